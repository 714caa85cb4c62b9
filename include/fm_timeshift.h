/*
 * fm_timeshift.h -- the player's timeshift ring, for a batch of channels.
 *
 * Replaces the ring part of output_thread_fn (reference src/rtl_fm_player.c:935-1030) and its sizing
 * in main (:1362-1365, :1543-1549): PCM is stored in whole clusters of CIRCBUFFCLUSTER = 32768 bytes
 * (include/rtl_fm_player.h:54) in a ring of `slots` clusters; every new cluster is written at the
 * ring's "bottom" slot (:964) and the cluster that is PLAYED (SDL_QueueAudio :1000, fwrite :1004) is
 * the one `shift` clusters behind it, with the reference's clamping of the shift:
 *     shift < 0                         -> 0                     (:982)
 *     ring not yet wrapped, shift > bottom -> bottom             (:985-987: cannot go back before the start)
 *     ring wrapped, shift > slots-2     -> slots-2               (:988-991)
 *     out = bottom - shift, wrapped by  slots - (shift - bottom) (:994-997)
 *     bottom advances and wraps after the playback slot was taken (:1007-1010)
 * The batched form keeps one ring per channel with a common bottom (all channels advance by one
 * cluster per push, as the batched demodulator produces them) and one shift, like the keyboard-driven
 * _circbuffeshift of the player (:1788 ff.).  Pure host memory; no compute.
 */
#ifndef FM_TIMESHIFT_H
#define FM_TIMESHIFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_TS_CLUSTER_BYTES 32768

typedef struct fm_timeshift fm_timeshift;

/* _circbufferslots = (kbytes * 1024) / CIRCBUFFCLUSTER (:1364); the player's default is 180 MiB. */
int fm_timeshift_slots_for_kbytes(long kbytes);
/* Returns 0, or -1 (bad argument / out of memory, the player's "Can't allocate memmory for timeshift", :1547). */
int fm_timeshift_create(fm_timeshift **ts, int slots, int n_streams);
void fm_timeshift_destroy(fm_timeshift *ts);
/*
 * Store one cluster per channel (`clusters` = [n_streams][cluster_pitch bytes], cluster_pitch >= 32768)
 * and select the playback cluster.  *shift is the requested shift in clusters and comes back clamped, as
 * the reference clamps _circbuffeshift in place.  If `out` is non-NULL the playback cluster of every
 * channel is copied to out[n_streams][32768].  Returns the playback slot index (>= 0) or -1.
 */
int fm_timeshift_push(fm_timeshift *ts, const void *clusters, size_t cluster_pitch, int *shift, void *out);
/* Playback cluster of one channel after the last push (valid until the next push). */
const void *fm_timeshift_playback(const fm_timeshift *ts, int stream);
/* Ring bookkeeping, for tests and status lines: bottom slot (next write), wrapped flag, slots. */
int fm_timeshift_state(const fm_timeshift *ts, int *bottom, int *wrapped, int *slots);

#ifdef __cplusplus
}
#endif
#endif
