/*
 * fm_wav.h -- the player's WAV output, byte for byte.
 *
 * Replaces InitWaveOut / CloseWaveOut (reference src/rtl_fm_player.c:1259-1328) and the file part of
 * output_thread_fn (:955-1005):
 *   - the file starts with the reference's fixed 260-byte header (include/rtl_fm_player.h:216-253): a
 *     44-byte PCM RIFF header hard-wired to 48000 Hz / 16 bit, stereo or mono, followed by 216 zero bytes
 *   - PCM reaches the file only in whole clusters of CIRCBUFFCLUSTER = 32768 bytes (h:54); what is left
 *     over at close never leaves the ring in the reference and is dropped here too (unless
 *     fm_wav_keep_tail() asks otherwise)
 *   - on close the RIFF size at byte 4 becomes file_size-8 and the data size at byte 40 becomes
 *     file_size-44 (:1266-1278) -- so the 216 padding bytes count as audio, as in the reference
 *   - path "-" writes to stdout and leaves the header's placeholder sizes alone (:1264, :1289-1296)
 */
#ifndef FM_WAV_H
#define FM_WAV_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_WAV_HEADER_BYTES 260
#define FM_WAV_CLUSTER_BYTES 32768

typedef struct fm_wav fm_wav;

/* mode 2 = stereo header, anything else = mono header (InitWaveOut's `mode`, :1313-1325).
 * Returns 0 or -1. */
int fm_wav_open(fm_wav **w, const char *path, int mode);
int fm_wav_write(fm_wav *w, const void *pcm, size_t bytes);
/* 1: also write the final partial cluster at close (NOT what the reference does). */
int fm_wav_keep_tail(fm_wav *w, int on);
int fm_wav_close(fm_wav *w);
/* The 260 header bytes as written at open. */
int fm_wav_header(int mode, unsigned char out[FM_WAV_HEADER_BYTES]);

#ifdef __cplusplus
}
#endif
#endif
