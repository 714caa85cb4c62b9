/*
 * fm_filesrc.h -- a file in rtl_sdr capture format (interleaved unsigned 8-bit I,Q, no header) behind
 * the librtlsdr streaming contract, so the reference's capture thread runs offline:
 *
 *     rtlsdr_read_async(dev, cb, ctx, buf_num, buf_len)     reference include/rtl-sdr.h:369-373,
 *                                                           src/librtlsdr.c:1867-1977
 *     typedef void(*rtlsdr_read_async_cb_t)(unsigned char *buf, uint32_t len, void *ctx)   rtl-sdr.h:340
 *     rtlsdr_cancel_async(dev)                              rtl-sdr.h:380
 *     rtlsdr_read_sync(dev, buf, len, n_read)               rtl-sdr.h:336
 *
 * Same callback type, same blocking behaviour (read_async returns when the stream ends or is cancelled),
 * same default chunk (buf_len 0 -> 16*32*512 = 262144 bytes, librtlsdr.c:354-355), same return convention
 * (0 / negative).  Differences that make an offline run deterministic:
 *   - a trailing chunk shorter than buf_len is never delivered (the player's ring needs
 *     max % len == 0, rtl_fm_player.c:823, and its demod thread only takes whole blocks, :863)
 *   - the dongle cannot be slowed down and the player's ring overwrites on overrun (:821-834); a file can,
 *     so delivery is paced: by a fill counter the consumer exposes (filesrc_set_backpressure), by wall
 *     clock (filesrc_set_realtime), or not at all.
 * Host-only C; no CUDA.
 */
#ifndef FM_FILESRC_H
#define FM_FILESRC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FILESRC_DEFAULT_BUF_LENGTH (16 * 32 * 512) /* librtlsdr.c:354-355 */

typedef struct filesrc_dev filesrc_dev_t;
typedef void (*filesrc_read_async_cb_t)(unsigned char *buf, uint32_t len, void *ctx);

/* path "-" reads stdin.  Returns 0 or -1 (errno set), like rtlsdr_open. */
int filesrc_open(filesrc_dev_t **dev, const char *path);
int filesrc_close(filesrc_dev_t *dev);

/* rtlsdr_set_sample_rate analogue: complex samples per second of the capture (8 * rate_in for the
 * player, rtl_fm_player.c:1053); only used by real-time pacing. */
int filesrc_set_sample_rate(filesrc_dev_t *dev, uint32_t rate);
/* speed 1.0 = deliver at the capture's own rate, 2.0 = twice as fast, 0 = no wall-clock pacing (default). */
int filesrc_set_realtime(filesrc_dev_t *dev, double speed);
/* Before delivering a chunk wait until *fill + len <= fill_max.  For the reference player:
 * fill = &_input_buffer_size, fill_max = _input_buffer_size_max (rtl_fm_player.h:65-70). */
int filesrc_set_backpressure(filesrc_dev_t *dev, const volatile uint32_t *fill, uint32_t fill_max);
/* Loop the file n times (0 = once, the default; UINT32_MAX = forever). */
int filesrc_set_loop(filesrc_dev_t *dev, uint32_t extra_passes);

/* Blocks; calls cb(buf, buf_len, ctx) for every whole chunk in file order from the calling thread.
 * buf_num is accepted for signature compatibility and ignored (no USB transfers to queue).
 * Returns 0 at end of file or after filesrc_cancel_async, negative on I/O error. */
int filesrc_read_async(filesrc_dev_t *dev, filesrc_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len);
/* May be called from the callback or from another thread. */
int filesrc_cancel_async(filesrc_dev_t *dev);
/* Reads exactly len bytes unless the file ends first; *n_read receives the count. */
int filesrc_read_sync(filesrc_dev_t *dev, void *buf, int len, int *n_read);

uint64_t filesrc_bytes_delivered(const filesrc_dev_t *dev);
uint64_t filesrc_chunks_delivered(const filesrc_dev_t *dev);

#ifdef __cplusplus
}
#endif
#endif
