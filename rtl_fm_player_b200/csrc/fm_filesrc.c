/* fm_filesrc.c -- see include/fm_filesrc.h.  Replaces dongle_thread_fn's data source
 * (reference src/rtl_fm_player.c:839-853 -> src/librtlsdr.c:1867) with a file. */
#define _POSIX_C_SOURCE 200809L
#include "fm_filesrc.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct filesrc_dev {
    FILE *f;
    int is_stdin;
    uint32_t rate;
    double speed;
    const volatile uint32_t *fill;
    uint32_t fill_max;
    uint32_t extra_passes;
    volatile int cancel;
    volatile int running;
    uint64_t bytes, chunks;
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void nap(double s)
{
    struct timespec ts;
    if (s <= 0) return;
    ts.tv_sec = (time_t) s;
    ts.tv_nsec = (long) ((s - (double) ts.tv_sec) * 1e9);
    nanosleep(&ts, NULL);
}

int filesrc_open(filesrc_dev_t **dev, const char *path)
{
    filesrc_dev_t *d;
    if (!dev || !path) { errno = EINVAL; return -1; }
    *dev = NULL;
    d = calloc(1, sizeof *d);
    if (!d) return -1;
    if (strcmp(path, "-") == 0) { d->f = stdin; d->is_stdin = 1; }
    else d->f = fopen(path, "rb");
    if (!d->f) { free(d); return -1; }
    d->rate = 8 * 240000; /* the player's default capture rate, :1053 with DEFAULT_SAMPLE_RATE */
    *dev = d;
    return 0;
}

int filesrc_close(filesrc_dev_t *d)
{
    if (!d) return -1;
    if (d->f && !d->is_stdin) fclose(d->f);
    free(d);
    return 0;
}

int filesrc_set_sample_rate(filesrc_dev_t *d, uint32_t rate)
{
    if (!d || rate == 0) return -1;
    d->rate = rate;
    return 0;
}

int filesrc_set_realtime(filesrc_dev_t *d, double speed)
{
    if (!d || speed < 0) return -1;
    d->speed = speed;
    return 0;
}

int filesrc_set_backpressure(filesrc_dev_t *d, const volatile uint32_t *fill, uint32_t fill_max)
{
    if (!d) return -1;
    d->fill = fill;
    d->fill_max = fill_max;
    return 0;
}

int filesrc_set_loop(filesrc_dev_t *d, uint32_t extra_passes)
{
    if (!d) return -1;
    d->extra_passes = extra_passes;
    return 0;
}

int filesrc_read_async(filesrc_dev_t *d, filesrc_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len)
{
    unsigned char *buf;
    uint32_t passes_left;
    double t0, due = 0.0;
    int rc = 0;
    (void) buf_num;
    if (!d || !cb) return -1;
    if (buf_len == 0) buf_len = FILESRC_DEFAULT_BUF_LENGTH;
    if (buf_len % 512) return -1; /* librtlsdr: "must be multiple of 512" (rtl-sdr.h:366) */
    buf = malloc(buf_len);
    if (!buf) return -1;
    passes_left = d->extra_passes;
    d->cancel = 0;
    d->running = 1;
    t0 = now_s();
    while (!d->cancel) {
        const size_t got = fread(buf, 1, buf_len, d->f);
        if (got < buf_len) {
            if (ferror(d->f)) { rc = -1; break; }
            /* end of file: the short tail is dropped */
            if (passes_left && !d->is_stdin) {
                if (passes_left != UINT32_MAX) --passes_left;
                if (fseek(d->f, 0, SEEK_SET) != 0) { rc = -1; break; }
                continue;
            }
            break;
        }
        if (d->fill) /* the consumer's ring overwrites on overrun: wait for room instead */
            while (!d->cancel && (uint64_t) *d->fill + buf_len > d->fill_max) nap(0.001);
        if (d->speed > 0) {
            due += (double) (buf_len / 2) / ((double) d->rate * d->speed);
            nap(t0 + due - now_s());
        }
        if (d->cancel) break;
        cb(buf, buf_len, ctx);
        d->bytes += buf_len;
        d->chunks += 1;
    }
    d->running = 0;
    free(buf);
    return rc;
}

int filesrc_cancel_async(filesrc_dev_t *d)
{
    if (!d) return -1;
    d->cancel = 1;
    return 0;
}

int filesrc_read_sync(filesrc_dev_t *d, void *buf, int len, int *n_read)
{
    size_t got;
    if (!d || !buf || len < 0) return -1;
    got = fread(buf, 1, (size_t) len, d->f);
    if (n_read) *n_read = (int) got;
    if (got < (size_t) len && ferror(d->f)) return -1;
    d->bytes += got;
    return 0;
}

uint64_t filesrc_bytes_delivered(const filesrc_dev_t *d) { return d ? d->bytes : 0; }
uint64_t filesrc_chunks_delivered(const filesrc_dev_t *d) { return d ? d->chunks : 0; }
