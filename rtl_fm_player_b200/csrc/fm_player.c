/*
 * fm_player.c -- `fmb_player`: offline, batched counterpart of the reference's capture -> demod -> file
 * pipeline (reference src/rtl_fm_player.c: dongle_thread_fn :839-853, demod_thread_fn :855-933, the file
 * branch of output_thread_fn :955-1005), in plain C on top of the C ABI:
 *
 *   one capture file per FM channel  --filesrc_read_async (one thread per channel, 262144-byte chunks)-->
 *   pinned batch buffer [channel][block]  --fmb_multi_submit / fmb_multi_wait (per device: H2D, CUDA kernels, D2H)-->
 *   per-channel PCM  --fm_wav (reference WAV header + 32768-byte clusters) or raw .pcm-->
 *
 * The channels are sharded by index over the devices given with -d (default: device 0), one worker thread per
 * device (include/fmb_multi.h); every device reads its slice of the one pinned input buffer and writes its
 * slice of the one pinned PCM buffer, so the host gathers nothing.
 *
 * Flags follow the reference's where they exist (-X -Y -s -r -E offset, :1389-1508).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fm_filesrc.h"
#include "fm_wav.h"
#include "fmb.h"
#include "fmb_multi.h"

#define NBUF FMB_PIPE_DEPTH

struct player;
struct chan {
    struct player *p;
    int index;
    filesrc_dev_t *src;
    pthread_t thread;
    long produced;      /* blocks copied into the batch buffers */
    int finished;       /* file ended */
    long n_blocks;      /* blocks this channel delivered in total (valid once finished) */
    fm_wav *wav;
    FILE *raw;
};

struct player {
    fmb_config cfg;
    int n;
    struct chan *ch;
    uint8_t *iq[NBUF];      /* pinned [n][block_bytes] */
    int16_t *pcm[NBUF];     /* pinned [n][pcm_pitch]   */
    size_t pcm_pitch;
    long consumed;          /* blocks whose buffers may be refilled: block k may be written iff k < consumed + NBUF */
    pthread_mutex_t mu;
    pthread_cond_t cv;
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* the rtlsdr_callback of this program (reference :790-837): copy the chunk into the channel's row of the
 * batch buffer for its next block; blocks while all NBUF buffers are still in flight (back-pressure). */
static void chunk_cb(unsigned char *buf, uint32_t len, void *ctx)
{
    struct chan *c = ctx;
    struct player *p = c->p;
    pthread_mutex_lock(&p->mu);
    while (c->produced >= p->consumed + NBUF) pthread_cond_wait(&p->cv, &p->mu);
    pthread_mutex_unlock(&p->mu);
    memcpy(p->iq[c->produced % NBUF] + (size_t) c->index * (size_t) p->cfg.block_bytes, buf, len);
    pthread_mutex_lock(&p->mu);
    c->produced++;
    pthread_cond_broadcast(&p->cv);
    pthread_mutex_unlock(&p->mu);
}

static void *reader(void *arg)
{
    struct chan *c = arg;
    struct player *p = c->p;
    if (filesrc_read_async(c->src, chunk_cb, c, 0, (uint32_t) p->cfg.block_bytes) != 0)
        fprintf(stderr, "channel %d: read error\n", c->index);
    pthread_mutex_lock(&p->mu);
    c->finished = 1;
    c->n_blocks = c->produced;
    pthread_cond_broadcast(&p->cv);
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static void usage(void)
{
    fprintf(stderr,
            "fmb_player -- batched offline FM demodulator (B200), one rtl_sdr-format uint8 IQ capture per channel\n"
            "usage: fmb_player [-X | -Y] [-s rate_in] [-r rate_out2] [-E offset] [-m lpr_mode] [-z lpr_size]\n"
            "                  [-v volume] [-w] [-o out_dir] [-d cuda_devices] [-P exact|fma] [-R speed] capture.u8 ...\n"
            "  -X  stereo preset 192k/48k, 90 taps   -Y  mono preset 192k/48k, 128 taps   (as rtl_fm_player)\n"
            "  -w  write <name>.wav with the reference's header and 32768-byte cluster rule (default: raw <name>.pcm)\n"
            "  -d  CUDA devices to shard the channels over, e.g. 0-7 or 0,2,4 (default 0); at least one channel per device\n"
            "  -R  pace the captures at `speed` x real time (default: as fast as the GPU takes them)\n");
}

static const char *base_name(const char *path)
{
    const char *s = strrchr(path, '/');
    return s ? s + 1 : path;
}

int main(int argc, char **argv)
{
    struct player P;
    const char *out_dir = NULL;
    int wav = 0, i, rc, first_file = -1;
    double speed = 0.0, t0;
    long k, submitted = 0, total_blocks = 0;
    int tickets[NBUF];
    int *n_outs;
    int devices[64], n_devices = 1;
    fmb_multi *h = NULL;

    memset(&P, 0, sizeof P);
    devices[0] = 0;
    fmb_default_config(&P.cfg);
    for (i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "-X")) fmb_preset_stereo_192k(&P.cfg);
        else if (!strcmp(a, "-Y")) fmb_preset_mono_192k(&P.cfg);
        else if (!strcmp(a, "-s") && i + 1 < argc) P.cfg.rate_in = atoi(argv[++i]);
        else if (!strcmp(a, "-r") && i + 1 < argc) P.cfg.rate_out2 = atoi(argv[++i]);
        else if (!strcmp(a, "-E") && i + 1 < argc) { if (!strcmp(argv[++i], "offset")) P.cfg.offset_tuning = 1; }
        else if (!strcmp(a, "-m") && i + 1 < argc) P.cfg.mode = atoi(argv[++i]);
        else if (!strcmp(a, "-z") && i + 1 < argc) P.cfg.size = atoi(argv[++i]);
        else if (!strcmp(a, "-v") && i + 1 < argc) P.cfg.volume = (float) atof(argv[++i]);
        else if (!strcmp(a, "-d") && i + 1 < argc) {
            n_devices = fmb_parse_device_list(argv[++i], devices, 64);
            if (n_devices < 1) { fprintf(stderr, "-d: %s\n", fmb_last_error()); return 2; }
        }
        else if (!strcmp(a, "-o") && i + 1 < argc) out_dir = argv[++i];
        else if (!strcmp(a, "-R") && i + 1 < argc) speed = atof(argv[++i]);
        else if (!strcmp(a, "-P") && i + 1 < argc) P.cfg.precision = !strcmp(argv[++i], "fma") ? FMB_PRECISION_FMA : FMB_PRECISION_EXACT;
        else if (!strcmp(a, "-w")) wav = 1;
        else if (!strcmp(a, "-h")) { usage(); return 0; }
        else if (a[0] == '-' && a[1]) { usage(); return 2; }
        else { first_file = i; break; }
    }
    if (first_file < 0) { usage(); return 2; }
    P.n = argc - first_file;
    P.cfg.n_streams = P.n;
    P.ch = calloc((size_t) P.n, sizeof *P.ch);
    n_outs = calloc((size_t) P.n, sizeof *n_outs);
    if (!P.ch || !n_outs) return 1;
    pthread_mutex_init(&P.mu, NULL);
    pthread_cond_init(&P.cv, NULL);

    if (n_devices > P.n) n_devices = P.n;          /* fewer channels than devices: use the first P.n */
    if (n_devices == 1) fmb_bind_thread_to_device_node(devices[0]); /* best effort: node-local pinned buffers */
    rc = fmb_multi_create(&P.cfg, devices, n_devices, &h);
    if (rc != FMB_OK) { fprintf(stderr, "fmb_multi_create: %s\n", fmb_last_error()); return 1; }
    P.pcm_pitch = ((size_t) fmb_multi_max_out_count(h) + 7) & ~(size_t) 7;
    for (i = 0; i < NBUF; ++i) {
        if (fmb_host_alloc((void **) &P.iq[i], (size_t) P.n * (size_t) P.cfg.block_bytes) != FMB_OK ||
            fmb_host_alloc((void **) &P.pcm[i], (size_t) P.n * P.pcm_pitch * sizeof(int16_t)) != FMB_OK) {
            fprintf(stderr, "pinned allocation: %s\n", fmb_last_error());
            return 1;
        }
    }
    for (i = 0; i < P.n; ++i) {
        struct chan *c = &P.ch[i];
        char path[4096];
        const char *in = argv[first_file + i];
        c->p = &P; c->index = i;
        if (filesrc_open(&c->src, in) != 0) { perror(in); return 1; }
        filesrc_set_sample_rate(c->src, (uint32_t) (8 * P.cfg.rate_in)); /* capture_rate, :1053 */
        filesrc_set_realtime(c->src, speed);
        if (out_dir) snprintf(path, sizeof path, "%s/%s.%s", out_dir, base_name(in), wav ? "wav" : "pcm");
        else snprintf(path, sizeof path, "%s.%s", in, wav ? "wav" : "pcm");
        if (wav) { if (fm_wav_open(&c->wav, path, P.cfg.rate_out2 > 0 ? P.cfg.mode : 1) != 0) { perror(path); return 1; } }
        else { c->raw = fopen(path, "wb"); if (!c->raw) { perror(path); return 1; } }
    }
    t0 = now_s();
    for (i = 0; i < P.n; ++i) pthread_create(&P.ch[i].thread, NULL, reader, &P.ch[i]);

    /* the demod thread of this program: block k of every channel -> one fmb_submit */
    for (k = 0;; ++k) {
        int live = 0, n_out = 0;
        pthread_mutex_lock(&P.mu);
        for (;;) { /* wait until every channel has either delivered block k or ended */
            int ready = 1;
            live = 0;
            for (i = 0; i < P.n; ++i) {
                if (P.ch[i].produced > k) live++;
                else if (!P.ch[i].finished) ready = 0;
            }
            if (ready) break;
            pthread_cond_wait(&P.cv, &P.mu);
        }
        pthread_mutex_unlock(&P.mu);
        if (live > 0) {
            for (i = 0; i < P.n; ++i) /* ended channels get mid-scale bytes; their PCM is not written */
                if (P.ch[i].produced <= k) memset(P.iq[k % NBUF] + (size_t) i * (size_t) P.cfg.block_bytes, 127, (size_t) P.cfg.block_bytes);
            rc = fmb_multi_submit(h, P.iq[k % NBUF], (size_t) P.cfg.block_bytes, P.pcm[k % NBUF], P.pcm_pitch, &tickets[k % NBUF]);
            if (rc != FMB_OK) { fprintf(stderr, "fmb_multi_submit: %s\n", fmb_last_error()); return 1; }
            submitted = k + 1;
        }
        /* retire block k-(NBUF-1) (or everything left once the inputs have ended) */
        {
            long upto = live > 0 ? submitted - (NBUF - 1) : submitted, j;
            for (j = P.consumed; j < upto; ++j) {
                rc = fmb_multi_wait(h, tickets[j % NBUF], n_outs);   /* result_len of this block, per channel */
                if (rc != FMB_OK) { fprintf(stderr, "fmb_multi_wait: %s\n", fmb_last_error()); return 1; }
                n_out = n_outs[0];
                for (i = 0; i < P.n; ++i) {
                    struct chan *c = &P.ch[i];
                    const int16_t *src = P.pcm[j % NBUF] + (size_t) i * P.pcm_pitch;
                    if (c->finished && j >= c->n_blocks) continue;
                    if (c->wav) fm_wav_write(c->wav, src, (size_t) n_out * 2);
                    else fwrite(src, 2, (size_t) n_out, c->raw);
                    total_blocks++;
                }
                pthread_mutex_lock(&P.mu);
                P.consumed = j + 1;
                pthread_cond_broadcast(&P.cv);
                pthread_mutex_unlock(&P.mu);
            }
        }
        if (live == 0) break;
    }
    {
        const double dt = now_s() - t0;
        fprintf(stderr, "fmb_player: %d channel(s) on %d device(s), %ld channel-blocks (%.1f M IQ samples) in %.3f s = %.1f Msamples/s\n",
                P.n, n_devices, total_blocks, (double) total_blocks * (P.cfg.block_bytes / 2) * 1e-6, dt,
                (double) total_blocks * (P.cfg.block_bytes / 2) * 1e-6 / (dt > 0 ? dt : 1));
    }
    for (i = 0; i < P.n; ++i) {
        pthread_join(P.ch[i].thread, NULL);
        if (P.ch[i].wav) fm_wav_close(P.ch[i].wav);
        if (P.ch[i].raw) fclose(P.ch[i].raw);
        filesrc_close(P.ch[i].src);
    }
    for (i = 0; i < NBUF; ++i) { fmb_host_free(P.iq[i]); fmb_host_free(P.pcm[i]); }
    fmb_multi_destroy(h);
    free(P.ch);
    free(n_outs);
    return 0;
}
