/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Compiles the UNMODIFIED reference translation unit
 * (/root/reference/src/rtl_fm_player.c) by #include-ing it from where it lies,
 * with its main() renamed, and drives its own DSP functions offline:
 *
 *     rotate_90_u8_f32 | u8_f32  ->  full_demod        (rtl_fm_player.c:879-889)
 *
 * Nothing from the reference is copied into this repository.  The build recipe
 * (oracle/Makefile) writes only into oracle/_ref/.  Two artefacts:
 *   oracle/_ref/libfmref.so   : ref_* C API below, loaded by tests/ via ctypes
 *   oracle/_ref/ref_offline   : CLI (file in -> raw PCM out), used by tests/ and
 *                               by bench.py's reference arm / cpu_baseline leg
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may use
 * these.  The product path (rtl_fm_player_b200/) never touches them.
 */
#define _GNU_SOURCE
#define main ref_player_main
#include REF_SOURCE_PATH           /* e.g. "/root/reference/src/rtl_fm_player.c" */
#undef main

#include <stddef.h>
#include <sys/time.h>

/* Configuration mirror of the CLI flags that change numerics. */
struct ref_cfg {
    int rate_in;        /* -s  (rate_in = rate_out)                :1412-1415 */
    int rate_out2;      /* -r  (also output.rate)                  :1416-1419 */
    int mode;           /* lpr.mode 0/1/2                          :1183,1473,1486 */
    int size;           /* lpr.size                                :1184,1474,1487 */
    int offset_tuning;  /* -E offset                               :1454-1457 */
    double deemph;      /* seconds, 0 = off                        :1169 */
    float volume;       /* :1181 */
    int rate_out;       /* 0: = rate_in.  Else demod.rate_out as main leaves it under -o N: rate_in is the
                           oversampled rate (rate_in *= post_downsample, :1510), rate_out stays at -s */
};

struct ref_inst {
    struct demod_state d;   /* the reference's own state struct */
    int output_rate;
};

static int g_tables_ready = 0;

void ref_default_cfg(struct ref_cfg *c)
{
    /* defaults of demod_init (:1156-1195) */
    c->rate_in = DEFAULT_SAMPLE_RATE;
    c->rate_out2 = 48000;
    c->mode = 2;
    c->size = 90;
    c->offset_tuning = 0;
    c->deemph = DEEMPHASIS_FM_EU;
    c->volume = 0.4f;
    c->rate_out = 0;
}

void *ref_create(const struct ref_cfg *c)
{
    struct ref_inst *r = calloc(1, sizeof(*r));
    if (!r) return NULL;
    demod_init(&r->d);                       /* :1156 */
    r->d.rate_in = c->rate_in;
    r->d.rate_out = c->rate_out > 0 ? c->rate_out : c->rate_in;
    r->d.rate_out2 = c->rate_out2;
    r->d.lpr.mode = c->mode;
    r->d.lpr.size = c->size;
    r->d.offset_tuning = c->offset_tuning;
    r->d.deemph = c->deemph;
    r->d.volume = c->volume;
    r->output_rate = c->rate_out2 ? c->rate_out2 : (int) r->d.rate_out;   /* :1416, :1512-1514 */
    if (r->d.deemph)                         /* :1575-1578 */
        r->d.deemph_lambda = (float) exp(-1.0 / ((double) r->output_rate * r->d.deemph));
    if (!g_tables_ready) {                   /* :1601-1602 */
        init_u8_f32_table();
        init_lp_f32();
        g_tables_ready = 1;
    }
    init_lp_real_f32(&r->d);                 /* :1603 */
    return r;
}

/* the reference's own (de)init of the decoder rings on a live instance (:413-470), for the re-init test */
void ref_deinit_lp_real(void *h) { deinit_lp_real_f32(&((struct ref_inst *) h)->d); }
void ref_init_lp_real(void *h) { init_lp_real_f32(&((struct ref_inst *) h)->d); }

void ref_destroy(void *h)
{
    struct ref_inst *r = h;
    if (!r) return;
    deinit_lp_real_f32(&r->d);
    demod_cleanup(&r->d);
    free(r);
}

/* Copy the reference's designed coefficient tables out (for pinning the port). */
void ref_get_tables(void *h, float *fb16, float *fm, float *fp, float *fs, float *swf_cwf_lambda)
{
    struct ref_inst *r = h;
    int i;
    for (i = 0; i < 16; i++) fb16[i] = lp_filter_f32[i];
    for (i = 0; i < r->d.lpr.rsize; i++) {
        fm[i] = r->d.lpr.fm[i];
        fp[i] = r->d.lpr.fp[i];
        fs[i] = r->d.lpr.fs[i];
    }
    swf_cwf_lambda[0] = r->d.lpr.swf;
    swf_cwf_lambda[1] = r->d.lpr.cwf;
    swf_cwf_lambda[2] = r->d.deemph_lambda;
}

/*
 * One reference block through the very call sequence of demod_thread_fn
 * (:879-889).  Returns result_len (int16 values).  pcm must hold >= 131072.
 */
int ref_block(void *h, const uint8_t *iq, uint32_t len, int16_t *pcm)
{
    struct ref_inst *r = h;
    struct demod_state *d = &r->d;
    memcpy(d->buf, iq, len);
    d->buf_len = len;
    if (!d->offset_tuning) rotate_90_u8_f32(d); else u8_f32(d);
    full_demod(d);
    memcpy(pcm, d->result, (size_t) d->result_len * 2);
    return d->result_len;
}

/*
 * Same block, but calling the stage functions one by one in full_demod's order
 * (:758-788) so every intermediate can be copied out.  Any NULL dump is skipped.
 *   z   : after lp_f32        f32[len/8]   (interleaved I,Q at rate_in)
 *   dem : after fm_demod_f32  f32[len/16]
 *   lr  : after lp_real_f32   f32[n_lr]    (L,R interleaved or mono)
 *   de  : after deemph        f32[n_lr]
 * n_stage[0..1] receive len(z), len(dem); return value = result_len = n_lr.
 */
int ref_block_stages(void *h, const uint8_t *iq, uint32_t len, float *z, float *dem, float *lr,
                     float *de, int16_t *pcm, int *n_stage)
{
    struct ref_inst *r = h;
    struct demod_state *d = &r->d;
    memcpy(d->buf, iq, len);
    d->buf_len = len;
    if (!d->offset_tuning) rotate_90_u8_f32(d); else u8_f32(d);
    lp_f32(d);
    if (z) memcpy(z, d->lowpassed, (size_t) d->lp_len * 4);
    if (n_stage) n_stage[0] = d->lp_len;
    fm_demod_f32(d);
    if (dem) memcpy(dem, d->result, (size_t) d->result_len * 4);
    if (n_stage) n_stage[1] = d->result_len;
    if (d->rate_out2 > 0) lp_real_f32(d);
    if (lr) memcpy(lr, d->result, (size_t) d->result_len * 4);
    if (d->deemph) deemph_filter_f32(d);
    if (de) memcpy(de, d->result, (size_t) d->result_len * 4);
    convert_f32_s16(d);
    if (pcm) memcpy(pcm, d->result, (size_t) d->result_len * 2);
    return d->result_len;
}

/* Whole capture: full 262144-byte blocks only, the tail is dropped exactly as
 * demod_thread_fn does (:863-868).  Returns number of int16 values written. */
long ref_run(void *h, const uint8_t *iq, size_t n_bytes, int16_t *pcm, size_t pcm_cap)
{
    size_t off = 0;
    long n = 0;
    int16_t *tmp = malloc(MAXIMUM_BUF_LENGTH * sizeof(int16_t));
    while (off + MAXIMUM_BUF_LENGTH <= n_bytes) {
        int k = ref_block(h, iq + off, MAXIMUM_BUF_LENGTH, tmp);
        if ((size_t) (n + k) > pcm_cap) { free(tmp); return -1; }
        memcpy(pcm + n, tmp, (size_t) k * 2);
        n += k;
        off += MAXIMUM_BUF_LENGTH;
    }
    free(tmp);
    return n;
}

/* The reference's own struct inside an instance: what the drop-in shim (include/fm_dropin.h) is handed. */
void *ref_demod_state(void *h) { return &((struct ref_inst *) h)->d; }

/* The reference's WAV output: InitWaveOut (:1280-1328) writes the fixed header, the output thread writes
 * whole CIRCBUFFCLUSTER clusters (:955-1005), CloseWaveOut (:1259-1278) patches the sizes. */
int ref_wav_header(int mode, unsigned char *out260)
{
    if (mode == 2) memcpy(out260, _WAVHeaderStereo, sizeof(_WAVHeaderStereo));
    else memcpy(out260, _WAVHeaderMono, sizeof(_WAVHeaderMono));
    return (int) (mode == 2 ? sizeof(_WAVHeaderStereo) : sizeof(_WAVHeaderMono));
}

int ref_wav_write(const char *path, int mode, const unsigned char *pcm, size_t n_bytes)
{
    FILE *f = InitWaveOut((char *) path, mode);
    size_t off;
    if (!f) return -1;
    for (off = 0; off + CIRCBUFFCLUSTER <= n_bytes; off += CIRCBUFFCLUSTER)
        fwrite(pcm + off, sizeof(char), CIRCBUFFCLUSTER, f);
    CloseWaveOut(f);
    return 0;
}

/* sizeof/offsetof of the reference struct, for pinning include/fm_dropin.h. */
long ref_layout(int what)
{
    switch (what) {
    case 0: return (long) sizeof(struct demod_state);
    case 1: return (long) offsetof(struct demod_state, buf);
    case 2: return (long) offsetof(struct demod_state, buf_len);
    case 3: return (long) offsetof(struct demod_state, lowpassed);
    case 4: return (long) offsetof(struct demod_state, lp_len);
    case 5: return (long) offsetof(struct demod_state, lowpass_tb);
    case 6: return (long) offsetof(struct demod_state, result);
    case 7: return (long) offsetof(struct demod_state, result_len);
    case 8: return (long) offsetof(struct demod_state, offset_tuning);
    case 9: return (long) offsetof(struct demod_state, rate_in);
    case 10: return (long) offsetof(struct demod_state, rate_out);
    case 11: return (long) offsetof(struct demod_state, rate_out2);
    case 12: return (long) offsetof(struct demod_state, pre_r_f32);
    case 13: return (long) offsetof(struct demod_state, deemph);
    case 14: return (long) offsetof(struct demod_state, deemph_l_f32);
    case 15: return (long) offsetof(struct demod_state, deemph_lambda);
    case 16: return (long) offsetof(struct demod_state, volume);
    case 17: return (long) offsetof(struct demod_state, prev_lpr_index);
    case 18: return (long) offsetof(struct demod_state, lpr);
    case 19: return (long) offsetof(struct demod_state, rw);
    case 20: return (long) offsetof(struct demod_state, output_target);
    case 21: return (long) sizeof(struct lp_real);
    case 22: return (long) offsetof(struct lp_real, swf);
    case 23: return (long) offsetof(struct lp_real, pos);
    case 24: return (long) offsetof(struct lp_real, mode);
    case 25: return (long) offsetof(struct demod_state, post_downsample);
    case 26: return (long) offsetof(struct demod_state, exit_flag);
    }
    return -1;
}

/* ==========================================================================================
 * The reference's own THREADS, offline.
 *
 * ref_player_run() does what main() does between option parsing and the key loop (:1382-1385,
 * :1575-1578, :1601-1618, :1647-1665) and then lets the reference's dongle_thread_fn (:839),
 * demod_thread_fn (:855) and output_thread_fn (:935) run until the capture file is exhausted; the WAV
 * is written by the reference's InitWaveOut / output thread / CloseWaveOut.  librtlsdr is replaced by
 * a capture-file source handed in as two function pointers with the rtlsdr_read_async /
 * rtlsdr_cancel_async contract (the product's include/fm_filesrc.h; this file does not link it).
 *
 * The threads call rotate_90_u8_f32 / u8_f32 / full_demod (and this function the init_* trio) through
 * the PLT of this shared object.  Loaded on its own, they bind to the reference's CPU code.  With
 * rtl_fm_player_b200/libfmb.so loaded RTLD_GLOBAL first (what LD_PRELOAD does for the real player) they
 * bind to the CUDA drop-in of include/fm_dropin.h -- the unmodified player, demodulating on the GPU.
 * ========================================================================================== */
typedef int (*ref_read_async_fn)(void *src, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len);
typedef int (*ref_cancel_async_fn)(void *src);
static ref_read_async_fn g_src_read;
static ref_cancel_async_fn g_src_cancel;
static void *g_src;

int rtlsdr_read_async(rtlsdr_dev_t *dev, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t buf_num, uint32_t buf_len)
{
    (void) dev;
    return g_src_read ? g_src_read(g_src, cb, ctx, buf_num, buf_len) : -1;
}
int rtlsdr_cancel_async(rtlsdr_dev_t *dev)
{
    (void) dev;
    return g_src_cancel ? g_src_cancel(g_src) : 0;
}
int SDL_QueueAudio(SDL_AudioDeviceID dev, const void *data, uint32_t len) { (void) dev; (void) data; (void) len; return 0; }
const char *SDL_GetError(void) { return ""; }

/* address of the input ring's fill counter, for the file source's back-pressure (the ring overwrites on
 * overrun, :821-834) */
volatile uint32_t *ref_player_input_fill(uint32_t *fill_max)
{
    if (fill_max) *fill_max = _input_buffer_size_max;
    return (volatile uint32_t *) &_input_buffer_size;
}

int ref_player_run(const struct ref_cfg *c, const char *wav_path, ref_read_async_fn read_async,
                   ref_cancel_async_fn cancel_async, void *src)
{
    int spins;
    g_src_read = read_async; g_src_cancel = cancel_async; g_src = src;
    _do_exit = 0;
    _input_buffer_rpos = _input_buffer_wpos = _input_buffer_size = 0;
    _output_buffer_rpos = _output_buffer_wpos = _output_buffer_size = 0;
    _circbufferslots = 8;                            /* main: 180 MiB (:1363-1364); 8 slots do for a file */
    _circbuffeshift = 0;
    _circbuffer = (char *) malloc((size_t) _circbufferslots * CIRCBUFFCLUSTER);
    if (!_circbuffer) return -1;

    dongle_init(&dongle);                            /* :1382-1385 */
    demod_init(&demod);
    output_init(&output);
    controller_init(&controller);
    demod.rate_in = c->rate_in;                      /* the flags that change numerics, as ref_create */
    demod.rate_out = c->rate_out > 0 ? c->rate_out : c->rate_in;
    demod.rate_out2 = c->rate_out2;
    demod.lpr.mode = c->mode;
    demod.lpr.size = c->size;
    demod.offset_tuning = c->offset_tuning;
    demod.deemph = c->deemph;
    demod.volume = c->volume;
    demod.output_target = &output;
    output.rate = c->rate_out2 ? c->rate_out2 : (int) demod.rate_out;
    if (demod.deemph)                                /* :1575-1578 */
        demod.deemph_lambda = (float) exp(-1.0 / ((double) output.rate * demod.deemph));

    init_u8_f32_table();                             /* :1601-1603 */
    init_lp_f32();
    init_lp_real_f32(&demod);

    output.filename = (char *) wav_path;             /* :1647-1656 */
    output.file = InitWaveOut(output.filename, demod.lpr.mode);
    if (!output.file) { free(_circbuffer); return -2; }
    _audio_muted = 1;                                /* no audio device: the SDL_QueueAudio branch is skipped */
    _isStartStream = true;                           /* :1665 */

    pthread_create(&output.thread, NULL, output_thread_fn, (void *) (&output));   /* :1614-1618 */
    pthread_create(&demod.thread, NULL, demod_thread_fn, (void *) (&demod));
    pthread_create(&dongle.thread, NULL, dongle_thread_fn, (void *) (&dongle));

    pthread_join(dongle.thread, NULL);               /* the file is exhausted */
    /* let the rings drain: whole blocks in, whole clusters out (what is left never leaves the rings, as
     * in the player) */
    for (spins = 0; spins < 20000 && (_input_buffer_size >= MAXIMUM_BUF_LENGTH || _output_buffer_size >= CIRCBUFFCLUSTER); spins++)
        usleep(1000);
    usleep(20000);                                   /* a block or cluster that was in flight */
    for (spins = 0; spins < 20000 && (_input_buffer_size >= MAXIMUM_BUF_LENGTH || _output_buffer_size >= CIRCBUFFCLUSTER); spins++)
        usleep(1000);
    _do_exit = 1;                                    /* the X key, :1846 */
    pthread_join(demod.thread, NULL);
    pthread_join(output.thread, NULL);

    output.filename = 0;                             /* :1856-1859 */
    CloseWaveOut(output.file);
    demod_cleanup(&demod);
    output_cleanup(&output);
    controller_cleanup(&controller);
    free(_circbuffer);
    _circbuffer = NULL;
    return 0;
}

#ifdef REF_CLI
static double now_s(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/*
 * ref_offline [-X|-Y] [-E offset] [-s rate] [-r rate] [-m mode] [-z size] [-n repeat] in.u8 out.pcm
 * Loops fread(262144) -> rotate|u8 -> full_demod -> fwrite, like demod_thread_fn
 * without the rings.  With -n N the capture is preloaded and demodulated N times
 * back to back (state carried on) with no output, for timing; prints one line:
 *   REFTIME samples=<IQ samples> dsp_s=<seconds in rotate+full_demod> wall_s=<...>
 */
int main(int argc, char **argv)
{
    struct ref_cfg c;
    int i, repeat = 0;
    const char *in = NULL, *out = NULL;
    ref_default_cfg(&c);
    for (i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-X")) { c.rate_in = 192000; c.rate_out2 = 48000; c.mode = 2; c.size = 90; }
        else if (!strcmp(argv[i], "-Y")) { c.rate_in = 192000; c.rate_out2 = 48000; c.mode = 1; c.size = 128; }
        else if (!strcmp(argv[i], "-E") && i + 1 < argc) { if (!strcmp(argv[++i], "offset")) c.offset_tuning = 1; }
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) c.rate_in = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) c.rate_out2 = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) { /* main: rate_in *= post_downsample (:1422, :1510) */
            const int n = atoi(argv[++i]);
            if (n > 1) { c.rate_out = c.rate_in; c.rate_in *= n; }
        }
        else if (!strcmp(argv[i], "-m") && i + 1 < argc) c.mode = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-z") && i + 1 < argc) c.size = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-n") && i + 1 < argc) repeat = atoi(argv[++i]);
        else if (!in) in = argv[i];
        else if (!out) out = argv[i];
    }
    if (!in) { fprintf(stderr, "usage: ref_offline [flags] in.u8 [out.pcm]\n"); return 2; }
    void *h = ref_create(&c);
    FILE *fi = fopen(in, "rb");
    if (!fi) { perror(in); return 1; }
    fseek(fi, 0, SEEK_END);
    long sz = ftell(fi);
    fseek(fi, 0, SEEK_SET);
    uint8_t *iq = malloc((size_t) sz);
    if (fread(iq, 1, (size_t) sz, fi) != (size_t) sz) { perror("fread"); return 1; }
    fclose(fi);
    int16_t *pcm = malloc(MAXIMUM_BUF_LENGTH * 2);
    double t_wall0 = now_s(), dsp = 0;
    long samples = 0;
    FILE *fo = (out && !repeat) ? fopen(out, "wb") : NULL;
    int rep, nrep = repeat ? repeat : 1;
    for (rep = 0; rep < nrep; rep++) {
        long off;
        for (off = 0; off + MAXIMUM_BUF_LENGTH <= sz; off += MAXIMUM_BUF_LENGTH) {
            double t0 = now_s();
            int k = ref_block(h, iq + off, MAXIMUM_BUF_LENGTH, pcm);
            dsp += now_s() - t0;
            samples += MAXIMUM_BUF_LENGTH / 2;
            if (fo) fwrite(pcm, 2, (size_t) k, fo);
        }
    }
    if (fo) fclose(fo);
    printf("REFTIME samples=%ld dsp_s=%.6f wall_s=%.6f\n", samples, dsp, now_s() - t_wall0);
    ref_destroy(h);
    return 0;
}
#endif
