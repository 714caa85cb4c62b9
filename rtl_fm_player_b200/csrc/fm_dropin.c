/*
 * fm_dropin.c -- the reference's demod entry points (include/fm_dropin.h) on top of the batched C ABI
 * (include/fmb.h) with n_streams = 1.  Plain C host code; all arithmetic happens in the CUDA kernels.
 * The reference's struct demod_state is accessed by the byte offsets in ref_layout.h (generated from the
 * reference header by oracle/gen_layout.c), never by a copied definition.
 */
#include "fm_dropin.h"

#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fmb_internal.h"
#include "ref_layout.h"

#define F(d, type, off) (*(type *) ((char *) (d) + (off)))
#define LPR(d, type, off) F(d, type, FMD_OFF_lpr + (off))

enum { CONV_NONE = 0, CONV_ROTATE = 1, CONV_PLAIN = 2 };

struct shim {
    struct demod_state *d;
    fmb_handle *h;
    fmb_config cfg;       /* configuration the handle was created with */
    int conv;             /* which of rotate_90_u8_f32 / u8_f32 ran since the last full_demod */
    int pos;              /* lpr.pos the reference would have now */
    uint8_t *pin_iq;      /* pinned staging */
    int16_t *pin_pcm;
    size_t pcm_cap;
    struct shim *next;
};

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static struct shim *g_list;
static int g_strict, g_device;
static void (*g_hook)(const char *);

static void fail(const char *what, int rc)
{
    char msg[768];
    snprintf(msg, sizeof msg, "fm_dropin: %s failed (%d): %s", what, rc, fmb_last_error());
    if (g_hook) { g_hook(msg); return; }
    fprintf(stderr, "%s\nThis build demodulates on the GPU only; there is no CPU fallback.\n", msg);
    abort();
}

static struct shim *find(struct demod_state *d, int create)
{
    struct shim *s;
    pthread_mutex_lock(&g_mu);
    for (s = g_list; s; s = s->next)
        if (s->d == d) break;
    if (!s && create) {
        s = calloc(1, sizeof *s);
        if (s) { s->d = d; s->next = g_list; g_list = s; }
    }
    pthread_mutex_unlock(&g_mu);
    return s;
}

static void drop_handle(struct shim *s)
{
    if (s->h) { fmb_destroy(s->h); s->h = NULL; }
    if (s->pin_iq) { fmb_host_free(s->pin_iq); s->pin_iq = NULL; }
    if (s->pin_pcm) { fmb_host_free(s->pin_pcm); s->pin_pcm = NULL; }
}

/* ---- state: reference representation <-> fmb_stream_state (time order, oldest first) ---- */
static int import_state(struct shim *s)
{
    struct demod_state *d = s->d;
    fmb_stream_state st;
    const int size = LPR(d, int, FMD_LPR_OFF_size), mode = LPR(d, int, FMD_LPR_OFF_mode);
    const float *br = LPR(d, float *, FMD_LPR_OFF_br), *bm = LPR(d, float *, FMD_LPR_OFF_bm),
                *bs = LPR(d, float *, FMD_LPR_OFF_bs);
    int pos = LPR(d, int, FMD_LPR_OFF_pos), i;
    memset(&st, 0, sizeof st);
    memcpy(st.lowpass_tb, &F(d, float, FMD_OFF_lowpass_tb), sizeof st.lowpass_tb);
    st.pre_r = F(d, float, FMD_OFF_pre_r_f32);
    st.pre_j = F(d, float, FMD_OFF_pre_j_f32);
    if (size > 0 && size <= FMB_HIST && br && pos >= 0 && pos < size) {
        /* ring[(pos + i) % size], i = 0..size-1, is oldest..newest (lp_real_f32 :500-566) */
        for (i = 0; i < size; ++i) {
            const int r = (pos + i) % size, t = FMB_HIST - size + i;
            st.br[t] = br[r];
            if (mode == 2 && bm && bs) { st.bm[t] = bm[r]; st.bs[t] = bs[r]; }
        }
    }
    st.pp = LPR(d, float, FMD_LPR_OFF_pp);
    st.deemph_l = F(d, float, FMD_OFF_deemph_l_f32);
    st.deemph_r = F(d, float, FMD_OFF_deemph_r_f32);
    st.raw_valid = 0; /* the reference keeps converted floats, not raw bytes */
    s->pos = pos;
    return fmb_set_state(s->h, 0, 1, &st, F(d, int, FMD_OFF_prev_lpr_index), 0);
}

static int export_state(struct shim *s)
{
    struct demod_state *d = s->d;
    fmb_stream_state st;
    int phase = 0, i, rc;
    const int size = LPR(d, int, FMD_LPR_OFF_size), mode = LPR(d, int, FMD_LPR_OFF_mode);
    float *br = LPR(d, float *, FMD_LPR_OFF_br), *bm = LPR(d, float *, FMD_LPR_OFF_bm),
          *bs = LPR(d, float *, FMD_LPR_OFF_bs);
    if (!s->h) return FMB_OK;
    rc = fmb_get_state(s->h, 0, 1, &st, &phase, NULL);
    if (rc != FMB_OK) return rc;
    memcpy(&F(d, float, FMD_OFF_lowpass_tb), st.lowpass_tb, sizeof st.lowpass_tb);
    F(d, float, FMD_OFF_pre_r_f32) = st.pre_r;
    F(d, float, FMD_OFF_pre_j_f32) = st.pre_j;
    if (s->cfg.rate_out2 > 0 && mode >= 1 && size > 0 && size <= FMB_HIST && br) {
        for (i = 0; i < size; ++i) {
            const int r = (s->pos + i) % size, t = FMB_HIST - size + i;
            br[r] = st.br[t];
            if (mode == 2 && bm && bs) { bm[r] = st.bm[t]; bs[r] = st.bs[t]; }
        }
        LPR(d, int, FMD_LPR_OFF_pos) = s->pos;
        if (mode == 2) LPR(d, float, FMD_LPR_OFF_pp) = st.pp;
    }
    if (s->cfg.rate_out2 > 0) F(d, int, FMD_OFF_prev_lpr_index) = phase;
    if (s->cfg.deemph != 0.0) {
        F(d, float, FMD_OFF_deemph_l_f32) = st.deemph_l;
        if (mode == 2) F(d, float, FMD_OFF_deemph_r_f32) = st.deemph_r; /* :692-706 pairs L,R whenever lpr.mode == 2 */
    }
    return FMB_OK;
}

static void config_from_struct(struct shim *s, fmb_config *c, uint32_t buf_len)
{
    struct demod_state *d = s->d;
    fmb_default_config(c);
    c->rate_in = F(d, int, FMD_OFF_rate_in);         /* the filters' rate (init_lp_real_f32, :419-429) */
    c->rate_out = F(d, int, FMD_OFF_rate_out);       /* the resampler's fast rate (lp_real_f32, :485); differs from
                                                        rate_in under -o N (main: rate_in *= post_downsample, :1510) */
    c->rate_out2 = F(d, int, FMD_OFF_rate_out2);
    c->mode = LPR(d, int, FMD_LPR_OFF_mode);
    c->size = LPR(d, int, FMD_LPR_OFF_size);
    c->offset_tuning = (s->conv == CONV_PLAIN);
    c->deemph = F(d, double, FMD_OFF_deemph);
    c->deemph_lambda = c->deemph != 0.0 ? F(d, float, FMD_OFF_deemph_lambda) : 0.0f; /* main computes it, :1577 */
    c->volume = F(d, float, FMD_OFF_volume);
    c->n_streams = 1;
    c->block_bytes = (int) buf_len;
    c->device = g_device;
    c->precision = FMB_PRECISION_EXACT;
    c->emulate_inplace_quirk = 1;
}

static int same_config(const fmb_config *a, const fmb_config *b)
{
    return a->rate_in == b->rate_in && a->rate_out == b->rate_out && a->rate_out2 == b->rate_out2 && a->mode == b->mode && a->size == b->size &&
           a->offset_tuning == b->offset_tuning && a->deemph == b->deemph && a->deemph_lambda == b->deemph_lambda &&
           a->block_bytes == b->block_bytes && a->device == b->device;
}

/* ---- the reference's entry points ---- */

/* The LUT and the channel filter live in the kernels' constant parameters; if the host program still
 * defines the reference's global tables (rtl_fm_player.h:274-281) they are filled for anyone who looks. */
extern float u8_f32_table[2][256] __attribute__((weak));
extern float lp_filter_f32[16] __attribute__((weak));

/* address of a weak symbol that may be undefined (NULL), hidden from constant folding */
static void *weak_addr(void *p)
{
    __asm__ volatile("" : "+r"(p));
    return p;
}

void init_u8_f32_table(void)
{
    int i;
    if (!weak_addr(u8_f32_table)) return;
    for (i = 0; i < 256; ++i) { /* :195-204 */
        u8_f32_table[0][i] = ((float) i - 127.5f) / 128.0f;
        u8_f32_table[1][i] = ((float) i - 127.5f) / -128.0f;
    }
}

void init_lp_f32(void)
{
    fmb_config c;
    fmb_tables t;
    if (!weak_addr(lp_filter_f32)) return;
    fmb_default_config(&c);
    if (fmb_design_tables(&c, &t) == FMB_OK) memcpy(lp_filter_f32, t.chan, sizeof t.chan);
}

void init_lp_real_f32(struct demod_state *d)
{
    fmb_config c;
    fmb_tables t;
    int rc;
    const int size = LPR(d, int, FMD_LPR_OFF_size), taps = size >> 1;
    {
        /* Called again on a struct that already has a GPU context: the reference's callocs below start the
         * decoder rings from zero (:430-436) and leave everything else (lowpass_tb, pre_r/j, de-emphasis, resampler
         * phase) alone.  So: bring the struct up to date, then forget the GPU copy -- the context is rebuilt
         * from the struct, fresh rings included, at the next full_demod. */
        struct shim *old = find(d, 0);
        if (old) {
            if (old->h && (rc = export_state(old)) != FMB_OK) fail("init_lp_real_f32: exporting the GPU state", rc);
            drop_handle(old);
            old->pos = 0;
        }
    }
    fmb_default_config(&c);
    c.rate_in = F(d, int, FMD_OFF_rate_in);
    c.size = size;
    rc = fmb_design_tables(&c, &t);
    if (rc != FMB_OK) { fail("init_lp_real_f32: filter design (lpr.size must be 2..128)", rc); return; }
    LPR(d, int, FMD_LPR_OFF_rsize) = taps;
    LPR(d, float, FMD_LPR_OFF_swf) = t.swf;
    LPR(d, float, FMD_LPR_OFF_cwf) = t.cwf;
    LPR(d, float, FMD_LPR_OFF_pp) = 0.0f;
    LPR(d, float *, FMD_LPR_OFF_br) = calloc((size_t) size, 4);
    LPR(d, float *, FMD_LPR_OFF_bm) = calloc((size_t) size, 4);
    LPR(d, float *, FMD_LPR_OFF_bs) = calloc((size_t) size, 4);
    LPR(d, float *, FMD_LPR_OFF_fm) = calloc((size_t) taps, 4);
    LPR(d, float *, FMD_LPR_OFF_fp) = calloc((size_t) taps, 4);
    LPR(d, float *, FMD_LPR_OFF_fs) = calloc((size_t) taps, 4);
    LPR(d, int, FMD_LPR_OFF_pos) = 0;
    if (LPR(d, float *, FMD_LPR_OFF_fm)) memcpy(LPR(d, float *, FMD_LPR_OFF_fm), t.fm, (size_t) taps * 4);
    if (LPR(d, float *, FMD_LPR_OFF_fp)) memcpy(LPR(d, float *, FMD_LPR_OFF_fp), t.fp, (size_t) taps * 4);
    if (LPR(d, float *, FMD_LPR_OFF_fs)) memcpy(LPR(d, float *, FMD_LPR_OFF_fs), t.fs, (size_t) taps * 4);
    find(d, 1);
}

void fm_dropin_release(struct demod_state *d)
{
    struct shim *s, **pp;
    pthread_mutex_lock(&g_mu);
    for (pp = &g_list; (s = *pp) != NULL; pp = &s->next)
        if (s->d == d) { *pp = s->next; break; }
    pthread_mutex_unlock(&g_mu);
    if (!s) return;
    if (s->h) export_state(s);
    drop_handle(s);
    free(s);
}

void deinit_lp_real_f32(struct demod_state *d)
{
    static const int ptrs[6] = {FMD_LPR_OFF_br, FMD_LPR_OFF_bm, FMD_LPR_OFF_bs, FMD_LPR_OFF_fm, FMD_LPR_OFF_fp, FMD_LPR_OFF_fs};
    int i;
    fm_dropin_release(d);
    LPR(d, int, FMD_LPR_OFF_rsize) = 0; /* :455-470 */
    for (i = 0; i < 6; ++i) {
        free(LPR(d, float *, ptrs[i]));
        LPR(d, float *, ptrs[i]) = NULL;
    }
}

static void note_conversion(struct demod_state *d, int conv)
{
    struct shim *s = find(d, 1);
    if (!s) { fail("out of memory", FMB_ERR_NOMEM); return; }
    s->conv = conv;
    F(d, int, FMD_OFF_lp_len) = (int) F(d, uint32_t, FMD_OFF_buf_len); /* :225, :238 */
}

/* The conversion itself is fused into the demod kernel (it reads d->buf's bytes directly). */
void rotate_90_u8_f32(struct demod_state *d) { note_conversion(d, CONV_ROTATE); }
void u8_f32(struct demod_state *d) { note_conversion(d, CONV_PLAIN); }

void full_demod(struct demod_state *d)
{
    struct shim *s = find(d, 1);
    fmb_config c;
    const uint32_t buf_len = F(d, uint32_t, FMD_OFF_buf_len);
    int rc, n_out = 0;
    if (!s) { fail("out of memory", FMB_ERR_NOMEM); return; }
    if (s->conv == CONV_NONE) {
        fail("full_demod without a preceding rotate_90_u8_f32()/u8_f32() on this block", FMB_ERR_STATE);
        return;
    }
    /* d->post_downsample > 1 is an empty block in full_demod itself (:776-779, "for float not implemented"): its
     * only effect on this path is main()'s rate_in *= post_downsample (:1510), which config_from_struct honours by
     * taking the filters' rate from rate_in and the resampler's from rate_out. */
    config_from_struct(s, &c, buf_len);
    if (!s->h || !same_config(&c, &s->cfg)) {
        if (s->h) { export_state(s); drop_handle(s); }
        rc = fmb_create(&c, &s->h);
        if (rc != FMB_OK) { s->h = NULL; fail("fmb_create", rc); return; }
        s->cfg = c;
        s->pcm_cap = (size_t) fmb_max_out_count(s->h);
        rc = fmb_host_alloc((void **) &s->pin_iq, (size_t) c.block_bytes);
        if (rc == FMB_OK) rc = fmb_host_alloc((void **) &s->pin_pcm, (s->pcm_cap + 8) * sizeof(int16_t));
        if (rc == FMB_OK) rc = import_state(s);
        if (rc != FMB_OK) {
            /* with an error hook installed fail() returns: leave no half-built context behind, so that the
             * next call starts over (and reports again) instead of copying into a NULL staging buffer or
             * continuing silently from zero state */
            fail("setting up the GPU context", rc);
            drop_handle(s);
            return;
        }
    } else if (c.volume != s->cfg.volume) {
        rc = fmb_set_volume(s->h, c.volume);
        if (rc != FMB_OK) { fail("fmb_set_volume", rc); return; }
        s->cfg.volume = c.volume;
    }
    memcpy(s->pin_iq, &F(d, uint8_t, FMD_OFF_buf), buf_len);
    rc = fmb_process(s->h, s->pin_iq, (size_t) c.block_bytes, s->pin_pcm, s->pcm_cap + 8, &n_out);
    if (rc != FMB_OK) { fail("fmb_process", rc); return; }
    memcpy(&F(d, int16_t, FMD_OFF_result), s->pin_pcm, (size_t) n_out * sizeof(int16_t));
    F(d, int, FMD_OFF_result_len) = n_out;                 /* :603 / :787 */
    F(d, int, FMD_OFF_lp_len) = (int) (buf_len >> 3);      /* :410 */
    if (c.rate_out2 > 0 && c.mode >= 1 && c.size > 0) s->pos = (int) ((s->pos + (long long) (buf_len >> 4)) % c.size);
    s->conv = CONV_NONE;
    if (g_strict) {
        rc = export_state(s);
        if (rc != FMB_OK) fail("fm_dropin_export_state", rc);
    }
}

int fm_dropin_export_state(struct demod_state *d)
{
    struct shim *s = find(d, 0);
    return s ? export_state(s) : FMB_OK;
}

int fm_dropin_import_state(struct demod_state *d)
{
    struct shim *s = find(d, 0);
    if (!s || !s->h) return FMB_OK; /* nothing on the GPU yet: the struct is read at first use anyway */
    return import_state(s);
}

void fm_dropin_set_strict(int on) { g_strict = on; }
void fm_dropin_set_device(int device) { g_device = device; }
void fm_dropin_set_error_hook(void (*hook)(const char *)) { g_hook = hook; }
unsigned long fm_dropin_sizeof_demod_state(void) { return FMD_SIZEOF_DEMOD_STATE; }
