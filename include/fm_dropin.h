/*
 * fm_dropin.h -- the reference's own demod entry points, backed by the B200 library.
 *
 * These are the seven symbols rtl_fm_player's demod thread and main() use for the IQ -> PCM path
 * (reference src/rtl_fm_player.c):
 *
 *     init_u8_f32_table()      :195-204     called once from main (:1601)
 *     init_lp_f32()            :241-251     called once from main (:1602)
 *     init_lp_real_f32(d)      :413-453     called once from main (:1603); allocates d->lpr.{br,bm,bs,fm,fp,fs}
 *     deinit_lp_real_f32(d)    :455-470
 *     rotate_90_u8_f32(d)      :206-226     demod_thread_fn :879-886, when !d->offset_tuning
 *     u8_f32(d)                :228-239     demod_thread_fn :879-886, when d->offset_tuning
 *     full_demod(d)            :758-788     demod_thread_fn :889
 *
 * with the SAME names, signatures and `struct demod_state` (include/rtl_fm_player.h:127-175): a build of
 * rtl_fm_player that drops its own definitions of these functions and links libfmb.so runs its demod
 * thread unchanged.  The struct is the reference's; this library never defines it and
 * touches only the fields listed in rtl_fm_player_b200/csrc/ref_layout.h, by offset.
 *
 * Behaviour (what stays the same for the caller):
 *   - d->buf / d->buf_len in, d->result (int16 PCM) / d->result_len out, d->lp_len as lp_f32 leaves it
 *   - configuration is read from the struct: rate_in, rate_out2, lpr.mode, lpr.size, deemph, deemph_lambda,
 *     volume (re-read on every call), and which of rotate_90_u8_f32 / u8_f32 preceded the call
 *   - PCM is bit-identical to the reference's (tests/test_gpu_dropin.py)
 * What differs:
 *   - between calls the carried state (lowpass_tb, pre_r/j_f32, lpr rings, pp, prev_lpr_index,
 *     deemph_l/r_f32) lives in GPU memory.  It is read FROM the struct when the GPU context for `d` is
 *     created (first full_demod) and written back by fm_dropin_export_state(); with
 *     fm_dropin_set_strict(1) it is written back after every full_demod.
 *   - d->lowpassed is not filled (nothing in the reference reads it after full_demod; RMSShadowBuf is
 *     write-only)
 *   - supported decoder shapes: lpr.mode 0, or lpr.mode 1/2 with lpr.size 90 or 128 (the only values the
 *     reference's CLI can set: demod_init :1184, -X :1474, -Y :1487); kernels are compiled for exactly these, any
 *     other lpr.size in the struct is reported through the error hook (FMB_ERR_UNSUPPORTED), as is a stereo
 *     rate_out/rate_out2 ratio between 2 and 3 whose in-place output (:593-597) would overwrite unread input
 *     beyond the emulated first-sample case -- both at the FIRST full_demod, never in the middle of playback
 *   - rate_out is honoured separately from rate_in (they differ under -o N, main :1510): the filters follow
 *     rate_in (:419-429), the resampler ticks rate_out/rate_out2 (:485)
 *   - the functions are `void` like the reference's; a CUDA failure or an unsupported configuration calls
 *     the error hook (default: message on stderr, then abort()).  There is no CPU fallback.
 */
#ifndef FM_DROPIN_H
#define FM_DROPIN_H

#ifdef __cplusplus
extern "C" {
#endif

struct demod_state; /* the reference's, include/rtl_fm_player.h:127-175 */

void init_u8_f32_table(void);
void init_lp_f32(void);
void init_lp_real_f32(struct demod_state *fm);
void deinit_lp_real_f32(struct demod_state *fm);
void rotate_90_u8_f32(struct demod_state *d);
void u8_f32(struct demod_state *d);
void full_demod(struct demod_state *d);

/* ---- additions (not in the reference) ---- */
/* GPU state -> struct fields, in the reference's own representation (rings in ring order at lpr.pos).
 * Returns 0, or a negative FMB_ERR_* code. */
int fm_dropin_export_state(struct demod_state *d);
/* struct fields -> GPU state (e.g. after the caller edited or restored them). */
int fm_dropin_import_state(struct demod_state *d);
/* 1: export after every full_demod (the struct is always current; costs a device round trip). */
void fm_dropin_set_strict(int on);
/* CUDA device ordinal used for contexts created from now on (default 0). */
void fm_dropin_set_device(int device);
/* Releases the GPU context of `d` (exporting its state first).  deinit_lp_real_f32 does this too. */
void fm_dropin_release(struct demod_state *d);
/* Error hook: called with a message instead of the default stderr + abort(). */
void fm_dropin_set_error_hook(void (*hook)(const char *msg));
/* Size the library believes struct demod_state has (sizeof in the reference build it was generated
 * from); lets a host assert ABI agreement at start-up. */
unsigned long fm_dropin_sizeof_demod_state(void);

#ifdef __cplusplus
}
#endif
#endif /* FM_DROPIN_H */
