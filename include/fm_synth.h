/*
 * fm_synth.h -- deterministic synthetic rtl_sdr-format captures, host only.
 *
 * Built into its OWN small library, rtl_fm_player_b200/libfmsynth.so (plain C + libm, no CUDA), so that
 * anything that only needs input data -- the tests, the oracle legs, bench.py's reference arm -- can
 * generate it without mapping the product library libfmb.so.  The reference ships neither captures nor a
 * generator; the format is what rtlsdr_read_async delivers (src/librtlsdr.c:1867) and `rtl_sdr` writes to
 * disk: interleaved unsigned 8-bit I,Q, no header.
 */
#ifndef FM_SYNTH_H
#define FM_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMB_SYNTH_FM_STEREO 0   /* two tones, 19 kHz pilot, 38 kHz DSB-SC L-R, +-75 kHz, noise */
#define FMB_SYNTH_FM_MONO 1     /* one tone, no pilot */
#define FMB_SYNTH_RANDOM 2      /* uniform random bytes: hits every atan2 branch */
#define FMB_SYNTH_CONST_0 3
#define FMB_SYNTH_CONST_127 4
#define FMB_SYNTH_CONST_128 5
#define FMB_SYNTH_CONST_255 6
#define FMB_SYNTH_ALT_0_255 7   /* I=0,Q=255: zero guards of the discriminator */
#define FMB_SYNTH_IMPULSE 8     /* one 255 in constant 127 */
#define FMB_SYNTH_CARRIER_OFF 9 /* carrier 90 kHz off centre: drives PCM into the clamp */
/* Writes 2*n_samples bytes of IQ for samples [first_sample, first_sample+n_samples) of `stream`.
 * Returns 0, or -1 on a bad argument. */
int fmb_synth_capture(int kind, int stream, int rate_in, int offset_tuning, uint64_t first_sample,
                      uint64_t n_samples, uint8_t *iq);

#ifdef __cplusplus
}
#endif
#endif /* FM_SYNTH_H */
