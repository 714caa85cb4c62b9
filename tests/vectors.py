"""Shared definitions of the parity cases.  Inputs are always regenerated from
the deterministic generator (fmb_synth_capture); tests/golden/golden.npz pins both the input
bytes (sha256) and the PCM the REFERENCE's own code produced for them."""
from __future__ import annotations

import hashlib

import numpy as np

B = 262144  # reference block, bytes

CONFIGS = {
    # name: oracle kwargs (= demod_state fields)
    "stereo192": dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=0),      # -X
    "stereo192_off": dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=1),  # -X -E offset
    "mono192": dict(rate_in=192000, rate_out2=48000, mode=1, size=128, offset_tuning=0),       # -Y
    "mono240_off": dict(rate_in=240000, rate_out2=48000, mode=1, size=128, offset_tuning=1),
    "stereo240": dict(rate_in=240000, rate_out2=48000, mode=2, size=90, offset_tuning=0),      # defaults (Q1 quirk)
    "stereo192_128": dict(rate_in=192000, rate_out2=48000, mode=2, size=128, offset_tuning=0),
    "mono192_90": dict(rate_in=192000, rate_out2=48000, mode=1, size=90, offset_tuning=0),
    "drop192": dict(rate_in=192000, rate_out2=48000, mode=0, size=90, offset_tuning=0),        # lpr.mode 0
    "nolpr192": dict(rate_in=192000, rate_out2=0, mode=2, size=90, offset_tuning=0),           # rate_out2 == 0
    "nodeemph192": dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=0, deemph=0.0),
    "loud192": dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=0, volume=1.5),
    # general -s / -r values: non-integer resampling ratios, other de-emphasis constants
    "stereo170_44": dict(rate_in=170000, rate_out2=44100, mode=2, size=90, offset_tuning=0),
    "stereo256_48": dict(rate_in=256000, rate_out2=48000, mode=2, size=90, offset_tuning=1),
    "mono240_32": dict(rate_in=240000, rate_out2=32000, mode=1, size=128, offset_tuning=0),
    "stereo192_us": dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=0, deemph=0.000075),
    "stereo96_48": dict(rate_in=96000, rate_out2=48000, mode=2, size=90, offset_tuning=0),    # ratio 2: the lowest stereo allows
    # the literal defaults of demod_init with -Y's decoder: mono at DEFAULT_SAMPLE_RATE (h:30), rotate path
    "mono240": dict(rate_in=240000, rate_out2=48000, mode=1, size=128, offset_tuning=0),
    "mono240_loud": dict(rate_in=240000, rate_out2=48000, mode=1, size=128, offset_tuning=0, volume=1.5),
    # -o 2 (main: rate_in *= post_downsample, :1510): filters designed for rate_in, resampler ticks at rate_out (:485)
    "stereo384_o2": dict(rate_in=384000, rate_out=192000, rate_out2=48000, mode=2, size=90, offset_tuning=0),
    "stereo480_o2": dict(rate_in=480000, rate_out=240000, rate_out2=48000, mode=2, size=90, offset_tuning=0),  # + Q1 quirk
    "mono384_o2": dict(rate_in=384000, rate_out=192000, rate_out2=48000, mode=1, size=128, offset_tuning=1),
}
# ratios between 2 and 3 whose block-start phases reach the reference's unemulated in-place overwrite: refused at create
REFUSED = {"stereo100_48": dict(rate_in=100000, rate_out2=48000, mode=2, size=90, offset_tuning=0)}

# (case id, config, synth kind, stream id, blocks)
CASES = [
    ("c2_stereo192_fm", "stereo192", "fm_stereo", 0, 3),
    ("stereo192_fm_s7", "stereo192", "fm_stereo", 7, 3),
    ("stereo192_random", "stereo192", "random", 1, 3),
    ("stereo192_const0", "stereo192", "const0", 0, 3),
    ("stereo192_const127", "stereo192", "const127", 0, 3),
    ("stereo192_const128", "stereo192", "const128", 0, 3),
    ("stereo192_const255", "stereo192", "const255", 0, 3),
    ("stereo192_alt", "stereo192", "alt_0_255", 0, 3),
    ("stereo192_impulse", "stereo192", "impulse", 0, 3),
    ("stereo192_carrier", "stereo192", "carrier_off", 0, 3),
    ("stereo192_off_fm", "stereo192_off", "fm_stereo", 2, 3),
    ("stereo192_off_random", "stereo192_off", "random", 3, 3),
    ("c1_mono192_fm", "mono192", "fm_mono", 0, 3),
    ("mono192_random", "mono192", "random", 4, 3),
    ("mono240_off_fm", "mono240_off", "fm_mono", 1, 4),
    ("stereo240_fm", "stereo240", "fm_stereo", 0, 7),
    ("stereo240_random", "stereo240", "random", 5, 7),
    ("stereo192_128_fm", "stereo192_128", "fm_stereo", 3, 3),
    ("mono192_90_random", "mono192_90", "random", 6, 3),
    ("drop192_fm", "drop192", "fm_stereo", 0, 3),
    ("nolpr192_random", "nolpr192", "random", 2, 2),
    ("nodeemph192_random", "nodeemph192", "random", 8, 3),
    ("loud192_random", "loud192", "random", 9, 3),
    ("stereo170_44_fm", "stereo170_44", "fm_stereo", 4, 5),
    ("stereo170_44_random", "stereo170_44", "random", 10, 5),
    ("stereo256_48_random", "stereo256_48", "random", 11, 5),
    ("mono240_32_fm", "mono240_32", "fm_mono", 2, 4),
    ("stereo192_us_fm", "stereo192_us", "fm_stereo", 5, 3),
    ("stereo96_48_random", "stereo96_48", "random", 12, 3),
    ("mono192_carrier", "mono192", "carrier_off", 0, 3),         # mono saturation (clamp at :722-729)
    ("mono240_fm", "mono240", "fm_mono", 3, 6),
    ("mono240_loud_random", "mono240_loud", "random", 1, 3),     # saturation on the generic tick path
    ("stereo384_o2_fm", "stereo384_o2", "fm_stereo", 6, 3),
    ("stereo480_o2_random", "stereo480_o2", "random", 13, 7),
    ("mono384_o2_fm", "mono384_o2", "fm_mono", 4, 3),
]
CASE_BY_ID = {c[0]: c for c in CASES}

# full-length cases of BASELINE.json configs[0]/[1]: 10 s captures; golden = sha256 of the PCM.  (id, config, kind,
# stream, full blocks): 10 s at 8*192 kHz = 30 720 000 B = 117 blocks + a tail that is never demodulated (:863-868);
# at the literal default rate 8*240 kHz = 38 400 000 B = 146 blocks + tail.
LONG_CASES = [
    ("c1_mono192_10s", "mono192", "fm_mono", 0, 117),
    ("c2_stereo192_10s", "stereo192", "fm_stereo", 0, 117),
    ("c0_mono240_10s", "mono240", "fm_mono", 0, 146),
]


def long_capture_bytes(cfg_name: str) -> int:
    """10 s of capture at 8 x rate_in, 2 bytes per IQ sample."""
    return 10 * 8 * CONFIGS[cfg_name]["rate_in"] * 2


def make_input(cfg_name: str, kind: str, stream: int, blocks: int) -> np.ndarray:
    from rtl_fm_player_b200 import synth
    kw = CONFIGS[cfg_name]
    return synth.capture(kind, stream, kw["rate_in"], kw.get("offset_tuning", 0), blocks * B // 2)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
