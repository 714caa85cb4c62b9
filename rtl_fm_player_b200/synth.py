"""Synthetic rtl_sdr-format captures (include/fm_synth.h, csrc/fm_synth.c).

Loads only rtl_fm_player_b200/libfmsynth.so (plain C, no CUDA): generating input data never maps the
product library libfmb.so, so the reference arm of bench.py and the oracle tests stay free of it.
Importable on its own path too (bench.py's reference arm loads this FILE without importing the package).
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

SYNTH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfmsynth.so")
SYNTH_KINDS = {
    "fm_stereo": 0, "fm_mono": 1, "random": 2, "const0": 3, "const127": 4, "const128": 5,
    "const255": 6, "alt_0_255": 7, "impulse": 8, "carrier_off": 9,
}
_synth = None


def synth_lib() -> C.CDLL:
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise FileNotFoundError(f"{SYNTH_LIB_PATH} not built; run `make -C rtl_fm_player_b200/csrc`")
        lib = C.CDLL(SYNTH_LIB_PATH)
        lib.fmb_synth_capture.restype = C.c_int
        lib.fmb_synth_capture.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        _synth = lib
    return _synth


def capture(kind: str, stream: int, rate_in: int, offset_tuning: int, n_samples: int, first_sample: int = 0,
            out: np.ndarray | None = None) -> np.ndarray:
    """uint8 [2*n_samples] interleaved I,Q for one stream."""
    buf = np.empty(2 * n_samples, dtype=np.uint8) if out is None else out
    assert buf.dtype == np.uint8 and buf.size == 2 * n_samples and buf.flags.c_contiguous
    if synth_lib().fmb_synth_capture(SYNTH_KINDS[kind], stream, rate_in, offset_tuning, first_sample, n_samples,
                                     buf.ctypes.data) != 0:
        raise ValueError("fmb_synth_capture rejected its arguments")
    return buf


def batch(kind: str, n_streams: int, rate_in: int, offset_tuning: int, n_samples: int, unique: int | None = None,
          threads: int = 8) -> np.ndarray:
    """uint8 [n_streams, 2*n_samples].  With `unique` < n_streams only that many distinct
    captures are synthesised and the rest are copies (timing runs only)."""
    out = np.empty((n_streams, 2 * n_samples), dtype=np.uint8)
    u = n_streams if unique is None else min(unique, n_streams)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda s: capture(kind, s, rate_in, offset_tuning, n_samples, out=out[s]), range(u)))
    for s in range(u, n_streams):
        out[s] = out[s % u]
    return out
