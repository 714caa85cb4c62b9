"""Synthetic rtl_sdr-format captures (fmb_synth_capture, csrc/fm_synth.c)."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib as L


def capture(kind: str, stream: int, rate_in: int, offset_tuning: int, n_samples: int, first_sample: int = 0,
            out: np.ndarray | None = None) -> np.ndarray:
    """uint8 [2*n_samples] interleaved I,Q for one stream."""
    buf = np.empty(2 * n_samples, dtype=np.uint8) if out is None else out
    assert buf.dtype == np.uint8 and buf.size == 2 * n_samples and buf.flags.c_contiguous
    L.check(L.lib().fmb_synth_capture(L.SYNTH_KINDS[kind], stream, rate_in, offset_tuning, first_sample, n_samples,
                                      buf.ctypes.data), "fmb_synth_capture")
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
