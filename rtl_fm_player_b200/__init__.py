"""rtl_fm_player_b200 -- B200-native batched FM demodulator (rtl_fm_player's IQ -> PCM path).

The product is the C-ABI shared library `libfmb.so` (include/fmb.h) built from
`csrc/` (hand-written sm_100a CUDA kernels + C host code); the Python modules
are thin ctypes mirrors of it used by the tests and the benchmark.
"""
from ._lib import (FMB_PRECISION_EXACT, FMB_PRECISION_FMA, FMB_REF_BLOCK_BYTES, FmbError, LIB_PATH, build, lib)
from .batch import DemodConfig, FmBatch, launch_count
from .multi import FmMulti, device_count, free_pinned, parse_device_list, pinned_array
from . import synth

__all__ = ["DemodConfig", "FmBatch", "FmMulti", "device_count", "parse_device_list", "pinned_array", "free_pinned", "FmbError", "build", "lib", "launch_count", "synth", "LIB_PATH",
           "FMB_PRECISION_EXACT", "FMB_PRECISION_FMA", "FMB_REF_BLOCK_BYTES"]
