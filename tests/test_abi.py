"""CPU-only checks of the C ABI: the shared library loads without a GPU, exports every symbol
the headers declare, rejects bad arguments, fails loudly without a device, and the host-side
pieces (filter design, synthetic captures) are deterministic."""
import ctypes as C
import glob
import json
import os
import re

import numpy as np
import pytest

import rtl_fm_player_b200 as R
from rtl_fm_player_b200 import _lib as L
from vectors import B, CASES, CONFIGS, make_input, sha

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = np.load(os.path.join(HERE, "golden", "golden.npz"))
META = json.load(open(os.path.join(HERE, "golden", "golden_meta.json")))


SYNTH_HEADER = "fm_synth.h"        # its symbols live in libfmsynth.so (plain C), everything else in libfmb.so


def declared_symbols(synth: bool = False):
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        if (os.path.basename(h) == SYNTH_HEADER) != synth:
            continue
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src):
            n = m.group(1)
            if n.startswith(("fmb_", "filesrc_", "fm_wav_", "fm_timeshift_", "fm_dropin_")) or n in DROPIN:
                names.add(n)
    return sorted(names)


DROPIN = {"init_u8_f32_table", "init_lp_f32", "init_lp_real_f32", "deinit_lp_real_f32", "rotate_90_u8_f32", "u8_f32",
          "full_demod"}


def test_library_loads_and_exports_every_declared_symbol():
    lib = C.CDLL(R.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 50 and DROPIN <= set(syms)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    slib = C.CDLL(R.synth.SYNTH_LIB_PATH)
    assert declared_symbols(synth=True) == ["fmb_synth_capture"] and hasattr(slib, "fmb_synth_capture")
    assert not hasattr(lib, "fmb_synth_capture"), "the capture generator must stay out of the product library"


def test_synth_library_is_plain_c_and_generating_data_does_not_map_the_product_library():
    """bench.py's reference arm and the oracle legs only need input data: producing it must not load libfmb.so
    (a reference-arm process that maps the product library voids the ratio)."""
    import subprocess
    import sys
    needed = subprocess.run(["readelf", "-d", R.synth.SYNTH_LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda" not in needed and "libcudart" not in needed and "libfmb" not in needed
    code = ("import importlib.util, sys\n"
            f"spec = importlib.util.spec_from_file_location('fm_synth_only', {os.path.join(ROOT, 'rtl_fm_player_b200', 'synth.py')!r})\n"
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
            "iq = m.capture('fm_stereo', 3, 192000, 0, 4096)\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libfmsynth.so' in maps and 'libfmb.so' not in maps and 'libcudart' not in maps, maps\n"
            "print(int(iq.sum()))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert int(r.stdout) == int(R.synth.capture("fm_stereo", 3, 192000, 0, 4096).sum())


def test_python_binding_covers_fmb_h():
    declared = {s for s in declared_symbols() if s.startswith("fmb_")}
    assert declared <= set(L.SIGNATURES), sorted(declared - set(L.SIGNATURES))


def test_defaults_match_demod_init():
    c = L.FmbConfig()
    assert R.lib().fmb_default_config(C.byref(c)) == 0
    # demod_init, reference src/rtl_fm_player.c:1156-1195
    assert (c.rate_in, c.rate_out2, c.mode, c.size, c.offset_tuning) == (240000, 48000, 2, 90, 0)
    assert c.deemph == 0.000050 and abs(c.volume - 0.4) < 1e-7 and c.block_bytes == 262144
    R.lib().fmb_preset_stereo_192k(C.byref(c))
    assert (c.rate_in, c.mode, c.size) == (192000, 2, 90)       # -X :1464-1476
    R.lib().fmb_preset_mono_192k(C.byref(c))
    assert (c.rate_in, c.mode, c.size) == (192000, 1, 128)      # -Y :1477-1488


def test_create_rejects_bad_arguments_before_touching_cuda():
    lib = R.lib()
    h = C.c_void_p()
    for field, val in [("n_streams", 0), ("block_bytes", 1000), ("block_bytes", 0), ("mode", 3), ("rate_in", 0),
                       ("precision", 7)]:
        c = L.FmbConfig()
        lib.fmb_default_config(C.byref(c))
        setattr(c, field, val)
        assert lib.fmb_create(C.byref(c), C.byref(h)) == L.FMB_ERR_ARG, field
        assert not h.value
    c = L.FmbConfig()
    lib.fmb_default_config(C.byref(c))
    c.size = 64
    assert lib.fmb_create(C.byref(c), C.byref(h)) == L.FMB_ERR_UNSUPPORTED
    c.size = 90
    c.rate_in, c.rate_out2 = 64000, 48000     # stereo needs rate_in >= 2*rate_out2 (:593-597)
    assert lib.fmb_create(C.byref(c), C.byref(h)) == L.FMB_ERR_UNSUPPORTED
    assert lib.fmb_create(None, C.byref(h)) == L.FMB_ERR_ARG
    assert b"" != lib.fmb_last_error()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.FmbError) as e:
        R.FmBatch(R.DemodConfig.stereo_192k())
    assert e.value.code == L.FMB_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


class FmbTables(C.Structure):
    _fields_ = [("chan", C.c_float * 16), ("chan_s", C.c_float * 16), ("fm", C.c_float * 64), ("fp", C.c_float * 64),
                ("fs", C.c_float * 64), ("swf", C.c_float), ("cwf", C.c_float), ("lam", C.c_float), ("pcm_scale", C.c_float)]


@pytest.mark.parametrize("name", ["192k_90", "192k_128", "240k_90", "240k_128"])
def test_host_filter_design_is_bit_identical_to_the_reference(name):
    """fm_design.c vs the tables init_lp_f32 / init_lp_real_f32 produced in the reference build."""
    kw = META["tables"][name]
    c = L.FmbConfig()
    R.lib().fmb_default_config(C.byref(c))
    c.rate_in, c.size = kw["rate_in"], kw["size"]
    t = FmbTables()
    fn = C.CDLL(R.LIB_PATH).fmb_design_tables
    fn.argtypes = [C.POINTER(L.FmbConfig), C.POINTER(FmbTables)]
    assert fn(C.byref(c), C.byref(t)) == 0
    taps = kw["size"] // 2
    u = lambda a, n: np.frombuffer(a, dtype=np.uint32)[:n]
    g = lambda k: GOLD[f"tab_{name}_{k}"].view(np.uint32)
    assert np.array_equal(u(t.chan, 16), g("fb"))
    assert np.array_equal(u(t.fm, taps), g("fm"))
    assert np.array_equal(u(t.fp, taps), g("fp"))
    assert np.array_equal(u(t.fs, taps), g("fs"))
    misc = np.array([t.swf, t.cwf, t.lam], dtype=np.float32).view(np.uint32)
    assert np.array_equal(misc, g("misc"))
    # the pre-scaled copy is an exact power-of-two scaling
    assert np.array_equal(np.frombuffer(t.chan_s, np.float32)[:16] * np.float32(128.0), np.frombuffer(t.chan, np.float32)[:16])


def test_synth_is_deterministic_and_piecewise_consistent():
    a = R.synth.capture("fm_stereo", 5, 192000, 0, 4096)
    b = R.synth.capture("fm_stereo", 5, 192000, 0, 4096)
    assert np.array_equal(a, b)
    tail = R.synth.capture("fm_stereo", 5, 192000, 0, 1024, first_sample=3072)
    assert np.array_equal(a[2 * 3072:], tail)
    assert not np.array_equal(a, R.synth.capture("fm_stereo", 6, 192000, 0, 4096))
    r = R.synth.capture("random", 1, 192000, 0, 2048, first_sample=100)
    assert np.array_equal(r, R.synth.capture("random", 1, 192000, 0, 2148)[200:])
    assert set(np.unique(R.synth.capture("alt_0_255", 0, 192000, 0, 64))) == {0, 255}


def test_synth_reproduces_golden_input_bytes():
    for cid, cfg, kind, stream, blocks in CASES[:4]:
        assert sha(make_input(cfg, kind, stream, blocks)) == META["cases"][cid]["input_sha256"]


def test_device_list_parser_and_multi_create_without_a_device():
    """fmb_multi.h host logic that needs no GPU: "-d 0-3,6" syntax; creating shards without a CUDA device
    fails as a whole with FMB_ERR_CUDA (no CPU fallback) and the failing shard is named."""
    assert R.parse_device_list("0-3,6") == [0, 1, 2, 3, 6]
    assert R.parse_device_list("7") == [7]
    assert R.parse_device_list("0,0") == [0, 0]
    for bad in ("", "a", "1-", "3-1", "0,,1", "0,", "-1", "0-99999"):
        with pytest.raises(R.FmbError) as e:
            R.parse_device_list(bad)
        assert e.value.code == L.FMB_ERR_ARG, bad
    import torch
    if not torch.cuda.is_available():
        assert R.device_count() == 0
        with pytest.raises(R.FmbError) as e:
            R.FmMulti(R.DemodConfig.stereo_192k(n_streams=8), [0, 1])
        assert e.value.code == L.FMB_ERR_CUDA and "shard" in str(e.value)
    with pytest.raises(R.FmbError) as e:
        R.FmMulti(R.DemodConfig.stereo_192k(n_streams=1), [0, 1])
    assert e.value.code == L.FMB_ERR_ARG
