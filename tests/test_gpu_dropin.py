"""GPU parity of the drop-in entry points (include/fm_dropin.h) and of the offline batch player.

The struct handed to our full_demod() is the REFERENCE's own `struct demod_state`, allocated and
initialised by the reference's own code (oracle/_ref/libfmref.so: demod_init + init_lp_real_f32,
src/rtl_fm_player.c:1156, :413); a twin struct runs the reference's CPU functions on the same bytes.
Bar: int16 PCM and every carried state field bit-identical."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rtl_fm_player_b200 as R
from oracle.oracle_py import REF_CLI, REF_PATH
from vectors import B, CONFIGS, make_input

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class RefCfg(C.Structure):
    _fields_ = [("rate_in", C.c_int), ("rate_out2", C.c_int), ("mode", C.c_int), ("size", C.c_int),
                ("offset_tuning", C.c_int), ("deemph", C.c_double), ("volume", C.c_float), ("rate_out", C.c_int)]


def layout():
    out = {}
    for line in open(os.path.join(ROOT, "rtl_fm_player_b200", "csrc", "ref_layout.h")):
        f = line.split()
        if len(f) == 3 and f[0] == "#define" and f[2].isdigit():
            out[f[1]] = int(f[2])
    return out


LAY = layout()


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(REF_PATH):
        pytest.skip("oracle/_ref/libfmref.so missing")
    ref = C.CDLL(REF_PATH)
    ref.ref_create.restype = C.c_void_p
    ref.ref_create.argtypes = [C.POINTER(RefCfg)]
    ref.ref_destroy.argtypes = [C.c_void_p]
    ref.ref_demod_state.restype = C.c_void_p
    ref.ref_demod_state.argtypes = [C.c_void_p]
    ref.ref_block.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    ours = C.CDLL(R.LIB_PATH)
    for n in ("rotate_90_u8_f32", "u8_f32", "full_demod", "fm_dropin_release", "init_lp_real_f32", "deinit_lp_real_f32"):
        getattr(ours, n).argtypes = [C.c_void_p]
        getattr(ours, n).restype = None
    for n in ("fm_dropin_export_state", "fm_dropin_import_state"):
        getattr(ours, n).argtypes = [C.c_void_p]
    return ref, ours


def new_ref(ref, cfgname, **over):
    kw = dict(CONFIGS[cfgname]); kw.update(over)
    c = RefCfg(kw["rate_in"], kw["rate_out2"], kw["mode"], kw["size"], kw.get("offset_tuning", 0),
               kw.get("deemph", 0.000050), kw.get("volume", 0.4), kw.get("rate_out", 0))
    h = ref.ref_create(C.byref(c))
    assert h
    return h, ref.ref_demod_state(h)


def field(d, off, ctype):
    return ctype.from_address(d + off)


def state_bytes(d, mode):
    """Every carried field of the reference struct, in the reference's own representation."""
    size = field(d, LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_size"], C.c_int).value
    out = {
        "lowpass_tb": bytes((C.c_char * 192).from_address(d + LAY["FMD_OFF_lowpass_tb"])),
        "pre": bytes((C.c_char * 8).from_address(d + LAY["FMD_OFF_pre_r_f32"])),
        "deemph": bytes((C.c_char * 8).from_address(d + LAY["FMD_OFF_deemph_l_f32"])),
        "prev_lpr_index": field(d, LAY["FMD_OFF_prev_lpr_index"], C.c_int).value,
        "pos": field(d, LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_pos"], C.c_int).value,
        "pp": bytes((C.c_char * 4).from_address(d + LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_pp"])),
    }
    for ring in ("br",) + (("bm", "bs") if mode == 2 else ()):
        ptr = field(d, LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_" + ring], C.c_void_p).value
        out[ring] = bytes((C.c_char * (4 * size)).from_address(ptr))
    return out


def ours_block(ours, d, iq, offset_tuning):
    """What demod_thread_fn does with one chunk (:870-889), calling OUR symbols on the reference's struct."""
    C.memmove(d + LAY["FMD_OFF_buf"], iq.ctypes.data, iq.size)
    field(d, LAY["FMD_OFF_buf_len"], C.c_uint32).value = iq.size
    (ours.u8_f32 if offset_tuning else ours.rotate_90_u8_f32)(d)
    assert field(d, LAY["FMD_OFF_lp_len"], C.c_int).value == iq.size
    ours.full_demod(d)
    n = field(d, LAY["FMD_OFF_result_len"], C.c_int).value
    return np.frombuffer((C.c_char * (2 * n)).from_address(d + LAY["FMD_OFF_result"]), dtype=np.int16).copy()


def ref_block(ref, h, iq):
    pcm = np.empty(131072, np.int16)
    n = ref.ref_block(h, iq.ctypes.data, iq.size, pcm.ctypes.data)
    return pcm[:n].copy()


@pytest.mark.parametrize("cfgname,kind", [("stereo192", "fm_stereo"), ("stereo240", "random"), ("mono192", "fm_mono"),
                                          ("stereo192_off", "random"), ("mono240_off", "fm_mono"), ("drop192", "fm_stereo"),
                                          ("nolpr192", "random"), ("nodeemph192", "random"),
                                          # -o 2: the struct has rate_in = 2*rate_out (main :1510); lp_real_f32 ticks
                                          # at rate_out (:485) while the filters follow rate_in
                                          ("stereo384_o2", "fm_stereo"), ("stereo480_o2", "random")])
def test_full_demod_on_the_references_struct_is_bit_identical(libs, cfgname, kind):
    ref, ours = libs
    kw = CONFIGS[cfgname]
    off, mode = kw.get("offset_tuning", 0), kw["mode"]
    blocks = 7
    iq = make_input(cfgname, kind, 11, blocks)
    h_ref, d_ref = new_ref(ref, cfgname)          # runs the reference's CPU code
    h_gpu, d_gpu = new_ref(ref, cfgname)          # same struct type, demodulated by libfmb.so
    try:
        for b in range(4):
            blk = iq[b * B:(b + 1) * B]
            assert np.array_equal(ours_block(ours, d_gpu, blk, off), ref_block(ref, h_ref, blk)), f"block {b}"
        # GPU state -> struct: every carried field equals the reference's, in its own representation
        assert ours.fm_dropin_export_state(d_gpu) == 0
        assert state_bytes(d_gpu, mode) == state_bytes(d_ref, mode)
        # ... so the REFERENCE can carry on from the exported struct (GPU -> CPU hand-over)
        ours.fm_dropin_release(d_gpu)
        blk = iq[4 * B:5 * B]
        assert np.array_equal(ref_block(ref, h_gpu, blk), ref_block(ref, h_ref, blk))
        # ... and the GPU can carry on from a struct the reference advanced (CPU -> GPU hand-over)
        for b in (5, 6):
            blk = iq[b * B:(b + 1) * B]
            assert np.array_equal(ours_block(ours, d_gpu, blk, off), ref_block(ref, h_ref, blk)), f"block {b}"
        assert ours.fm_dropin_export_state(d_gpu) == 0
        assert state_bytes(d_gpu, mode) == state_bytes(d_ref, mode)
    finally:
        ours.fm_dropin_release(d_gpu)
        ref.ref_destroy(h_ref)
        ref.ref_destroy(h_gpu)


def test_init_lp_real_f32_again_restarts_the_gpu_history(libs):
    """The reference's init_lp_real_f32 callocs fresh rings (:430-436); called again on a struct that already has
    a GPU context, ours must forget the GPU copy of the history too."""
    ref, ours = libs
    iq = make_input("stereo192", "random", 5, 2)
    h_ref, d_ref = new_ref(ref, "stereo192")
    h_gpu, d_gpu = new_ref(ref, "stereo192")
    try:
        blk = iq[:B]
        assert np.array_equal(ours_block(ours, d_gpu, blk, 0), ref_block(ref, h_ref, blk))
        # both sides re-init WITHOUT deinit (leaks the old rings, as the reference would): rings zeroed, pos 0,
        # pp 0, everything else (lowpass_tb, pre_*, de-emphasis, resampler phase) carried on
        ours.init_lp_real_f32(d_gpu)
        ref2 = C.CDLL(REF_PATH)
        if not hasattr(ref2, "ref_init_lp_real"):
            pytest.skip("reference harness without the re-init hook")
        ref2.ref_init_lp_real.argtypes = [C.c_void_p]
        ref2.ref_init_lp_real(h_ref)
        blk = iq[B:2 * B]
        assert np.array_equal(ours_block(ours, d_gpu, blk, 0), ref_block(ref, h_ref, blk))
    finally:
        ours.fm_dropin_release(d_gpu)
        ref.ref_destroy(h_ref)
        ref.ref_destroy(h_gpu)


def test_volume_change_and_strict_mode(libs):
    ref, ours = libs
    iq = make_input("stereo192", "random", 3, 3)
    h_ref, d_ref = new_ref(ref, "stereo192")
    h_gpu, d_gpu = new_ref(ref, "stereo192")
    ours.fm_dropin_set_strict(1)
    try:
        for b, vol in enumerate((0.4, 1.0, 0.05)):                      # the player's volume keys change demod.volume
            for d in (d_ref, d_gpu):
                field(d, LAY["FMD_OFF_volume"], C.c_float).value = vol
            blk = iq[b * B:(b + 1) * B]
            assert np.array_equal(ours_block(ours, d_gpu, blk, 0), ref_block(ref, h_ref, blk))
            assert state_bytes(d_gpu, 2) == state_bytes(d_ref, 2)       # strict: struct current after every call
    finally:
        ours.fm_dropin_set_strict(0)
        ours.fm_dropin_release(d_gpu)
        ref.ref_destroy(h_ref)
        ref.ref_destroy(h_gpu)


def test_our_init_lp_real_f32_fills_the_struct_like_the_reference(libs):
    ref, ours = libs
    h_ref, d_ref = new_ref(ref, "stereo240")
    h_x, d_x = new_ref(ref, "stereo240")
    try:
        ours.deinit_lp_real_f32(d_x)                                     # frees the reference's arrays (:455-470)
        assert field(d_x, LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_rsize"], C.c_int).value == 0
        ours.init_lp_real_f32(d_x)                                       # ours: allocate + design
        for name in ("fm", "fp", "fs"):
            pa = field(d_ref, LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_" + name], C.c_void_p).value
            pb = field(d_x, LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_" + name], C.c_void_p).value
            assert bytes((C.c_char * 180).from_address(pa)) == bytes((C.c_char * 180).from_address(pb)), name
        lo, hi = LAY["FMD_OFF_lpr"] + LAY["FMD_LPR_OFF_swf"], LAY["FMD_OFF_lpr"] + LAY["FMD_SIZEOF_LP_REAL"]
        assert bytes((C.c_char * (hi - lo)).from_address(d_ref + lo)) == bytes((C.c_char * (hi - lo)).from_address(d_x + lo))
        iq = make_input("stereo240", "fm_stereo", 1, 2)
        for b in range(2):
            blk = iq[b * B:(b + 1) * B]
            assert np.array_equal(ours_block(ours, d_x, blk, 0), ref_block(ref, h_ref, blk))
    finally:
        ours.fm_dropin_release(d_x)
        ref.ref_destroy(h_ref)
        ref.ref_destroy(h_x)


@pytest.mark.parametrize("flag,cfgname,kind", [("-X", "stereo192", "fm_stereo"), ("-Y", "mono192", "fm_mono")])
def test_fmb_player_wav_files_equal_the_reference_pipeline(tmp_path, flag, cfgname, kind):
    """BASELINE configs[0]/[1] end to end in C: capture files -> fmb_player (file source, CUDA, WAV writer) vs
    the reference's own code (ref_offline PCM -> the reference's InitWaveOut/CloseWaveOut)."""
    player = os.path.join(ROOT, "rtl_fm_player_b200", "fmb_player")
    if not (os.path.exists(player) and os.path.exists(REF_CLI)):
        pytest.skip("fmb_player / ref_offline not built")
    ref = C.CDLL(REF_PATH)
    ref.ref_wav_write.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]
    lens = [9, 12, 5]                                                   # channels of different lengths (blocks) + a tail
    files = []
    for s, nb in enumerate(lens):
        p = tmp_path / f"ch{s}.u8"
        np.concatenate([make_input(cfgname, kind, s, nb), np.full(1000 * (s + 1), 127, np.uint8)]).tofile(p)
        files.append(str(p))
    r = subprocess.run([player, flag, "-w", *files], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for s, f in enumerate(files):
        pcm_path = tmp_path / f"ref{s}.pcm"
        subprocess.run([REF_CLI, flag, f, str(pcm_path)], check=True, capture_output=True)
        pcm = np.fromfile(pcm_path, dtype=np.uint8)
        want = tmp_path / f"ref{s}.wav"
        assert ref.ref_wav_write(str(want).encode(), CONFIGS[cfgname]["mode"], pcm.ctypes.data, pcm.size) == 0
        assert open(f + ".wav", "rb").read() == want.read_bytes(), f"channel {s}"
    # the same channels sharded over several devices by the C multi-GPU host (-d: every visible GPU, or three shards
    # sharing GPU 0 on a 1-GPU box): byte-identical WAV files again
    ndev = R.device_count()
    dlist = f"0-{ndev - 1}" if ndev >= 2 else "0,0,0"
    for f in files:
        os.rename(f + ".wav", f + ".single.wav")
    r = subprocess.run([player, flag, "-w", "-d", dlist, *files], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"on {min(max(ndev, 3) if ndev < 2 else ndev, len(files))} device(s)" in r.stderr, r.stderr
    for f in files:
        assert open(f + ".wav", "rb").read() == open(f + ".single.wav", "rb").read(), f
    # raw PCM output keeps every sample (no cluster rule)
    r = subprocess.run([player, flag, files[0]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    subprocess.run([REF_CLI, flag, files[0], str(tmp_path / "r.pcm")], check=True, capture_output=True)
    assert open(files[0] + ".pcm", "rb").read() == (tmp_path / "r.pcm").read_bytes()


@pytest.mark.parametrize("cfgname,kind,blocks", [("stereo192", "fm_stereo", 9), ("stereo240", "random", 13),
                                                 ("mono192", "fm_mono", 9), ("stereo192_off", "fm_stereo", 9)])
def test_unmodified_player_threads_with_libfmb_interposed_write_the_same_wav(tmp_path, cfgname, kind, blocks):
    """The strongest form of the drop-in claim: the reference's OWN dongle/demod/output threads (compiled
    unmodified into oracle/_ref/libfmref.so) run a capture file twice in separate processes -- once bound to
    the reference's CPU functions, once with rtl_fm_player_b200/libfmb.so first in the global symbol scope
    (what LD_PRELOAD does for the real binary), so that their calls to init_*/rotate_90_u8_f32/u8_f32/
    full_demod land in the CUDA drop-in.  The two WAV files must be byte-identical."""
    import sys
    c = CONFIGS[cfgname]
    iq = make_input(cfgname, kind, 5, blocks)
    cap = tmp_path / "cap.u8"
    iq.tofile(cap)
    outs = {}
    for how in ("ref", "dropin"):
        wav = tmp_path / f"{how}.wav"
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_ref_player.py"), how, str(cap), str(wav),
               str(c["rate_in"]), str(c["rate_out2"]), str(c["mode"]), str(c["size"]), str(c.get("offset_tuning", 0))]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and f"{blocks} chunks delivered" in r.stdout, r.stdout + r.stderr
        outs[how] = wav.read_bytes()
    assert len(outs["ref"]) > 260 + 32768
    assert outs["dropin"] == outs["ref"]
