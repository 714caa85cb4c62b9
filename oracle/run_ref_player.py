"""TEST INFRASTRUCTURE.  Runs the reference's own threaded pipeline (oracle/_ref/libfmref.so:
ref_player_run = dongle_thread_fn + demod_thread_fn + output_thread_fn of src/rtl_fm_player.c) on a
capture file, in THIS process, and writes the WAV the reference writes.

    python oracle/run_ref_player.py ref    capture.u8 out.wav rate_in rate_out2 mode size offset
    python oracle/run_ref_player.py dropin capture.u8 out.wav ...

`ref`    : the threads call the reference's CPU functions.
`dropin` : rtl_fm_player_b200/libfmb.so is loaded RTLD_GLOBAL first, so the unmodified threads' calls to
           init_u8_f32_table / init_lp_f32 / init_lp_real_f32 / rotate_90_u8_f32 / u8_f32 / full_demod
           bind to the CUDA drop-in (include/fm_dropin.h) -- what LD_PRELOAD=libfmb.so does for the player.
Both use the product's capture-file source (include/fm_filesrc.h) behind rtlsdr_read_async.
A separate process per run keeps the symbol scopes apart (the test suite loads both libraries)."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FMB = os.path.join(ROOT, "rtl_fm_player_b200", "libfmb.so")
REF = os.path.join(HERE, "_ref", "libfmref.so")


class RefCfg(C.Structure):
    _fields_ = [("rate_in", C.c_int), ("rate_out2", C.c_int), ("mode", C.c_int), ("size", C.c_int),
                ("offset_tuning", C.c_int), ("deemph", C.c_double), ("volume", C.c_float)]


def main():
    how, cap, wav = sys.argv[1:4]
    rate_in, rate_out2, mode, size, offset = [int(x) for x in sys.argv[4:9]]
    fmb = C.CDLL(FMB, mode=C.RTLD_GLOBAL if how == "dropin" else C.RTLD_LOCAL)
    ref = C.CDLL(REF, mode=C.RTLD_LOCAL)
    if how == "dropin":                      # prove which definition the player's PLT will reach
        ours = C.cast(fmb.full_demod, C.c_void_p).value
        assert C.cast(C.CDLL(None).full_demod, C.c_void_p).value == ours, "libfmb.so is not first in the global scope"
    cfg = RefCfg()
    ref.ref_default_cfg(C.byref(cfg))
    cfg.rate_in, cfg.rate_out2, cfg.mode, cfg.size, cfg.offset_tuning = rate_in, rate_out2, mode, size, offset

    src = C.c_void_p()
    fmb.filesrc_open.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
    assert fmb.filesrc_open(C.byref(src), cap.encode()) == 0
    fill_max = C.c_uint32()
    ref.ref_player_input_fill.restype = C.c_void_p
    fill = ref.ref_player_input_fill(C.byref(fill_max))
    fmb.filesrc_set_backpressure.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    assert fmb.filesrc_set_backpressure(src, fill, fill_max.value) == 0
    ref.ref_player_run.argtypes = [C.POINTER(RefCfg), C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = ref.ref_player_run(C.byref(cfg), wav.encode(), C.cast(fmb.filesrc_read_async, C.c_void_p),
                            C.cast(fmb.filesrc_cancel_async, C.c_void_p), src)
    fmb.filesrc_chunks_delivered.restype = C.c_uint64
    fmb.filesrc_chunks_delivered.argtypes = [C.c_void_p]
    print(f"{how}: rc {rc}, {fmb.filesrc_chunks_delivered(src)} chunks delivered")
    fmb.filesrc_close.argtypes = [C.c_void_p]
    fmb.filesrc_close(src)
    return rc


if __name__ == "__main__":
    sys.exit(main())
