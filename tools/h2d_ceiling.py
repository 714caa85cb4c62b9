"""Raw host<->device copy ceiling of this box at 1/2/4/8 GPUs (tools/fm_copyprobe.cu -> tools/libfmprobe.so).

    python tools/h2d_ceiling.py [--mib 256] [--reps 20] > profiles/rNN_h2d_ceiling.txt

For every GPU count N <= visible devices: N concurrent streams of `reps` cudaMemcpyAsync of `mib` MiB from pinned
host buffers (plain and write-combined), H2D alone, D2H alone, and the demodulator's own mix (H2D + D2H/16),
  * "1 process"    : one process drives all N devices (the layout of the C multi-GPU host, fmb_multi)
  * "N processes"  : one process per device, started together (the torchrun layout of bench.py)
This is the number the end-to-end throughput (2 B in + 0.125 B out per IQ sample) is bounded by.
"""
import argparse
import ctypes as C
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def probe(devices, nbytes, reps, direction, wc):
    lib = C.CDLL(os.path.join(ROOT, "tools", "libfmprobe.so"))
    lib.fmprobe_copy.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p]
    dv = (C.c_int * len(devices))(*devices)
    a, b, s = C.c_double(), C.c_double(), C.c_double()
    err = C.create_string_buffer(256)
    if lib.fmprobe_copy(dv, len(devices), nbytes, reps, direction, wc, C.byref(a), C.byref(b), C.byref(s), err) != 0:
        raise RuntimeError(err.value.decode())
    return a.value, b.value, s.value


def _worker(dev, nbytes, reps, direction, wc, barrier, q):
    """One process per device.  The barrier is passed INTO the probe and taken after its buffers are allocated,
    touched and warmed up, right before the timed copies: all processes copy at the same time."""
    try:
        lib = C.CDLL(os.path.join(ROOT, "tools", "libfmprobe.so"))
        CB = C.CFUNCTYPE(None)
        lib.fmprobe_copy_sync.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p,
                                          CB, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        dv = (C.c_int * 1)(dev)
        a, b, s, t0, t1 = (C.c_double() for _ in range(5))
        err = C.create_string_buffer(256)
        cb = CB(lambda: barrier.wait())
        if lib.fmprobe_copy_sync(dv, 1, nbytes, reps, direction, wc, C.byref(a), C.byref(b), C.byref(s), err, cb,
                                 C.byref(t0), C.byref(t1)) != 0:
            raise RuntimeError(err.value.decode())
        q.put((dev, t0.value, t1.value, None))
    except Exception as e:                               # noqa: BLE001
        try:
            barrier.abort()
        except Exception:                                # noqa: BLE001
            pass
        q.put((dev, 0.0, 0.0, str(e)))


def probe_processes(n, nbytes, reps, direction, wc):
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(d, nbytes, reps, direction, wc, barrier, q)) for d in range(n)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    if any(r[3] for r in res):
        raise RuntimeError(str(res))
    # aggregate over the UNION of the processes' timed intervals (they start together behind the barrier)
    span = max(r[2] for r in res) - min(r[1] for r in res)
    overlap = min(r[2] for r in res) - max(r[1] for r in res)
    in_b = nbytes if direction != 1 else 0
    out_b = (nbytes // 16 if direction == 2 else nbytes) if direction != 0 else 0
    return in_b * reps * n / span * 1e-9, out_b * reps * n / span * 1e-9, overlap / span


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--gpus", default="1,2,4,8")
    args = ap.parse_args()
    import torch
    ndev = torch.cuda.device_count()
    nbytes = args.mib << 20
    print(f"# raw copy ceiling: {args.reps} x {args.mib} MiB per device per direction, pinned host memory; "
          f"{ndev} visible GPUs, {os.cpu_count()} host cores; GB/s aggregate over the N devices")
    print("# N processes: every process allocates and warms up its buffers, then all pass a barrier and copy together; rates are "
          "over the union of their timed intervals (last column: fraction of that span during which ALL were copying)")
    print(f"{'N':>2} {'layout':>12} {'buffer':>6} {'H2D alone':>10} {'D2H alone':>10} {'mix H2D':>9} {'mix D2H':>8} {'overlap':>8}")
    for n in [int(x) for x in args.gpus.split(",")]:
        if n > ndev:
            continue
        for layout in ("1 process", f"{n} processes"):
            if n == 1 and layout != "1 process":
                continue
            for wc in (0, 1):
                f = (lambda d, w: probe(list(range(n)), nbytes, args.reps, d, w)) if layout == "1 process" else \
                    (lambda d, w: probe_processes(n, nbytes, args.reps, d, w))
                h2d = f(0, wc)[0]
                d2h = f(1, wc)[1]
                mix = f(2, wc)
                ov = f"{mix[2]:>8.2f}" if layout != "1 process" else f"{'-':>8}"
                print(f"{n:>2} {layout:>12} {'wc' if wc else 'plain':>6} {h2d:>10.1f} {d2h:>10.1f} {mix[0]:>9.1f} {mix[1]:>8.1f} {ov}",
                      flush=True)


if __name__ == "__main__":
    sys.exit(main())
