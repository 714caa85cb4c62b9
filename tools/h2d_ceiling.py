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
    try:
        probe([dev], nbytes, 2, direction, wc)          # context + first-touch outside the timed part
        barrier.wait()
        t0 = time.perf_counter()
        a, b, s = probe([dev], nbytes, reps, direction, wc)
        q.put((dev, a, b, s, t0))
    except Exception as e:                               # noqa: BLE001
        q.put((dev, 0.0, 0.0, -1.0, str(e)))


def probe_processes(n, nbytes, reps, direction, wc):
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(d, nbytes, reps, direction, wc, barrier, q)) for d in range(n)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    if any(r[3] < 0 for r in res):
        raise RuntimeError(str(res))
    # the processes ran the same number of copies side by side: aggregate = sum of per-process rates
    return sum(r[1] for r in res), sum(r[2] for r in res), max(r[3] for r in res)


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
    print(f"{'N':>2} {'layout':>12} {'buffer':>6} {'H2D alone':>10} {'D2H alone':>10} {'mix H2D':>9} {'mix D2H':>8}")
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
                print(f"{n:>2} {layout:>12} {'wc' if wc else 'plain':>6} {h2d:>10.1f} {d2h:>10.1f} {mix[0]:>9.1f} {mix[1]:>8.1f}",
                      flush=True)


if __name__ == "__main__":
    sys.exit(main())
