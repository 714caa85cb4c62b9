#!/usr/bin/env python
"""bench.py -- throughput of the batched IQ -> PCM path (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libfmb.so)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU code, all host cores

A step = one pass of the hot path over one batch: every stream of the rank's shard consumes one
reference block (262144 bytes = 131072 IQ samples, rtl_fm_player.h:31-33) and produces its PCM.
Workload = BASELINE.json configs[3]/[4]: 1024 stereo FM streams per GPU (8192 over 8 GPUs),
-X preset (192 kHz -> 48 kHz stereo, 90 taps), rotate path.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 262144                    # bytes per stream per step (MAXIMUM_BUF_LENGTH)
SAMPLES_PER_BLOCK = BLOCK // 2
ALG_BYTES = {"stereo": 2.125, "mono": 2.0625}   # u8 IQ in + int16 PCM out per IQ sample
FP32_OPS_PER_SAMPLE = {"stereo": 69.0, "mono": 23.0}  # bit-exact (no FMA) FP32 instructions per IQ sample, SURVEY s8(d)
METRIC = "aggregate IQ Msamples/s to stereo PCM"
UNIT = "Msamples/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(streams, mode, precision):
    """DRAM bytes per launch of the demod kernel from the committed ncu capture (profiles/demod_traffic.json);
    only valid for the workload it was captured on."""
    try:
        with open(os.path.join(ROOT, "profiles", "demod_traffic.json")) as f:
            t = json.load(f)
        if streams == 1024 and mode == "stereo" and precision == "exact":
            return t["traffic"], t["source"]
    except Exception:
        pass
    return None, None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled through NVML every ~2 ms
    (nvidia-smi, one process per sample, is too slow for a region of a few milliseconds; it is the
    fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.bits = [(pynvml.nvmlClocksEventReasonHwSlowdown if hasattr(pynvml, "nvmlClocksEventReasonHwSlowdown") else 0x8),
                         0x40, 0x20, 0x4]  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        except Exception:
            self.nvml = None

    def _run(self):
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
                    try:
                        mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        mask = int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.samples.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & b else "Not Active" for b in self.bits])
                    self._stop.wait(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = sorted({n for s in self.samples for n, v in zip(self.NAMES, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own rtl_fm_player.c (oracle/_ref/ref_offline)
# ------------------------------------------------------------------------------------------
def cpu_reference_run(mode: str, target_seconds: float, cores: int | None = None):
    """One reference process per core, each demodulating its own preloaded capture
    (BASELINE.md s3).  Returns dict(value Msamples/s, cores, kind, sample, seconds)."""
    from rtl_fm_player_b200 import synth
    cli = os.path.join(ROOT, "oracle", "_ref", "ref_offline")
    kind = "reference"
    if not os.path.exists(cli):
        return cpu_port_run(mode, target_seconds, cores)
    cores = cores or os.cpu_count() or 1
    flag = "-X" if mode == "stereo" else "-Y"
    nblk = 4
    tmp = tempfile.mkdtemp(prefix="fmref_")
    files = []
    for c in range(min(cores, 8)):  # 8 distinct captures, reused round-robin
        fn = os.path.join(tmp, f"s{c}.u8")
        synth.capture("fm_stereo" if mode == "stereo" else "fm_mono", c, 192000, 0, nblk * SAMPLES_PER_BLOCK).tofile(fn)
        files.append(fn)
    # calibrate one process
    t0 = time.perf_counter()
    subprocess.run([cli, flag, "-n", "8", files[0]], capture_output=True, text=True, check=True)
    per_rep = (time.perf_counter() - t0) / 8
    reps = max(4, int(target_seconds / max(per_rep, 1e-4)))
    t0 = time.perf_counter()
    procs = [subprocess.Popen([cli, flag, "-n", str(reps), files[c % len(files)]], stdout=subprocess.PIPE, text=True)
             for c in range(cores)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    samples, dsp_max = 0, 0.0
    for o in outs:
        for tok in o.split():
            if tok.startswith("samples="):
                samples += int(tok[8:])
            if tok.startswith("dsp_s="):
                dsp_max = max(dsp_max, float(tok[6:]))
    for fn in files:
        os.unlink(fn)
    os.rmdir(tmp)
    return {"value": samples / wall * 1e-6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{cores} reference processes (oracle/_ref/ref_offline {flag}), each {reps}x a {nblk}-block synthetic capture "
                      f"({reps * nblk} blocks of 131072 IQ samples), whole-process wall {wall:.1f}s, slowest DSP-only {dsp_max:.1f}s",
            "seconds": wall, "dsp_only_value": samples / max(dsp_max, 1e-9) * 1e-6}


def cpu_port_run(mode: str, target_seconds: float, cores: int | None = None):
    """Fallback when oracle/_ref was not built: the C restatement, one thread per core."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle_py import PortOracle
    from rtl_fm_player_b200 import synth
    cores = cores or os.cpu_count() or 1
    kw = dict(rate_in=192000, rate_out2=48000, mode=2 if mode == "stereo" else 1, size=90 if mode == "stereo" else 128)
    iq = synth.capture("fm_stereo" if mode == "stereo" else "fm_mono", 0, 192000, 0, 4 * SAMPLES_PER_BLOCK)
    o = PortOracle(**kw)
    t0 = time.perf_counter(); o.bench(iq, 2); per_rep = (time.perf_counter() - t0) / 2
    reps = max(2, int(target_seconds / per_rep))
    orcs = [PortOracle(**kw) for _ in range(cores)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:   # ctypes releases the GIL
        samples = sum(ex.map(lambda oo: oo.bench(iq, reps), orcs))
    wall = time.perf_counter() - t0
    return {"value": samples / wall * 1e-6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} threads of oracle/fm_oracle.c, each {reps}x a 4-block capture, wall {wall:.1f}s", "seconds": wall}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_step = max(1.0, min(8.0, 60.0 / (steps + warm)))
    for _ in range(warm):
        cpu_reference_run(args.mode, per_step)
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(steps):
        last = cpu_reference_run(args.mode, per_step)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": wall / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 0),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, segs):
    return {"workload": f"{args.streams} {args.mode} FM streams per GPU x 1 reference block (262144 B = 131072 IQ samples) per step; "
                        f"{'-X' if args.mode == 'stereo' else '-Y'} preset (192 kHz -> 48 kHz, "
                        f"{'90-tap stereo' if args.mode == 'stereo' else '128-tap mono'}), rotate_90 path; BASELINE.json configs[3]/[4]",
            "streams_per_gpu": args.streams, "block_bytes": BLOCK, "precision": args.precision, "segments": segs,
            "l2": f"{args.nbuf} distinct {args.streams * BLOCK >> 20} MiB input batches in rotation (each > 126 MB L2), no flush needed"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rtl_fm_player_b200 as R

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = R.lib().fmb_bind_thread_to_device_node(local)   # pinned buffers below become node-local
    dist = None
    if world > 1:
        import torch.distributed as dist
        backend = os.environ.get("FMB_BENCH_BACKEND", "nccl")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    red_dev = "cuda" if dist is None or dist.get_backend() == "nccl" else "cpu"

    S, K, W = args.streams, max(1, args.steps), max(3, args.warmup)
    stereo = args.mode == "stereo"
    mk = R.DemodConfig.stereo_192k if stereo else R.DemodConfig.mono_192k
    cfg = mk(n_streams=S, device=local, precision=R.FMB_PRECISION_FMA if args.precision == "fma" else R.FMB_PRECISION_EXACT,
             segments=args.segments)
    fb = R.FmBatch(cfg)
    n_out = fb.next_out_count()          # constant for the 192k presets
    pitch = (n_out + 7) & ~7

    # synthetic input: `unique` distinct channels per rank (global stream ids), replicated to S;
    # nbuf consecutive blocks of each so that successive steps read different HBM
    uniq = min(args.unique, S)
    base = int(os.environ.get("FMB_BENCH_DATA_RANK", rank)) * S     # (override: time another rank's channels on one GPU)
    kind = "fm_stereo" if stereo else "fm_mono"
    host = np.empty((args.nbuf, S, BLOCK), dtype=np.uint8)
    from concurrent.futures import ThreadPoolExecutor
    def gen(u):
        cap = R.synth.capture(kind, base + u, 192000, 0, args.nbuf * SAMPLES_PER_BLOCK)
        for b in range(args.nbuf):
            host[b, u] = cap[b * BLOCK:(b + 1) * BLOCK]
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(gen, range(uniq)))
    for s in range(uniq, S):
        host[:, s] = host[:, s % uniq]
    dev_in = [torch.from_numpy(host[b]).cuda() for b in range(args.nbuf)]
    dev_pcm = torch.empty((S, pitch), dtype=torch.int16, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        fb.process_device(dev_in[i % args.nbuf].data_ptr(), BLOCK, dev_pcm.data_ptr(), pitch, stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: `value` ----
    for i in range(W):
        step(i)
    fb.join(stream)
    barrier()
    fb.profile_enable(True)
    fb.profile_reset()
    launches0 = R.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        t_host = time.perf_counter()
        for i in range(K):
            step(W + i)
        t_host = time.perf_counter() - t_host       # host time to enqueue the K steps (must stay below the device time)
        fb.join(stream)
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = R.launch_count() - launches0
    prof = fb.profile_read()
    fb.profile_enable(False)
    if dist is not None:
        t = torch.tensor([ms_total], device=red_dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    samples_step_all = S * SAMPLES_PER_BLOCK * world
    value = samples_step_all * K / (ms_total * 1e-3) * 1e-6

    # ---- the same steps in FMB_PRECISION_FMA (FIR multiply-adds fused: within +-1 LSB of int16 PCM, the
    # tolerance north_star states for the floating-point stages); reported beside the bit-exact headline ----
    fma_alt = None
    if args.precision == "exact" and not args.no_fma_alt:
        fb2 = R.FmBatch(mk(n_streams=S, device=local, precision=R.FMB_PRECISION_FMA, segments=args.segments))
        for i in range(W):
            fb2.process_device(dev_in[i % args.nbuf].data_ptr(), BLOCK, dev_pcm.data_ptr(), pitch, stream)
        fb2.join(stream)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            fb2.process_device(dev_in[(W + i) % args.nbuf].data_ptr(), BLOCK, dev_pcm.data_ptr(), pitch, stream)
        fb2.join(stream)
        f1.record()
        barrier()
        fma_ms = f0.elapsed_time(f1)
        if dist is not None:
            t = torch.tensor([fma_ms], device=red_dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fma_ms = float(t.item())
        fma_alt = {"value": samples_step_all * K / (fma_ms * 1e-3) * 1e-6, "unit": UNIT, "ms_per_step": fma_ms / K,
                   "tolerance": "+-1 LSB int16 PCM vs the reference (tests/test_gpu_parity.py::test_fma_precision_within_one_lsb)"}
        fb2.close()

    # ---- end to end through the host API: pinned host -> H2D -> kernels -> D2H pinned ----
    nh = 2
    wc = os.environ.get("FMB_BENCH_WC", "1") == "1"          # upload buffers: pinned + write-combined (fmb_host_alloc_wc)
    if wc:
        import ctypes as C
        h_in_np, h_in_ptr = [], []
        for b in range(nh):
            ptr = C.c_void_p()
            assert R.lib().fmb_host_alloc_wc(C.byref(ptr), S * BLOCK) == 0
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (S, BLOCK))
            a[:] = host[b % args.nbuf]
            h_in_np.append(a); h_in_ptr.append(ptr.value)
    else:
        h_in = [torch.empty((S, BLOCK), dtype=torch.uint8, pin_memory=True) for _ in range(nh)]
        for b in range(nh):
            h_in[b].numpy()[:] = host[b % args.nbuf]
        h_in_ptr = [t.data_ptr() for t in h_in]
    h_pcm = [torch.empty((S, pitch), dtype=torch.int16, pin_memory=True) for _ in range(R._lib.FMB_PIPE_DEPTH)]
    def e2e_loop(n):
        tickets = []
        for i in range(n):
            if len(tickets) >= R._lib.FMB_PIPE_DEPTH - 1:
                fb.wait(tickets.pop(0))
            tickets.append(fb.submit(h_in_ptr[i % nh], BLOCK, h_pcm[i % len(h_pcm)].data_ptr(), pitch))
        for t in tickets:
            fb.wait(t)
    torch.cuda.synchronize()
    e2e_loop(max(3, W))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=red_dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = samples_step_all * K / e2e_s * 1e-6
    checksum = int(h_pcm[(K - 1) % len(h_pcm)].numpy()[:, :n_out].astype(np.int64).sum())

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    demod_ms = prof["demod_ms"] / max(prof["demod_launches"], 1)
    deemph_ms = prof["deemph_ms"] / max(prof["deemph_launches"], 1)
    alg = ALG_BYTES[args.mode] * S * SAMPLES_PER_BLOCK           # algorithmic bytes per launch (one rank)
    achieved = alg / (demod_ms * 1e-3) * 1e-9
    traffic, traffic_src = ncu_traffic(S, args.mode, args.precision)
    clocks = clk.summary()
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6                         # FP32 lane-instructions/s at the clock seen under load
    fp32_ach = FP32_OPS_PER_SAMPLE[args.mode] * S * SAMPLES_PER_BLOCK / (demod_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, fb.cfg.segments),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": S * BLOCK * world,
                "d2h_bytes_per_step": S * n_out * 2 * world, "ms_per_step": e2e_s / K * 1e3,
                "api": "fmb_submit/fmb_wait, pinned host buffers (IQ: write-combined), 2 steps in flight", "pcm_checksum": checksum,
                "host_numa_node": numa_node},
        "gpu_launches": launches, "host_enqueue_ms_per_step": t_host / K * 1e3,
        "roofline": {"bound": "hbm", "kernel": "fmb_demod_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "alg_bytes_per_launch": alg, "peak_source": peak_src,
                     "alg_bytes_per_iq_sample": ALG_BYTES[args.mode], "kernel_ms": demod_ms,
                     "deemph_kernel_ms": deemph_ms,
                     "fp32_pipe": {"ops_per_iq_sample": FP32_OPS_PER_SAMPLE[args.mode], "achieved_Tops": fp32_ach * 1e-12,
                                   "peak_Tops": fp32_peak * 1e-12, "frac": fp32_ach / fp32_peak,
                                   "note": "the path is FP32-issue bound (SURVEY s8d); peak = 148 SM x 128 lanes x SM clock under load"}},
    }
    if fma_alt is not None:
        line["precision_fma"] = fma_alt
    if world == 1 and not args.no_cpu:
        cb = cpu_reference_run(args.mode, args.cpu_seconds)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, ...) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                     # C-level and Python-level stdout -> stderr from here on
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU")
    ap.add_argument("--mode", default="stereo", choices=["stereo", "mono"])
    ap.add_argument("--precision", default="exact", choices=["exact", "fma"])
    ap.add_argument("--segments", type=int, default=0)
    ap.add_argument("--unique", type=int, default=64, help="distinct synthetic channels per GPU (rest are copies)")
    ap.add_argument("--nbuf", type=int, default=4, help="distinct input batches rotated through (each > L2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fma-alt", action="store_true", help="skip the extra FMB_PRECISION_FMA timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
