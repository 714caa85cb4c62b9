"""Device-resident timing of the demod step for several values of one environment knob, in one process.
usage: python tools/sweep_env.py FMB_STAGGER stereo 0 4000 8000 ...   (the library reads the knob in fmb_create)
       python tools/sweep_env.py FMB_CHUNK,FMB_TAIL_PCT stereo 0:100 2:100 1:20 ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtl_fm_player_b200 as R

knob, mode, vals = sys.argv[1], sys.argv[2], sys.argv[3:]
S, BLOCK, NBUF, K, W = int(os.environ.get("SWEEP_STREAMS", "1024")), R.FMB_REF_BLOCK_BYTES, 4, 20, 4
NBUF = max(1, min(NBUF, (6 << 30) // (S * BLOCK)))        # keep the inputs under 6 GiB
stereo = mode == "stereo"
uniq = 16
host = np.empty((NBUF, S, BLOCK), dtype=np.uint8)
for u in range(uniq):
    cap = R.synth.capture("fm_stereo" if stereo else "fm_mono", u, 192000, 0, NBUF * BLOCK // 2)
    for b in range(NBUF):
        host[b, u] = cap[b * BLOCK:(b + 1) * BLOCK]
for s in range(uniq, S):
    host[:, s] = host[:, s % uniq]
dev_in = [torch.from_numpy(host[b]).cuda() for b in range(NBUF)]
if os.environ.get("SWEEP_STREAM", "0") == "1":      # a created (non-default) stream instead of torch's current (legacy default) one
    _st = torch.cuda.Stream()
    torch.cuda.set_stream(_st)
stream = torch.cuda.current_stream().cuda_stream
ref = None
for v in vals:
    for kn, vv in zip(knob.split(","), v.split(":")):   # several knobs: names a,b  values 1:2
        os.environ[kn] = vv
    mk = R.DemodConfig.stereo_192k if stereo else R.DemodConfig.mono_192k
    fb = R.FmBatch(mk(n_streams=S, device=0))
    n_out = fb.next_out_count(); pitch = (n_out + 7) & ~7
    pcm = torch.empty((S, pitch), dtype=torch.int16, device="cuda")
    best = 1e9
    for rep in range(3):
        for i in range(W): fb.process_device(dev_in[i % NBUF].data_ptr(), BLOCK, pcm.data_ptr(), pitch, stream)
        fb.join(stream); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K): fb.process_device(dev_in[(W + i) % NBUF].data_ptr(), BLOCK, pcm.data_ptr(), pitch, stream)
        fb.join(stream); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    chk = int(pcm[:, :n_out].to(torch.int64).sum().item())
    if ref is None: ref = chk
    print(f"{knob}={v:>8s} {mode}: {best:.4f} ms/step  checksum {chk} {"same" if chk == ref else "DIFFERENT"}", flush=True)
    fb.close() if hasattr(fb, "close") else None
    del fb
