"""Device-resident step time of any configuration (1024 streams by default).
usage: python tools/time_config.py rate_in rate_out2 mode size offset_tuning [streams]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtl_fm_player_b200 as R

rate_in, rate_out2, mode, size, off = [int(x) for x in sys.argv[1:6]]
S = int(sys.argv[6]) if len(sys.argv) > 6 else 1024
BLOCK, NBUF, K, W = R.FMB_REF_BLOCK_BYTES, 4, 20, 4
kind = "fm_stereo" if mode == 2 else "fm_mono"
uniq = 16
host = np.empty((NBUF, S, BLOCK), dtype=np.uint8)
for u in range(uniq):
    cap = R.synth.capture(kind, u, rate_in, off, NBUF * BLOCK // 2)
    for b in range(NBUF):
        host[b, u] = cap[b * BLOCK:(b + 1) * BLOCK]
for s in range(uniq, S):
    host[:, s] = host[:, s % uniq]
dev_in = [torch.from_numpy(host[b]).cuda() for b in range(NBUF)]
stream = torch.cuda.current_stream().cuda_stream
fb = R.FmBatch(R.DemodConfig(n_streams=S, rate_in=rate_in, rate_out2=rate_out2, mode=mode, size=size, offset_tuning=off))
pitch = (fb.max_out_count() + 7) & ~7 if hasattr(fb, "max_out_count") else 8200
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
print(f"rate_in={rate_in} rate_out2={rate_out2} mode={mode} size={size} offset={off} streams={S}: {best:.4f} ms/step = {S*131072/best*1e-6:.1f} G samples/s", flush=True)
