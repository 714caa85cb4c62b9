"""Probe: the same 1024-stream step issued as G independent groups of streams on G CUDA streams, so that the
ragged end of one group's launch overlaps the other groups' work.  usage: python tools/two_group_probe.py [G ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtl_fm_player_b200 as R

S, BLOCK, NBUF, K, W = 1024, R.FMB_REF_BLOCK_BYTES, 4, 20, 4
uniq = 16
host = np.empty((NBUF, S, BLOCK), dtype=np.uint8)
for u in range(uniq):
    cap = R.synth.capture("fm_stereo", u, 192000, 0, NBUF * BLOCK // 2)
    for b in range(NBUF):
        host[b, u] = cap[b * BLOCK:(b + 1) * BLOCK]
for s in range(uniq, S):
    host[:, s] = host[:, s % uniq]
dev_in = [torch.from_numpy(host[b]).cuda() for b in range(NBUF)]
ref = None
for G in [int(x) for x in (sys.argv[1:] or ["1", "2", "4"])]:
    n = S // G
    fbs = [R.FmBatch(R.DemodConfig.stereo_192k(n_streams=n, device=0)) for _ in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    pitch = (fbs[0].next_out_count() + 7) & ~7
    pcm = torch.empty((S, pitch), dtype=torch.int16, device="cuda")
    def step(i):
        for g in range(G):
            fbs[g].process_device(dev_in[i % NBUF].data_ptr() + g * n * BLOCK, BLOCK, pcm.data_ptr() + g * n * pitch * 2, pitch, streams[g].cuda_stream)
    best = 1e9
    for rep in range(3):
        for i in range(W): step(i)
        for g in range(G): fbs[g].join(streams[g].cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for g in range(G): streams[g].wait_event(e0)
        for i in range(K): step(W + i)
        for g in range(G):
            fbs[g].join(streams[g].cuda_stream)
            torch.cuda.current_stream().wait_stream(streams[g])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    chk = int(pcm[:, :fbs[0].next_out_count()].to(torch.int64).sum().item())
    if ref is None: ref = chk
    print(f"groups={G}: {best:.4f} ms per 1024-stream step, checksum {'same' if chk == ref else 'DIFFERENT'}", flush=True)
    for f in fbs: f.close()
