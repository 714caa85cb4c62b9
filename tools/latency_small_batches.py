import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import rtl_fm_player_b200 as R
B = R.FMB_REF_BLOCK_BYTES
for n in (1, 8, 64):
    cfg = R.DemodConfig.stereo_192k(n_streams=n, device=0)
    iq = np.stack([R.synth.capture("fm_stereo", s, 192000, 0, 8 * B // 2) for s in range(min(n, 4))])
    iq = np.concatenate([iq] * ((n + 3) // 4))[:n]
    with R.FmBatch(cfg) as fb:
        for b in range(3): fb.process(iq[:, b * B:(b + 1) * B])
        torch.cuda.synchronize()
        t = []
        for rep in range(40):
            b = 3 + rep % 5
            t0 = time.perf_counter(); fb.process(iq[:, b * B:(b + 1) * B]); t.append(time.perf_counter() - t0)
        t = np.array(t) * 1e6
        print(f"fmb_process, {n} stream(s), one 262144-byte block each, pageable numpy buffers: median {np.median(t):.0f} us, min {t.min():.0f} us per call "
              f"(real time for the block: {131072/1.536e6*1e6:.0f} us; reference CPU ~2100 us per stream)")
