"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
stereo 240 kHz (quirk path, ragged de-emphasis chunks), 5 streams x 3 blocks, plus mono (4:1 and generic tick path of the
warp-specialised kernel) and the drop-sample mode."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtl_fm_player_b200 as R
from oracle.oracle_py import PortOracle

B = 262144
for kw, kind in ((dict(rate_in=240000, rate_out2=48000, mode=2, size=90), "random"),
                 (dict(rate_in=192000, rate_out2=48000, mode=1, size=128), "fm_mono"),
                 (dict(rate_in=240000, rate_out2=48000, mode=1, size=128), "fm_mono"),   # warp-specialised kernel, generic tick path
                 (dict(rate_in=192000, rate_out2=48000, mode=0, size=90), "fm_stereo")):
    n, blocks = 5, 3
    iq = np.stack([R.synth.capture(kind, s, kw["rate_in"], 0, blocks * B // 2) for s in range(n)])
    with R.FmBatch(R.DemodConfig(n_streams=n, **kw)) as fb:
        pcm = fb.run(iq)
    for s in range(n):
        assert np.array_equal(pcm[s], PortOracle(**kw).run(iq[s])), (kw, s)
# the dynamic work assignment (a batch that fills the GPU): whole-stream runs chained through shared memory,
# fine-grain runs with lead-ins, the named-barrier hand-overs, on random bytes
os.environ["FMB_CHUNK"], os.environ["FMB_TAIL_PCT"] = "2", "30"
kw = dict(rate_in=192000, rate_out2=48000, mode=2, size=90)
n, uniq = 60, 3
iq = np.stack([R.synth.capture("random", s % uniq, 192000, 0, B // 2) for s in range(n)])
with R.FmBatch(R.DemodConfig(n_streams=n, **kw)) as fb:
    pcm = fb.run(iq)
want = [PortOracle(**kw).run(iq[s]) for s in range(uniq)]
for s in range(n):
    assert np.array_equal(pcm[s], want[s % uniq]), ("dynamic", s)
# round 2: overlapping launches (programmatic dependent launch: per-stream hand-over counters, whole-stream tickets)
# and the warp-specialised mono kernel (producer/consumer named barriers), device-resident back to back
import torch
del os.environ["FMB_CHUNK"], os.environ["FMB_TAIL_PCT"]
for kw, kind, n in ((dict(rate_in=192000, rate_out2=48000, mode=2, size=90), "fm_stereo", 40),
                    (dict(rate_in=192000, rate_out2=48000, mode=1, size=128), "fm_mono", 40),
                    (dict(rate_in=250000, rate_out2=44100, mode=1, size=90), "fm_mono", 40)):
    blocks, uniq = 3, 2
    os.environ["FMB_CHUNK"], os.environ["FMB_TAIL_PCT"] = "8", "0"      # force the ticketed path on a small batch
    iq = np.stack([R.synth.capture(kind, s % uniq, kw["rate_in"], 0, blocks * B // 2) for s in range(n)])
    d_in = torch.from_numpy(iq).cuda()
    with R.FmBatch(R.DemodConfig(n_streams=n, **kw)) as fb:
        n_out = fb.next_out_count()
        pitch = (n_out + 7) & ~7
        d_out = torch.zeros((blocks, n, pitch), dtype=torch.int16, device="cuda")
        st = torch.cuda.Stream()
        for b in range(blocks):
            fb.process_device(d_in.data_ptr() + b * B, iq.shape[1], d_out[b].data_ptr(), pitch, st.cuda_stream)
        fb.join(st.cuda_stream)
        torch.cuda.synchronize()
    got = np.concatenate([d_out[b, :, :n_out].cpu().numpy() for b in range(blocks)], axis=1)
    want = [PortOracle(**kw).run(iq[s]) for s in range(uniq)]
    for s in range(n):
        assert np.array_equal(got[s], want[s % uniq]), ("overlap", kw["mode"], s)
print("sanitize case ok")
