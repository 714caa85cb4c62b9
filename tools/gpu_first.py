"""Development probe run on the GPU box: stage-by-stage parity vs the port oracle + rough timing."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtl_fm_player_b200 as R
from oracle.oracle_py import PortOracle

B = 262144

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}"); return False
    if a.dtype == np.float32:
        same = np.array_equal(a.view(np.uint32), b.view(np.uint32))
        if same: print(f"  {name}: bit-exact ({a.size})"); return True
        eq = (a == b)
        bad = np.flatnonzero(~eq)
        if bad.size == 0:
            print(f"  {name}: equal up to zero sign"); return True
        print(f"  {name}: {bad.size}/{a.size} differ, first at {bad[:8]}, max abs {np.nanmax(np.abs(a-b))}")
        return False
    same = np.array_equal(a, b)
    if same: print(f"  {name}: exact ({a.size})")
    else:
        bad = np.flatnonzero(a != b)
        print(f"  {name}: {bad.size}/{a.size} differ, first at {bad[:8]}, max abs {np.abs(a.astype(int)-b.astype(int)).max()}")
    return same

def check(label, kw, kind='fm_stereo', n_streams=3, nblocks=3, segments=0, precision=0):
    print(f"== {label} kind={kind} streams={n_streams} blocks={nblocks} segs={segments} prec={precision}")
    cfg = R.DemodConfig(n_streams=n_streams, segments=segments, precision=precision, **kw)
    fb = R.FmBatch(cfg)
    fb.debug_enable(True)
    iq = np.stack([R.synth.capture(kind, s, kw['rate_in'], kw.get('offset_tuning', 0), nblocks * B // 2) for s in range(n_streams)])
    oracles = [PortOracle(**kw) for _ in range(n_streams)]
    ok = True
    for b in range(nblocks):
        blk = iq[:, b * B:(b + 1) * B]
        pcm = fb.process(blk)
        dem, lr = fb.debug_read()
        for s in range(n_streams):
            p, st = oracles[s].block(blk[s], stages=True)
            print(f" block {b} stream {s}")
            ok &= cmp('dem', dem[s], st['dem'])
            ok &= cmp('lr ', lr[s, :len(st['lr'])], st['lr'])
            ok &= cmp('pcm', pcm[s], p)
    print("RESULT", label, kind, "OK" if ok else "FAIL")
    return ok

if __name__ == '__main__':
    st192 = dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=0)
    ok = check('stereo192', st192, nblocks=2, n_streams=2)
    ok &= check('stereo192-random', st192, kind='random', nblocks=2, n_streams=1)
    ok &= check('stereo192-seg4', st192, nblocks=2, n_streams=2, segments=4)
    ok &= check('mono192', dict(rate_in=192000, rate_out2=48000, mode=1, size=128, offset_tuning=0), nblocks=2, n_streams=2)
    ok &= check('stereo192-off', dict(rate_in=192000, rate_out2=48000, mode=2, size=90, offset_tuning=1), nblocks=2, n_streams=1)
    ok &= check('stereo240', dict(rate_in=240000, rate_out2=48000, mode=2, size=90, offset_tuning=0), nblocks=7, n_streams=1, kind='random')
    ok &= check('mode0', dict(rate_in=192000, rate_out2=48000, mode=0, size=90, offset_tuning=0), nblocks=2, n_streams=1)
    print("ALL", "OK" if ok else "FAIL")
    # rough timing, device-resident
    import torch
    for (ns, segs) in [(1024, 0), (1024, 1), (1024, 2), (64, 0)]:
        cfg = R.DemodConfig.stereo_192k(n_streams=ns, segments=segs)
        fb = R.FmBatch(cfg)
        fb.profile_enable(True)
        one = R.synth.batch('fm_stereo', min(ns, 16), 192000, 0, B // 2)
        iq = torch.from_numpy(np.tile(one, (ns // one.shape[0], 1))).cuda()
        pcm = torch.empty((ns, 8192), dtype=torch.int16, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            fb.process_device(iq.data_ptr(), B, pcm.data_ptr(), 8192, st)
        fb.join(st); torch.cuda.synchronize(); fb.profile_reset()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record()
        for _ in range(K):
            fb.process_device(iq.data_ptr(), B, pcm.data_ptr(), 8192, st)
        fb.join(st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        prof = fb.profile_read()
        samples = ns * B / 2
        print(f"streams={ns} segs={fb.cfg.segments if segs else 'auto'} {ms:.3f} ms/step  {samples/ms*1e-6:.1f} GS/s  "
              f"HBM-alg {samples*2.125/ms*1e-6:.1f} GB/s  demod {prof['demod_ms']/max(prof['demod_launches'],1):.3f} ms  deemph {prof['deemph_ms']/max(prof['deemph_launches'],1):.3f} ms")
        fb.close()
