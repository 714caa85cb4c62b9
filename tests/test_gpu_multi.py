"""The C multi-GPU host (include/fmb_multi.h, csrc/fmb_multi.c): channels sharded by index over devices, one worker
thread per device, every device's copies landing in its slice of ONE pinned buffer.  Each shard's PCM is compared
with the oracle ON HARDWARE -- on every visible GPU when there are several (BASELINE.json configs[4]: parity on 8
streams per GPU), and as several shards sharing GPU 0 otherwise, so the host logic is covered on a 1-GPU box too."""
import numpy as np
import pytest

import rtl_fm_player_b200 as R
from oracle.oracle_py import PortOracle
from vectors import B, CONFIGS, make_input

pytestmark = pytest.mark.gpu


def _devices(min_shards):
    n = R.device_count()
    return list(range(n)) if n >= min_shards else [0] * min_shards


@pytest.mark.parametrize("cfgname,kind", [("stereo192", "fm_stereo"), ("stereo240", "random"), ("mono192", "fm_mono")])
def test_every_shard_matches_the_oracle_eight_streams_per_gpu(cfgname, kind):
    devs = _devices(2)
    per, blocks = 8, 6 if cfgname == "stereo240" else 3
    n = per * len(devs)
    kw = CONFIGS[cfgname]
    iq = np.stack([make_input(cfgname, kind, s, blocks) for s in range(n)])     # every stream distinct
    with R.FmMulti(R.DemodConfig(n_streams=n, **kw), devs) as fm:
        assert [(f, c) for f, c, _ in fm.shards] == [(g * per, per) for g in range(len(devs))]
        assert [d for _, _, d in fm.shards] == devs
        pcm = fm.run(iq)
    for s in range(n):
        assert np.array_equal(pcm[s], PortOracle(**kw).run(iq[s])), f"stream {s} (shard {s // per}, device {devs[s // per]})"


def test_uneven_shards_and_pipelined_submits_into_one_pinned_buffer():
    """13 streams over 3 shards (4, 4, 5); three submits in flight; PCM lands in the caller's ONE buffer per step."""
    devs = _devices(3)[:3]
    n, blocks = 13, 5
    kw = CONFIGS["stereo192"]
    iq = np.stack([make_input("stereo192", "random", s, blocks) for s in range(n)])
    with R.FmMulti(R.DemodConfig(n_streams=n, **kw), devs) as fm:
        assert [c for _, c, _ in fm.shards] == [4, 4, 5]
        n_out = fm.next_out_count()
        pitch = (n_out + 7) & ~7
        h_in = [R.pinned_array((n, B), np.uint8) for _ in range(blocks)]
        h_out = [R.pinned_array((n, pitch), np.int16) for _ in range(blocks)]
        for b in range(blocks):
            h_in[b][...] = iq[:, b * B:(b + 1) * B]
            h_out[b][...] = -1
        tickets = []
        for b in range(blocks):
            if len(tickets) == 3:
                fm.wait(tickets.pop(0))
            tickets.append(fm.submit(h_in[b].ctypes.data, B, h_out[b].ctypes.data, pitch))
        with pytest.raises(R.FmbError):
            fm.wait(99)
        for t in tickets:
            fm.wait(t)
        got = np.concatenate([h[:, :n_out] for h in h_out], axis=1)
        for a in h_in + h_out:
            R.free_pinned(a)
    for s in range(n):
        assert np.array_equal(got[s], PortOracle(**kw).run(iq[s])), s


def test_device_resident_multi_and_reset():
    import torch
    devs = _devices(2)
    per, blocks = 16, 3
    n = per * len(devs)
    kw = CONFIGS["stereo192"]
    iq = np.stack([make_input("stereo192", "fm_stereo", s % 4, blocks) for s in range(n)])
    want = [PortOracle(**kw).run(iq[s]) for s in range(4)]
    with R.FmMulti(R.DemodConfig(n_streams=n, **kw), devs) as fm:
        d_in, d_out = [], []
        for g, (first, count, dev) in enumerate(fm.shards):
            d_in.append(torch.from_numpy(iq[first:first + count]).to(f"cuda:{dev}"))
            d_out.append(torch.zeros((blocks, count, 8192), dtype=torch.int16, device=f"cuda:{dev}"))
        for d in set(devs):
            torch.cuda.synchronize(d)
        for rep in range(2):
            for b in range(blocks):
                fm.process_device([t.data_ptr() + b * B for t in d_in], iq.shape[1], [t[b].data_ptr() for t in d_out], 8192)
            fm.sync()
            for g, (first, count, dev) in enumerate(fm.shards):
                got = np.concatenate([d_out[g][b].cpu().numpy() for b in range(blocks)], axis=1)
                for s in range(count):
                    assert np.array_equal(got[s], want[(first + s) % 4]), (rep, g, s)
            fm.reset()


def test_create_fails_as_a_whole():
    with pytest.raises(R.FmbError) as e:
        R.FmMulti(R.DemodConfig.stereo_192k(n_streams=4), [0, 4096])
    assert "shard 1" in str(e.value)
    with pytest.raises(R.FmbError):
        R.FmMulti(R.DemodConfig.stereo_192k(n_streams=1), [0, 0])          # fewer streams than devices
