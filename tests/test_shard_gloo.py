"""N>1 host logic on CPU: world_size-2 gloo run of the stream sharding + PCM gather.
The per-rank compute stand-in is the CPU oracle (tests may use it); the product's CUDA path is
covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from rtl_fm_player_b200.shard import concat_shards, owner_of, shard_range, shard_sizes

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_ranges_partition_the_streams():
    for n in (0, 1, 7, 64, 1000, 1024, 8192):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
            for s in range(0, n, max(1, n // 17)):
                lo, hi = shard_range(n, w, owner_of(s, n, w))
                assert lo <= s < hi
    assert shard_range(8192, 8, 3) == (3072, 4096)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_concat_shards():
    a, b = np.ones((2, 5), np.int16), 2 * np.ones((3, 5), np.int16)
    out = concat_shards([a, np.empty((0, 5), np.int16), b])
    assert out.shape == (5, 5) and out[:2].max() == 1 and out[2:].min() == 2
    with pytest.raises(ValueError):
        concat_shards([a, np.ones((1, 4), np.int16)])


def _worker(rank, world, port, n_streams, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from oracle.oracle_py import PortOracle
    from rtl_fm_player_b200 import synth
    from rtl_fm_player_b200.shard import gather_pcm, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_streams, world, rank)
    kw = dict(rate_in=192000, rate_out2=48000, mode=2, size=90)
    local = np.stack([PortOracle(**kw).run(synth.capture("fm_stereo", s, 192000, 0, 32768 // 2), block_bytes=32768)
                      for s in range(lo, hi)])
    full = gather_pcm(local, n_streams, dst=0)
    dist.barrier()
    if rank == 0:
        q.put(full)
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_unsharded():
    from oracle.oracle_py import PortOracle
    from rtl_fm_player_b200 import synth
    n_streams, world = 5, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    kw = dict(rate_in=192000, rate_out2=48000, mode=2, size=90)
    want = np.stack([PortOracle(**kw).run(synth.capture("fm_stereo", s, 192000, 0, 32768 // 2), block_bytes=32768)
                     for s in range(n_streams)])
    assert full.shape == want.shape and np.array_equal(full, want)
