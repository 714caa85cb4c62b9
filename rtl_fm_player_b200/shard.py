"""Stream sharding across the GPUs of one box.

FM channels share no state (each has its own demod_state in the reference,
include/rtl_fm_player.h:127-175), so the batch is partitioned by stream index into
contiguous ranges, one per GPU / process, with NO data-path collective.  The host gathers
PCM by concatenating the per-rank slices in rank order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """Streams [lo, hi) owned by `rank`: lo = rank*S//G, hi = (rank+1)*S//G."""
    if world < 1 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad shard arguments")
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


def shard_sizes(n_streams: int, world: int) -> List[int]:
    return [shard_range(n_streams, world, r)[1] - shard_range(n_streams, world, r)[0] for r in range(world)]


def owner_of(stream: int, n_streams: int, world: int) -> int:
    """Rank that owns `stream` under shard_range."""
    if not (0 <= stream < n_streams):
        raise ValueError("stream out of range")
    r = min(world - 1, (stream * world) // max(n_streams, 1))
    while stream < shard_range(n_streams, world, r)[0]:
        r -= 1
    while stream >= shard_range(n_streams, world, r)[1]:
        r += 1
    return r


def concat_shards(parts: Sequence[np.ndarray]) -> np.ndarray:
    """Host-side gather: per-rank PCM [n_local, n_out] in rank order -> [n_streams, n_out]."""
    parts = [p for p in parts if p.shape[0] > 0]
    if not parts:
        return np.empty((0, 0), dtype=np.int16)
    n_out = parts[0].shape[1]
    if any(p.shape[1] != n_out for p in parts):
        raise ValueError("shards disagree on the output length")
    return np.concatenate(parts, axis=0)


def gather_pcm(local_pcm: np.ndarray, n_streams: int, dst: int = 0):
    """Gather per-rank PCM to `dst` over torch.distributed (any backend; gloo for host arrays).

    Returns the [n_streams, n_out] array on `dst`, None elsewhere.  This is the only
    cross-rank traffic of the framework and it is OFF the demodulation path: 0.125 bytes per
    IQ sample.
    """
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(n_streams, world)
    if local_pcm.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local_pcm.shape[0]} streams, expected {sizes[rank]}")
    n_out = int(local_pcm.shape[1])
    # gloo has no int16 collectives: ship the PCM as raw bytes
    t = torch.from_numpy(np.ascontiguousarray(local_pcm).view(np.uint8).reshape(sizes[rank], 2 * n_out))
    pad = max(sizes)
    buf = torch.zeros((pad, 2 * n_out), dtype=torch.uint8)
    buf[: sizes[rank]] = t
    if rank == dst:
        outs = [torch.zeros((pad, 2 * n_out), dtype=torch.uint8) for _ in range(world)]
        dist.gather(buf, outs, dst=dst)
        return concat_shards([o[: sizes[r]].numpy().view(np.int16) for r, o in enumerate(outs)])
    dist.gather(buf, None, dst=dst)
    return None
