"""Multi-GPU sharding of the MUSIC DOA path.

Windows are independent (the reference block keeps no state between work() items except the
read-only table, /root/reference/lib/baz_music_doa.cc:72-161), so the stream shards
round-robin: window ``w`` -> rank ``w mod G``.  Each rank holds a full copy of the steering
table; there is no exchange during compute.  The single collective per batch is an all-gather
of the int32 peak-bin indices (NCCL over NVLink on GPUs, gloo in the CPU tests); the gathered
layout is rank-major - ``gathered[r][i]`` is stream window ``w = i*G + r`` - and
``gathered_to_stream`` restores stream order.
"""
from __future__ import annotations

import numpy as np


def shard_indices(total_windows: int, world: int, rank: int) -> np.ndarray:
    """Global window numbers owned by ``rank``: rank, rank+G, rank+2G, ..."""
    return np.arange(rank, total_windows, world, dtype=np.int64)


def shard_sizes(total_windows: int, world: int):
    return [len(range(r, total_windows, world)) for r in range(world)]


def gathered_to_stream(gathered, total_windows: int):
    """gathered: (G, Wmax, n) rank-major (padded shards); returns (total_windows, n) in stream order."""
    g = np.asarray(gathered) if not hasattr(gathered, "permute") else gathered
    G = g.shape[0]
    if hasattr(g, "permute"):
        out = g.permute(1, 0, 2).reshape(g.shape[1] * G, g.shape[2])
    else:
        out = np.transpose(g, (1, 0, 2)).reshape(g.shape[1] * G, g.shape[2])
    return out[:total_windows]


def all_gather_bins(local_bins, total_windows: int, group=None):
    """All-gather the per-rank int32 peak bins (torch tensor (W_local, n)) and return them in stream
    order on every rank.  Shards may differ by one window; they are padded with -1 to the largest."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    wmax = (total_windows + world - 1) // world
    n = local_bins.shape[1]
    pad = torch.full((wmax, n), -1, dtype=local_bins.dtype, device=local_bins.device)
    pad[: local_bins.shape[0]] = local_bins
    gathered = torch.empty((world, wmax, n), dtype=local_bins.dtype, device=local_bins.device)
    dist.all_gather_into_tensor(gathered.view(-1), pad.view(-1), group=group)
    return gathered_to_stream(gathered, total_windows)
