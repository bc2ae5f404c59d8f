"""N > 1 host logic on CPU: world_size-2 gloo run of the round-robin shard -> compute -> all-gather ->
de-interleave path (the compute function here is the oracle; on GPUs it is the CUDA library)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_baz_b200 import sharding, synth
from oracle import c_oracle as co

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.config(1)
    table = helpers.table_for(cfg)
    idx = sharding.shard_indices(total, world, rank)
    x = synth.gen_windows_numpy(cfg, 77, indices=idx)
    bins = torch.from_numpy(co.work_batch(x, cfg["m"], cfg["n"], table, want_P=False)["bins"])
    stream = sharding.all_gather_bins(bins, total)
    np.save(os.path.join(out_dir, "bins_%d.npy" % rank), stream.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_shard_and_gather_world2(tmp_path):
    total, world = 11, 2  # odd: shards of 6 and 5 windows
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    cfg = synth.config(1)
    table = helpers.table_for(cfg)
    x = synth.gen_windows_numpy(cfg, 77, 0, total)
    ref = co.work_batch(x, cfg["m"], cfg["n"], table, want_P=False)["bins"]
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "bins_%d.npy" % r))
        assert got.shape == ref.shape and np.array_equal(got, ref)


def test_index_maps():
    assert sharding.shard_indices(10, 4, 1).tolist() == [1, 5, 9]
    assert sharding.shard_sizes(10, 4) == [3, 3, 2, 2]
    g = np.full((3, 2, 1), -1, np.int32)  # G=3, Wmax=2
    for r in range(3):
        for i, w in enumerate(range(r, 5, 3)):
            g[r, i, 0] = w
    assert sharding.gathered_to_stream(g, 5)[:, 0].tolist() == [0, 1, 2, 3, 4]
