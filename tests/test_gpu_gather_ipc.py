"""GPU (>= 2 devices; skipped on a one-GPU box): the fused all-gather of the peak bins with ONE PROCESS PER GPU, as
bench.py runs it under torchrun - CUDA-IPC-mapped gather buffers, peer stores from the scan epilogue, epoch flags -
against ncclAllGather of the same bins (the reference implementation it replaces) and against the oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from gr_baz_b200 import sharding, synth
    from gr_baz_b200.music_doa import music_doa

    import helpers

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    results = {}
    for name, base, over, Wl in (("fused_m4", 2, {}, 333), ("unfused_m8", 4, {"snapshots": 512}, 40), ("n2", 1, {"n": 2, "geometry": "uca"}, 25)):
        cfg = synth.config(base, **over)
        table = helpers.table_for(cfg)
        n = cfg["n"]
        total = Wl * world
        idx = sharding.shard_indices(total, world, rank)
        x = synth.gen_windows_numpy(cfg, 4711 + base, indices=idx)
        blk = music_doa(cfg["m"], n, cfg["nsamples"], table.tolist(), cfg["resolution"], device=rank)
        mine = torch.from_numpy(np.frombuffer(blk.gather_create(total), dtype=np.uint8).copy()).to(dev)
        allh = torch.empty((world, 128), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh.view(-1), mine)
        blk.gather_attach(world, rank, [allh[r].cpu().numpy().tobytes() for r in range(world)])
        d_in = torch.from_numpy(x.view(np.float32)).to(dev)
        d_ang = torch.empty((Wl, n), dtype=torch.float32, device=dev)
        d_bins = torch.empty((Wl, n), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):  # epochs 1..3
            blk.process_device(d_in.data_ptr(), Wl, d_ang.data_ptr(), None, None, d_bins.data_ptr(), stream=st)
        blk.gather_wait(st)
        torch.cuda.synchronize()
        got = blk.gather_read(total)
        ref = sharding.all_gather_bins(d_bins, total).cpu().numpy()  # ncclAllGather + de-interleave
        results[name] = (got, ref)
        dist.barrier()  # nobody unmaps while a peer may still be storing
        blk.close()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k + "_got": v[0] for k, v in results.items()},
             **{k + "_ref": v[1] for k, v in results.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_ipc_peer_store_gather_equals_nccl_all_gather(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from gr_baz_b200 import synth
    from oracle import c_oracle as co

    import helpers

    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for name in ("fused_m4", "unfused_m8", "n2"):
            assert np.array_equal(z[name + "_got"], z[name + "_ref"]), (r, name)
    # and the gathered stream is the oracle's
    cfg = synth.config(2)
    x = synth.gen_windows_numpy(cfg, 4711 + 2, 0, 333 * world)
    ref = co.work_batch(x, 4, 1, helpers.table_for(cfg), want_P=False)["bins"]
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "rank0.npz"))["fused_m4_got"], ref)
