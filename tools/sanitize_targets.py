#!/usr/bin/env python
"""Small invocations of the kernels compute-sanitizer should look at: run as
  compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_targets.py <target>
targets: fused (C1-size windows, 2000 of them + a ragged 1187 + a drain-only 9), fused_jacobi, fused8, covn8, covn16, eig16, multi"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from gr_baz_b200 import synth
from gr_baz_b200.music_doa import music_doa

target = sys.argv[1] if len(sys.argv) > 1 else "fused"
dev = torch.device("cuda:0")


def run(cfg, counts, devices=None):
    resp, _ = bench.table_for(cfg)
    Wmax = max(counts)
    d_in = synth.gen_windows_torch(cfg, 1234, 0, Wmax, dev)
    blk = music_doa(cfg["m"], cfg["n"], cfg["nsamples"], resp, cfg["resolution"], devices=devices)
    a = torch.empty((Wmax, cfg["n"]), dtype=torch.float32, device=dev)
    l = torch.empty_like(a)
    b = torch.empty((Wmax, cfg["n"]), dtype=torch.int32, device=dev)
    for W in counts:
        blk.process_device(d_in.data_ptr(), W, a.data_ptr(), l.data_ptr(), None, b.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    print(target, "ok: bins[:8] =", b[:8, 0].cpu().numpy().tolist(), "launches", blk.launch_count())
    blk.close()


if target in ("fused", "fused_jacobi"):
    if target == "fused_jacobi":
        os.environ["MUSIC_B200_EIG"] = "jacobi"
    run(synth.config(1), (2000, 1187, 9))
elif target in ("covn8", "fused8"):  # M = 8, n = 1 takes the fused M = 8 kernel by default; covn8 = the three-kernel path
    if target == "covn8":
        os.environ["MUSIC_B200_FUSED"] = "0"
    run(synth.config(4, snapshots=1024, resolution=360), (600, 37))
elif target == "covn16":
    run(synth.config(5, snapshots=512, resolution=360), (300, 19))
elif target == "eig16":
    run(synth.config(5, snapshots=128, resolution=180), (64,))
elif target == "multi":
    cfg = synth.config(1)
    resp, table = bench.table_for(cfg)
    W = 300
    x = synth.gen_windows_numpy(cfg, 7, 0, W)
    blk = music_doa(4, 1, cfg["nsamples"], resp, cfg["resolution"], devices=list(range(torch.cuda.device_count())))
    ang = np.zeros((W, 1), np.float32)
    blk.work(W, [x], [ang])
    print("multi ok", ang[:4, 0])
    blk.close()
