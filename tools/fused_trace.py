#!/usr/bin/env python
"""Per-CTA clock64 trace of the fused kernel (MUSIC_B200_TRACE=1): when each role finished and how
busy the eigensolver / scan warps were.  Run on the GPU box: python tools/fused_trace.py"""
import os
import sys

os.environ["MUSIC_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from gr_baz_b200 import synth
from gr_baz_b200.music_doa import music_doa

cfg, W = bench.workload(2)
resp, _ = bench.table_for(cfg)
dev = torch.device("cuda:0")
d_in = synth.gen_windows_torch(cfg, synth.BASE_SEED + 2, 0, W, dev)
blk = music_doa(cfg["m"], cfg["n"], cfg["nsamples"], resp, cfg["resolution"])
a = torch.empty((W, 1), dtype=torch.float32, device=dev)
l = torch.empty_like(a)
b = torch.empty((W, 1), dtype=torch.int32, device=dev)
for _ in range(3):
    blk.process_device(d_in.data_ptr(), W, a.data_ptr(), l.data_ptr(), None, b.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
tr = np.zeros((148, 32), np.int64)
blk._lib.music_b200_debug_fused_trace(blk._h, tr.ctypes.data, 148)
cov_end = tr[:, 0]
print("cycles (mean over CTAs):")
print("  last cov warp done   %9.0f   (first cov warp done %9.0f)" % (cov_end.mean(), tr[:, 1].mean()))
print("  eig warp exit        %9.0f   busy %9.0f in %5.1f rounds -> %7.0f cyc/round" % (tr[:, 8].mean(), tr[:, 9].mean(), tr[:, 10].mean(), (tr[:, 9] / np.maximum(tr[:, 10], 1)).mean()))
print("  scan warps exit      %9.0f   busy %9.0f in %5.1f passes -> %7.0f cyc/pass" % (tr[:, 11].mean(), tr[:, 12].mean(), tr[:, 13].mean(), (tr[:, 12] / np.maximum(tr[:, 13], 1)).mean()))
print("  scan: exact-evaluation phase %9.0f cycles, candidates/CTA %7.1f, full-fp64 fallbacks/CTA %5.2f" % (tr[:, 14].mean(), tr[:, 7].mean(), tr[:, 15].mean()))
print("  scan thread 0 inside the sweeps: issue+slot-wait %9.0f, tile-arrival wait %9.0f, loads+MMA+post %9.0f" % (tr[:, 4].mean(), tr[:, 5].mean(), tr[:, 6].mean()))
print("  CTA lifetime by %%globaltimer: %.1f us mean -> SM clock %.0f MHz; CTA start skew %.1f us; first start to last end %.1f us" % (tr[:, 3].mean() / 1e3, (tr[:, 11] / (tr[:, 3] / 1e3)).mean(), (tr[:, 2].max() - tr[:, 2].min()) / 1e3, ((tr[:, 2] + tr[:, 3]).max() - tr[:, 2].min()) / 1e3))
print("  tail after last cov  %9.0f   (eig warp exit - last cov %9.0f, scan exit - eig exit %9.0f)" % ((tr[:, 11] - cov_end).mean(), (tr[:, 8] - cov_end).mean(), (tr[:, 11] - tr[:, 8]).mean()))
print("  tensor-core passes ended %9.0f, windows taken by the fp64 drain workers %5.1f / CTA, Jacobi rounds %5.2f / CTA" % (tr[:, 16].mean(), tr[:, 17].mean(), tr[:, 19].mean()))
print("  CTA done (all outputs written) %9.0f mean, %9.0f max; after the last covariance warp: %9.0f mean" % (tr[:, 18].mean(), tr[:, 18].max(), (tr[:, 18] - cov_end).mean()))
print("  drain: %5.2f groups / CTA, unit busy cycles (sum over warps) %9.0f -> %7.0f per group-unit (%d units per group); last group finished %9.0f (%9.0f after the last covariance warp)" % (tr[:, 20].mean(), tr[:, 22].mean(), (tr[:, 22] / np.maximum(tr[:, 20] * 16, 1)).mean(), 16, tr[:, 23].mean(), (tr[:, 23] - cov_end).mean()))
u = np.maximum(tr[:, 29], 1)
print("  drain units: %5.1f / CTA; cycles per unit: table sweep %7.0f, lane merge + publish %7.0f; units that found lane 0's warp split at the sweep %5.1f / at the merge %5.1f; worker warps waiting for a unit: %9.0f cycles / CTA (sum over 16 warps)"
      % (tr[:, 29].mean(), (tr[:, 25] / u).mean(), (tr[:, 26] / u).mean(), tr[:, 24].mean(), tr[:, 27].mean(), tr[:, 28].mean()))
