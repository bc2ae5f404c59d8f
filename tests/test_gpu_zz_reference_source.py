"""GPU (runs last): the CUDA path against the reference's OWN work() - /root/reference/lib/baz_music_doa.cc compiled
unmodified against stand-in GNU Radio / Armadillo headers (oracle/_ref, see oracle/Makefile and tests/test_ref_shim.py).
The library is prebuilt in the build container and travels with the repo snapshot; where it is missing or cannot be
loaded the test is skipped (the oracle-based parity tests do not depend on it)."""
import numpy as np
import pytest

from gr_baz_b200 import synth
from gr_baz_b200.music_doa import music_doa

import helpers

pytestmark = pytest.mark.gpu

P_RTOL = 1e-5


def reference_source():
    try:
        from oracle import ref_build
        if not ref_build.available():
            pytest.skip("oracle/_ref not built")
        ref_build.lib()
        return ref_build
    except OSError as e:  # e.g. the OpenBLAS it links is not where the build container had it
        pytest.skip("oracle/_ref cannot be loaded here: %s" % e)


@pytest.mark.parametrize("base,over,W", [
    (1, {}, 24), (1, {"n": 2}, 12), (2, {"snapshots": 1024}, 12), (4, {"snapshots": 512}, 10),
    (5, {"snapshots": 512, "resolution": 1800}, 8), (1, {"m": 6, "geometry": "uca", "n": 2}, 8),
    # the BASELINE shapes at FULL size (configs[1..4]): C2 4x4096x3600, C3 8x8192x7200, C4 8x4096x3600, C5 16x4096x3600 n=2
    (2, {}, 16), (3, {}, 8), (4, {}, 8), (5, {}, 8),
])
def test_cuda_path_matches_the_reference_source(base, over, W):
    rb = reference_source()
    cfg = synth.config(base, **over)
    K, n = cfg["resolution"], cfg["n"]
    table = helpers.table_for(cfg)
    x = synth.gen_windows_numpy(cfg, 606 + base, 0, W)
    ref = rb.work_batch(x, cfg["m"], n, table)
    blk = music_doa(cfg["m"], n, cfg["nsamples"], table.tolist(), K)
    ang = np.full((W, n), -7, np.float32)
    lvl = np.full((W, n), -7, np.float32)
    spec = np.zeros((W, K), np.float32)
    assert blk.work(W, [x], [ang, lvl, spec]) == W
    assert np.array_equal(ang, ref["angles"])  # = peak bins bit-exact (angle = (float)(k * 360 / K) is injective)
    assert helpers.rel_err(lvl, ref["levels"]) <= P_RTOL
    assert helpers.rel_err(spec, ref["spectrum"]) <= P_RTOL
    # peak-only call (the fused kernel for M = 4, n = 1)
    ang2 = np.zeros_like(ang)
    assert blk.work(W, [x], [ang2]) == W and np.array_equal(ang2, ref["angles"])
