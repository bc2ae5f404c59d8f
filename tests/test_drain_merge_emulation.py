"""CPU: the lane merge of the fused kernel's drain workers (gr-baz_b200/csrc/music_fused.cuh, fused_drain_worker) restated in
numpy: three warp reductions per window - max of the high word of P, max of the low word among the lanes that hold it, min of the
bin among the lanes that hold both - must pick exactly what the reference's insertion loop picks when it walks the bins in ascending
order with a strict '>' (/root/reference/lib/baz_music_doa.cc:129-141): the largest strength, the lowest bin on exact ties, and
nothing (bin -1, strength 0) when no lane holds a bin."""
import numpy as np
import pytest


def redux_merge(P, k):
    """P: 32 non-negative float64 (0.0 for a lane without a bin), k: 32 int32 bins (-1 for none)."""
    bits = P.view(np.uint64)
    hi, lo = (bits >> np.uint64(32)).astype(np.uint32), (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    mh = hi.max()
    ml = np.where(hi == mh, lo, 0).max()
    cand = np.where((hi == mh) & (lo == ml) & (k >= 0), k.astype(np.uint32), np.uint32(0x7FFFFFFF))
    mk = cand.min()
    Pm = np.array([(np.uint64(mh) << np.uint64(32)) | np.uint64(ml)], np.uint64).view(np.float64)[0]
    return Pm, (-1 if mk == 0x7FFFFFFF else int(mk))


def reference_rule(P, k):
    """the reference's loop over ALL bins in ascending order: replace iff P > best (initial pair (0, 0) -> bin -1 here)"""
    best, bk = 0.0, -1
    for i in np.argsort(k, kind="stable"):
        if k[i] >= 0 and P[i] > best:
            best, bk = P[i], int(k[i])
    return best, bk


@pytest.mark.parametrize("seed", range(6))
def test_redux_merge_equals_the_reference_rule(seed):
    rng = np.random.default_rng(seed)
    for trial in range(300):
        k = rng.permutation(4000)[:32].astype(np.int32)
        P = rng.random(32) * 10.0 ** rng.integers(-3, 6)
        kind = trial % 6
        if kind == 1:    # exact ties between several lanes (mirror bins of an x-axis ULA)
            P[rng.integers(0, 32, 5)] = P.max() * 1.0
            P[rng.integers(0, 32, 3)] = P.max()
        elif kind == 2:  # strengths that differ in the last bit only
            P[:] = 3.25
            P[rng.integers(0, 32)] = np.nextafter(3.25, 4.0)
        elif kind == 3:  # lanes without a bin
            none = rng.random(32) < 0.5
            P[none] = 0.0
            k[none] = -1
        elif kind == 4:  # no lane holds a bin
            P[:] = 0.0
            k[:] = -1
        elif kind == 5:  # infinite strength (d = 0) beats everything, lowest bin among them
            P[rng.integers(0, 32, 2)] = np.inf
        got = redux_merge(P.copy(), k.copy())
        ref = reference_rule(P, k)
        assert got[1] == ref[1], (trial, got, ref)
        assert (got[0] == ref[0]) or (ref[1] < 0 and got[0] == 0.0)
