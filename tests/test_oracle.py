"""CPU tests: the oracle (numpy/LAPACK restatement + C/Jacobi restatement) against our golden
vectors and against each other; reference semantics the CUDA path must reproduce."""
import hashlib
import os

import numpy as np
import pytest

from gr_baz_b200 import synth
from oracle import c_oracle as co
from oracle import music_oracle as mo

import helpers


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("path", helpers.golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracles_match_golden(path):
    cfg, seed, table, wins, table_sha = helpers.load_golden(path)
    assert _sha(table) == table_sha  # steering builder (helper restatement) has not drifted
    for g in wins:
        assert _sha(g["in"]) == g["in_sha256"]  # synthetic generator has not drifted
        for impl in (mo, co):
            r = impl.work(g["in"], cfg["m"], cfg["n"], table, want_spectrum=True, return_internals=True)
            assert np.array_equal(r["bins"], g["bins"])  # bit-exact peak bins
            assert np.array_equal(r["angles"], g["angles"])
            assert helpers.rel_err(r["P"], g["P"]) <= 1e-9  # gate for the product is 1e-5
            assert helpers.rel_err(r["spectrum"], g["spectrum"]) <= 1e-6
            assert helpers.rel_err(r["levels"], g["levels"]) <= 1e-6
            assert np.max(np.abs(r["R"] - g["R"])) <= 1e-12 * np.max(np.abs(g["R"]))
            assert np.max(np.abs(r["eigvals"] - g["eigvals"])) <= 1e-12 * np.max(np.abs(g["eigvals"]))
            assert np.max(np.abs(r["noise_projector"] - g["noise_projector"])) <= 1e-10


def test_c_jacobi_matches_lapack():
    rng = np.random.default_rng(7)
    for M in (2, 3, 4, 8, 16):
        for _ in range(5):
            X = rng.standard_normal((M, 3 * M)) + 1j * rng.standard_normal((M, 3 * M))
            A = X @ X.conj().T
            w, V = co.herm_eig(A)
            wl, _ = np.linalg.eigh(A)
            assert np.all(np.diff(w) >= 0)
            assert np.max(np.abs(w - wl)) <= 1e-12 * wl[-1]
            assert np.max(np.abs(V.conj().T @ V - np.eye(M))) <= 1e-13
            assert np.max(np.abs(A @ V - V * w)) <= 1e-12 * wl[-1]


def test_interleaved_layout_and_covariance():
    # x(r, c) = in[c*M + r]  (lib/baz_music_doa.cc:82-85)
    rng = np.random.default_rng(1)
    M, N = 3, 5
    x = (rng.standard_normal((M, N)) + 1j * rng.standard_normal((M, N))).astype(np.complex64)
    inter = x.T.reshape(-1)  # sample-interleaved
    R = mo.covariance(inter, M)
    want = x.astype(np.complex128) @ x.astype(np.complex128).conj().T / N
    assert np.allclose(R, want, rtol=0, atol=1e-15)


def test_top_n_semantics():
    # strict '>' insertion (lib/baz_music_doa.cc:129-141): n largest bins, ties -> lower bin
    # first, NaN and non-positive values never inserted, unfilled slots stay (0, 0).
    P = np.array([1.0, 5.0, 5.0, np.nan, 3.0, 0.0, -2.0, np.inf, 5.0])
    for n in (1, 2, 3, 4, 6, 8):
        a = mo.pick_top_n_literal(P, n, len(P))
        b = mo.pick_top_n(P, n, len(P))
        assert a == b
    top4 = mo.pick_top_n(P, 4, len(P))
    assert [t[2] for t in top4] == [7, 1, 2, 8]
    assert [t[2] for t in mo.pick_top_n(np.array([np.nan, 0.0, -1.0]), 2, 3)] == [-1, -1]
    rng = np.random.default_rng(3)
    for _ in range(50):
        P = rng.integers(0, 6, size=40).astype(np.float64)  # many exact ties
        assert mo.pick_top_n_literal(P, 5, 40) == mo.pick_top_n(P, 5, 40)


def test_mirror_tie_picks_lower_bin():
    # x-axis ULA: P[k] == P[K-k] bit-for-bit; the reference keeps the lower bin.
    cfg = synth.config(1)
    table = helpers.table_for(cfg)
    x = synth.gen_windows_numpy(cfg, 1234, 0, 8)
    for w in range(8):
        r = mo.work(x[w], 4, 1, table)
        k = int(r["bins"][0])
        K = cfg["resolution"]
        assert r["P"][k] == r["P"][(K - k) % K]
        assert k <= K // 2
        c = co.work(x[w], 4, 1, table)
        assert c["P"][k] == c["P"][(K - k) % K]
        assert int(c["bins"][0]) == k


def test_zero_window_and_single_snapshot():
    cfg = synth.config(1)
    table = helpers.table_for(cfg)
    z = np.zeros(cfg["nsamples"], np.complex64)
    a = mo.work(z, 4, 1, table)
    b = co.work(z, 4, 1, table)
    assert np.array_equal(a["bins"], b["bins"])
    assert helpers.rel_err(b["P"], a["P"]) < 1e-12
    # one snapshot (nsamples == m): rank-1 R
    x = synth.gen_windows_numpy(synth.config(1, snapshots=1), 5, 0, 1)[0]
    a = mo.work(x, 4, 1, table)
    b = co.work(x, 4, 1, table)
    assert np.array_equal(a["bins"], b["bins"])


def test_param_checks():
    cfg = synth.config(1)
    table = helpers.table_for(cfg)
    x = np.zeros(cfg["nsamples"], np.complex64)
    for m, n in ((4, 4), (4, 0), (4, 5)):
        with pytest.raises(ValueError):
            mo.work(x, m, n, table)
    with pytest.raises(ValueError):
        mo.work(x[:-1], 4, 1, table)
    with pytest.raises(ValueError):
        co.work(x, 4, 4, table)


def test_steering_table_matches_closed_form():
    # python/music_doa_helper.py:32-46 restated literally vs the vectorised closed form
    for geom, m in (("ula_x", 4), ("uca", 8)):
        cfg = synth.config(1, geometry=geom, m=m)
        t = helpers.table_for(cfg)
        k = np.arange(cfg["resolution"])
        want = synth.steering(cfg["antenna_array"], k * 360.0 / cfg["resolution"]).astype(np.complex64)
        assert t.shape == (cfg["resolution"], m) and t.dtype == np.complex64
        assert np.max(np.abs(t - want)) <= 2e-7


def test_angle_is_injective_in_float32():
    for K in (360, 3600, 7200):
        ang = (np.arange(K) * 360.0 / K).astype(np.float32)
        assert np.array_equal(np.rint(ang.astype(np.float64) * K / 360.0).astype(np.int64), np.arange(K))


def test_pick_local_maxima_rule():
    """the opt-in peak rule (not in the reference): definition checks on hand-made spectra"""
    K = 12
    P = np.array([1, 5, 5, 2, 1, 9, 3, 3, 7, 1, 0.5, 4], np.float64)
    # local maxima: k=1 (plateau 5,5 -> lowest bin), k=5, k=8, k=11 (4 > 0.5 and 4 >= P[0]=1, circular)
    got = mo.pick_local_maxima(P, 4, K)
    assert [g[2] for g in got] == [5, 8, 1, 11]
    assert got[0][0] == 5 * 360.0 / K and got[0][1] == 9.0
    # exclusion: 8 is 3 bins from 5 -> dropped with exclusion 3; 1 is 4 away -> kept; 11 is 2 from 1 (circular) -> dropped
    assert [g[2] for g in mo.pick_local_maxima(P, 4, K, exclusion=3)] == [5, 1, -1, -1]
    assert [g[2] for g in mo.pick_local_maxima(P, 2, K, exclusion=0)] == [5, 8]
    # NaN and non-positive bins are never candidates and never beat a neighbour
    Q = P.copy(); Q[5] = np.nan; Q[8] = 0.0
    assert [g[2] for g in mo.pick_local_maxima(Q, 3, K)] == [1, 11, 9]  # bin 6 fails 3 > NaN; bin 9 (1 > 0, 1 >= 0.5) qualifies
