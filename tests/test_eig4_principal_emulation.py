"""CPU: the principal-eigenvector solver of the fused kernel (gr-baz_b200/csrc/music_eig4p.cuh) restated in numpy
(tools/emulate_eig4_principal.py) against LAPACK on covariance matrices of the synthetic stream: the vector, the
orthonormal complement basis that replaces the three noise eigenvectors, power-of-two scaling invariance and the
cases that must fall back to the Jacobi solver."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import emulate_eig4_principal as emu  # noqa: E402

from gr_baz_b200 import synth  # noqa: E402


def covariances(cfg, seed, W):
    x = synth.gen_windows_numpy(cfg, seed, 0, W)
    M, N = cfg["m"], cfg["snapshots"]
    for w in range(W):
        X = x[w].reshape(N, M).T.astype(np.complex128)
        yield X @ X.conj().T / N


@pytest.mark.parametrize("base,snr,max_sq", [(2, 20.0, 3), (2, 40.0, 3), (1, 0.0, 5), (2, -10.0, 9)])
def test_principal_vector_and_complement_match_lapack(base, snr, max_sq):
    cfg = synth.config(base, snr_db=snr)
    for R in covariances(cfg, 77, 24):
        Vt, nsq = emu.principal(R)
        assert Vt is not None and nsq <= max_sq
        ev, V = np.linalg.eigh(R)
        e = V[:, 3] * np.conj(V[0, 3]) / abs(V[0, 3])
        assert np.linalg.norm(Vt[3] - e) <= 4e-15
        assert Vt[3][0].imag == 0.0 and Vt[3][0].real >= 0.0
        Q = Vt.T
        assert np.max(np.abs(Q.conj().T @ Q - np.eye(4))) <= 2e-15
        # the three complement vectors span the noise subspace: same projector as LAPACK's three smallest eigenvectors
        G = Vt[:3].T
        Gl = V[:, :3]
        assert np.max(np.abs(G @ G.conj().T - Gl @ Gl.conj().T)) <= 4e-15


def test_power_of_two_scaling_gives_identical_vectors():
    cfg = synth.config(2)
    for R in covariances(cfg, 5, 6):
        a, _ = emu.principal(R)
        for k in (-40, 3, 64):
            b, _ = emu.principal(R * 2.0 ** k)
            assert np.array_equal(a.view(np.float64), b.view(np.float64))


def test_fallback_cases():
    assert emu.principal(np.zeros((4, 4)))[0] is None                      # R = 0: no trace to scale by
    assert emu.principal(np.eye(4))[0] is None                             # degenerate: never becomes rank one
    R = np.eye(4, dtype=np.complex128)
    R[1, 1] = np.nan
    assert emu.principal(R)[0] is None
    rng = np.random.default_rng(1)
    X = rng.standard_normal((4, 4096)) + 1j * rng.standard_normal((4, 4096))  # noise only: eigenvalue ratio ~1.02
    Vt, nsq = emu.principal(X @ X.conj().T / 4096)
    assert Vt is None or nsq >= 9


def test_eight_antennas():
    """the same algorithm on the 8 x 8 covariance (music_fused8.cuh)"""
    cfg = synth.config(4)
    for R in covariances(cfg, 11, 8):
        Vt, nsq = emu.principal(R)
        assert Vt is not None and nsq <= 4
        ev, V = np.linalg.eigh(R)
        e = V[:, 7] * np.conj(V[0, 7]) / abs(V[0, 7])
        assert np.linalg.norm(Vt[7] - e) <= 4e-15
        G, Gl = Vt[:7].T, V[:, :7]
        assert np.max(np.abs(G @ G.conj().T - Gl @ Gl.conj().T)) <= 4e-15
