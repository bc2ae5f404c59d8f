"""CPU: the four-lanes-per-window Jacobi of the fused kernel's eigensolver warp
(gr-baz_b200/csrc/music_kernels.cuh::herm_eig4_coop) restated lane by lane in Python - same XOR-relative slot layout,
same shuffle pattern, same four phases per step - must reproduce the one-lane solver (herm_eig_body<4, true>) bit for
bit, sweep counts included.  The GPU-side check of the same property is
tests/test_gpu_parity.py::test_fused_kernel_matches_unfused_path_on_every_window."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_eig4_coop as emu  # noqa: E402

from gr_baz_b200 import synth  # noqa: E402
from oracle import music_oracle as mo  # noqa: E402


def check(R):
    s = emu.eig_seq(R)
    c = emu.eig_coop(R)
    assert s[4] == c[4]
    for a, b in zip(s[:4], c[:4]):
        assert np.array_equal(a, b)
    return s


def test_coop_equals_sequential_on_random_hermitian_matrices():
    rng = np.random.default_rng(11)
    for trial in range(120):
        X = rng.standard_normal((4, 48)) + 1j * rng.standard_normal((4, 48))
        R = X @ X.conj().T / 48
        R = (R + R.conj().T) / 2
        R[np.diag_indices(4)] = R[np.diag_indices(4)].real
        check(R)


def test_coop_equals_sequential_on_music_covariances_and_diagonalises_them():
    cfg = synth.config(1)
    x = synth.gen_windows_numpy(cfg, 5, 0, 24)
    for w in range(24):
        R = mo.covariance(x[w], 4)
        Ar, Ai, Vr, Vi, sweeps = check(R)
        V = Vr + 1j * Vi
        assert sweeps <= 6
        assert np.allclose(V.conj().T @ R @ V, np.diag(np.diag(Ar)), atol=1e-12 * np.abs(R).max())


def test_degenerate_inputs():
    check(np.diag([3.0, 1.0, 2.0, 1.0]).astype(complex))  # already diagonal: zero sweeps
    check(np.zeros((4, 4), complex))
    R = np.ones((4, 4), complex)  # rank one, exact ties between three zero eigenvalues
    check(R)
