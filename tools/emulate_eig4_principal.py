#!/usr/bin/env python
"""CPU emulation (numpy fp64) of gr-baz_b200/csrc/music_eig4p.cuh: principal eigenvector of the 4 x 4 covariance by
repeated squaring + two power steps + residual certificate, and the Householder basis of its orthogonal complement.
Used by tests/test_eig4_principal_emulation.py; prints accuracy against LAPACK when run directly."""
import numpy as np

MAXSQ = 12


def pow2_scale(t):
    if not (np.isfinite(t) and t >= 2.0 ** -1022):
        return None
    e = int(np.floor(np.log2(t)))
    while 2.0 ** e > t:
        e -= 1
    while 2.0 ** (e + 1) <= t:
        e += 1
    if e >= 1023:
        return None
    return 2.0 ** (-e)


def principal(R):
    """Returns (Vt, nsq): Vt[rank][i] with ranks 0..M-2 = complement basis, rank M-1 = principal eigenvector, or
    (None, nsq) where the CUDA code falls back to Jacobi.  M = 4: music_eig4p.cuh; M = 8: music_fused8.cuh."""
    R = np.asarray(R, np.complex128)
    M = R.shape[0]
    sc = pow2_scale(np.trace(R).real)
    if sc is None:
        return None, 0
    A0 = R * sc
    A = A0.copy()
    st, nsq = 0, 0
    for _ in range(MAXSQ):
        if st >= 2:
            break
        N = A @ A
        N[np.diag_indices(M)] = N[np.diag_indices(M)].real
        nsq += 1
        t = np.trace(N).real
        f = float(np.sum(np.abs(N) ** 2))
        s = pow2_scale(t)
        if s is None:
            st = 3
            break
        A = N * s
        st = 2 if st == 1 else (1 if f >= 0.999999999 * t * t else 0)
    if st != 2:
        return None, nsq
    js = int(np.argmax(np.diag(A).real))
    u = np.conj(A[js, :])
    for _ in range(2):
        w = A0 @ u
        u = w * (1.0 / np.sqrt(np.sum(np.abs(w) ** 2)))
    w = A0 @ u
    lam = float(np.sum(u.real * w.real + u.imag * w.imag))
    res2 = float(np.sum(np.abs(w - lam * u) ** 2))
    if not (res2 <= 1e-24 * lam * lam and lam > 0):
        return None, nsq
    mag = abs(u[0])
    p = np.conj(u[0]) / mag if mag > 0 else 1.0
    e = u * p
    e[0] = e[0].real
    Vt = np.zeros((M, M), np.complex128)
    Vt[M - 1] = e
    h = 1.0 / (1.0 + e[0].real)
    for q in range(1, M):
        g = -e * np.conj(e[q]) * h
        g[q] += 1.0
        g[0] = -np.conj(e[q])
        Vt[q - 1] = g
    return Vt, nsq


if __name__ == "__main__":
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gr_baz_b200 import synth

    for base, snr in ((2, 20.0), (2, 0.0), (2, 40.0), (1, 0.0), (2, -10.0), (2, -20.0)):
        cfg = synth.config(base, snr_db=snr)
        x = synth.gen_windows_numpy(cfg, 123, 0, 200)
        M, N = cfg["m"], cfg["snapshots"]
        errs, orth, nsqs, fails = [], [], [], 0
        for w in range(200):
            X = x[w].reshape(N, M).T.astype(np.complex128)
            R = X @ X.conj().T / N
            ev, V = np.linalg.eigh(R)
            Vt, nsq = principal(R)
            if Vt is None:
                fails += 1
                continue
            e = V[:, 3] * np.conj(V[0, 3]) / abs(V[0, 3])
            errs.append(np.linalg.norm(Vt[3] - e))
            Q = Vt.T  # columns = vectors
            orth.append(np.max(np.abs(Q.conj().T @ Q - np.eye(4))))
            nsqs.append(nsq)
        print("config %d snr %5.1f dB: Jacobi fallbacks %3d/200, max |e - e_lapack| %.2e, max |Q^H Q - I| %.2e, squarings %s"
              % (base, snr, fails, max(errs) if errs else 0, max(orth) if orth else 0, np.bincount(nsqs).tolist() if nsqs else []))
