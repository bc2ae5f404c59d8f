#!/usr/bin/env python
"""CPU study for the NEXT step of the M = 8 / 16 kernels: the 3xTF32 tensor-core screen of
music_fused.cuh generalised from 8 to 2M columns.  For an M-antenna table it emulates the screen (truncating fp32
accumulation, the pessimistic model of the derivation), measures the worst |d~ - d| / ||a||^2 against the bound
B_M = 2^-15 * M / 4 (the accumulation count grows with M: 6M fp32 additions per component), and counts how many bins per
window would survive into the exact fp64 evaluation on the synthetic streams of BASELINE configs 3 / 4 (n = 1).
Run directly for the numbers quoted in DESIGN.md; tests/test_screen_bound.py imports the emulation for M = 8."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr_baz_b200 import synth  # noqa: E402
from oracle import music_oracle as mo  # noqa: E402


def bound(M):
    return 2.0 ** -15 * M / 4.0


def tf32_trunc(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rna(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32)


def add_trunc32(a, b):
    s = np.asarray(a, np.float64) + np.asarray(b, np.float64)
    r = s.astype(np.float32)
    too_big = np.abs(r.astype(np.float64)) > np.abs(s)
    return np.where(too_big, np.nextafter(r, np.float32(0.0)), r).astype(np.float32)


def screen(a_rows, e, truncate=True):
    """a_rows: (K, M) complex64; e: (M,) complex128 unit vector -> (d~ (K,) float32, fl32(||a||^2))"""
    K, M = a_rows.shape
    A = np.empty((K, 2 * M), np.float32)
    A[:, 0::2], A[:, 1::2] = a_rows.real, a_rows.imag
    ah = tf32_trunc(A)
    al = tf32_trunc(A - ah)
    col_re = np.empty(2 * M, np.float32)
    col_im = np.empty(2 * M, np.float32)
    col_re[0::2], col_re[1::2] = e.real, e.imag
    col_im[0::2], col_im[1::2] = -e.imag, e.real
    out = []
    for col in (col_re, col_im):
        bh = tf32_rna(col)
        bl = tf32_rna(col - bh)
        acc = np.zeros(K, np.float32)
        for x, y in ((al, bh), (ah, bl), (ah, bh)):
            for k in range(2 * M):
                p = (x[:, k].astype(np.float64) * np.float64(y[k])).astype(np.float32)
                acc = add_trunc32(acc, p) if truncate else (acc + p).astype(np.float32)
        out.append(acc)
    cr, ci = out
    na = np.sum(A.astype(np.float64) ** 2, axis=1).astype(np.float32)
    return (na - (cr * cr + ci * ci).astype(np.float32)).astype(np.float32), na


def exact(a_rows, e):
    a = a_rows.astype(np.complex128)
    na = np.sum(np.abs(a) ** 2, axis=1)
    return na - np.abs(a @ np.conj(e)) ** 2, na


def survivors(a_rows, e, B):
    """bins the kernel's two sweeps would hand to the exact evaluation (U from the full table here: the decimated first
    sweep of the kernel gives a slightly larger U, hence a few more)"""
    dt, na = screen(a_rows, e)
    dt, na = dt.astype(np.float64), na.astype(np.float64)
    U = np.min(dt + B * na)
    return int(np.sum(dt - B * na <= U * (1 + 2.0 ** -10)))


def study(cfg_id, windows=24, seed=31):
    cfg = synth.config(cfg_id)
    M, N, K = cfg["m"], cfg["snapshots"], cfg["resolution"]
    arr = mo.scaled_antenna_array(synth.SPACING, cfg["antenna_array"])
    table = mo.steering_table_c64(arr, K, synth.C_LIGHT / synth.FREQUENCY)
    x = synth.gen_windows_numpy(cfg, seed, 0, windows)
    B = bound(M)
    worst, counts = 0.0, []
    for w in range(windows):
        X = x[w].reshape(N, M).T.astype(np.complex128)
        ev, V = np.linalg.eigh(X @ X.conj().T / N)
        e = V[:, -1] * np.exp(-1j * np.angle(V[0, -1]))
        dt, na32 = screen(table, e)
        d, na = exact(table, e)
        worst = max(worst, float(np.max(np.abs(dt.astype(np.float64) - d) / na)))
        counts.append(survivors(table, e, B))
    return {"config": cfg_id, "M": M, "K": K, "B": B, "worst_err_over_na": worst, "margin": B / worst,
            "survivors_mean": float(np.mean(counts)), "survivors_max": int(np.max(counts))}


if __name__ == "__main__":
    for cid in (2, 4, 3):
        r = study(cid)
        print("config %d (M = %d, K = %d): B_M = 2^%.1f, worst |d~ - d| / ||a||^2 = %.2e (%.0fx below B_M), survivors per window: mean %.1f, max %d of %d bins"
              % (r["config"], r["M"], r["K"], np.log2(r["B"]), r["worst_err_over_na"], r["margin"], r["survivors_mean"], r["survivors_max"], r["K"]))
