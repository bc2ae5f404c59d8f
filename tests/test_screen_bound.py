"""CPU: emulation of the fused kernel's 3xTF32 tensor-core screen (gr-baz_b200/csrc/music_fused.cuh, "Screen error
bound") - table entries split by truncation to TF32 hi/lo, eigenvector entries rounded to fp32 and split with
round-to-nearest-away TF32, the three products a_lo e_hi, a_hi e_lo, a_hi e_hi accumulated in fp32, then
d~ = fl32(||a||^2) - (Re^2 + Im^2) in fp32 - against the exact fp64 value d = ||a||^2 - |e^H a|^2.
The kernel keeps every bin whose lower bound d~ - B ||a||^2 does not exceed the smallest upper bound d~ + B ||a||^2,
B = 2^-15, so |d~ - d| <= B ||a||^2 is what makes the result identical to an all-fp64 scan.  The emulation uses the
PESSIMISTIC accumulator model of the derivation (every fp32 addition truncated toward zero) as well as round to
nearest; the measured worst case must stay under the bound with the margin the source comment claims (> 3x)."""
import numpy as np

from gr_baz_b200 import synth
from oracle import music_oracle as mo

B = 2.0 ** -15


def tf32_trunc(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32)


def add_trunc32(a, b):
    """fp32 addition truncated toward zero (the pessimistic tensor-core accumulator)"""
    s = np.asarray(a, np.float64) + np.asarray(b, np.float64)  # exact for fp32 operands of comparable scale
    r = s.astype(np.float32)
    too_big = np.abs(r.astype(np.float64)) > np.abs(s)
    r = np.where(too_big, np.nextafter(r, np.float32(0.0)), r)
    return r.astype(np.float32)


def screen(a_rows, e, truncate):
    """a_rows: (K, 4) complex64; e: (4,) complex128 unit vector -> d~ (K,) float32"""
    A = np.empty((a_rows.shape[0], 8), np.float32)
    A[:, 0::2], A[:, 1::2] = a_rows.real, a_rows.imag
    ah = tf32_trunc(A)
    al = tf32_trunc(A - ah)  # the subtraction is exact
    col_re = np.empty(8, np.float32)
    col_im = np.empty(8, np.float32)
    col_re[0::2], col_re[1::2] = e.real, e.imag       # Re(e^H a) = sum er*ar + ei*ai
    col_im[0::2], col_im[1::2] = -e.imag, e.real      # Im(e^H a) = sum er*ai - ei*ar
    out = []
    for col in (col_re, col_im):
        bh = tf32_rna(col)
        bl = tf32_rna(col - bh)
        acc = np.zeros(A.shape[0], np.float32)
        for x, y in ((al, bh), (ah, bl), (ah, bh)):  # small terms first, as in the kernel
            for k in range(8):
                p = (x[:, k].astype(np.float64) * np.float64(y[k])).astype(np.float32)  # 11 x 11 bit significands: exact
                acc = add_trunc32(acc, p) if truncate else (acc + p).astype(np.float32)
        out.append(acc)
    cr, ci = out
    na = np.sum(A.astype(np.float64) ** 2, axis=1).astype(np.float32)
    mag = (cr * cr + ci * ci).astype(np.float32)
    return (na - mag).astype(np.float32), na


def exact(a_rows, e):
    a = a_rows.astype(np.complex128)
    na = np.sum(np.abs(a) ** 2, axis=1)
    c = a @ np.conj(e)
    return na - np.abs(c) ** 2, na


def worst_ratio(a_rows, e, truncate):
    dt, na32 = screen(a_rows, e, truncate)
    d, na = exact(a_rows, e)
    return float(np.max(np.abs(dt.astype(np.float64) - d) / na))


def test_screen_error_stays_below_the_bound_on_steering_tables():
    rng = np.random.default_rng(5)
    worst = 0.0
    for geometry, m in (("ula_x", 4), ("ula_y", 4), ("uca", 4)):
        cfg = synth.config(2, geometry=geometry, m=m)
        arr = mo.scaled_antenna_array(synth.SPACING, cfg["antenna_array"])
        table = mo.steering_table_c64(arr, 3600, synth.C_LIGHT / synth.FREQUENCY)
        for trial in range(12):
            if trial % 2 == 0:  # a signal eigenvector that IS (almost) a table row: d ~ 0, the peak, the hard case
                e = table[int(rng.integers(0, 3600))].astype(np.complex128) + 1e-3 * (rng.standard_normal(4) + 1j * rng.standard_normal(4))
            else:
                e = rng.standard_normal(4) + 1j * rng.standard_normal(4)
            e = e / np.linalg.norm(e)
            e = e * np.exp(-1j * np.angle(e[0]))  # component 0 real, as the eigensolver delivers it
            for truncate in (False, True):
                worst = max(worst, worst_ratio(table, e, truncate))
    assert worst <= B / 3.0, "screen error %.3e ||a||^2 exceeds a third of the bound %.3e" % (worst, B)


def test_screen_error_on_adversarial_magnitudes():
    """table rows of very different magnitudes and mixed signs (not steering vectors): the bound is relative to ||a||^2"""
    rng = np.random.default_rng(6)
    worst = 0.0
    for scale in (1e-3, 1.0, 37.0):
        rows = ((rng.standard_normal((2000, 4)) + 1j * rng.standard_normal((2000, 4))) * scale * rng.uniform(0.01, 1.0, (2000, 1))).astype(np.complex64)
        for _ in range(6):
            e = rng.standard_normal(4) + 1j * rng.standard_normal(4)
            e = e / np.linalg.norm(e)
            e = e * np.exp(-1j * np.angle(e[0]))
            worst = max(worst, worst_ratio(rows, e, True))
    assert worst <= B / 3.0, worst


def test_generalised_screen_for_eight_antennas_stays_below_its_bound():
    """the next step for the M = 8 kernel: the same screen with 16 columns; its bound grows with the
    number of fp32 accumulations (B_M = 2^-15 M / 4) and few bins survive it on the config-4 stream"""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import emulate_screen_general as g

    r = g.study(4, windows=6)
    assert r["B"] == 2.0 ** -14
    assert r["worst_err_over_na"] <= r["B"] / 3.0
    assert r["survivors_max"] <= 64
