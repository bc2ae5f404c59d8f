#!/usr/bin/env python
"""Paper study of an integer-sliced (Ozaki-style) covariance on the int8 tensor cores for M = 16 (round-1 review, item 4
stretch).  The sliced products accumulate EXACTLY in int32, so the only error of the scheme is the fixed-point alignment of
every antenna row to its largest sample in the window: x_r[c] -> round(x_r[c] / max_r * 2^(S-1)) with S = 7 * slices bits.
This script quantises the config-5 stream that way, forms R from the quantised samples in fp64 (= what the sliced MMAs would
deliver) and reports, per S, the error of R, of P(theta) and the number of windows whose peak bins move - i.e. how many 7-bit
slices the parity gate needs - next to the operation counts that decide whether it is worth building."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr_baz_b200 import synth  # noqa: E402
from oracle import music_oracle as mo  # noqa: E402


def spectrum(R, table, n):
    ev, V = np.linalg.eigh(R)
    G = V[:, : R.shape[0] - n]
    return 1.0 / np.sum(np.abs(table.astype(np.complex128).conj() @ G) ** 2, axis=1)


def topn(P, n):
    order = np.lexsort((np.arange(P.size), -P))  # strength descending, bin ascending
    return order[:n]


def study(cfg_id=5, windows=12, seed=77):
    cfg = synth.config(cfg_id)
    M, N, K, n = cfg["m"], cfg["snapshots"], cfg["resolution"], cfg["n"]
    arr = mo.scaled_antenna_array(synth.SPACING, cfg["antenna_array"])
    table = mo.steering_table_c64(arr, K, synth.C_LIGHT / synth.FREQUENCY)
    x = synth.gen_windows_numpy(cfg, seed, 0, windows)
    rows = []
    for slices in (3, 4, 5, 6, 7):
        S = 7 * slices
        eR, eP, moved = 0.0, 0.0, 0
        for w in range(windows):
            X = x[w].reshape(N, M).T.astype(np.complex128)
            R = X @ X.conj().T / N
            P = spectrum(R, table, n)
            scale = np.maximum(np.max(np.abs(X.real), axis=1), np.max(np.abs(X.imag), axis=1))[:, None] / 2.0 ** (S - 1)
            Xq = (np.round(X.real / scale) + 1j * np.round(X.imag / scale)) * scale
            Rq = Xq @ Xq.conj().T / N
            Pq = spectrum(Rq, table, n)
            eR = max(eR, float(np.max(np.abs(Rq - R)) / np.max(np.abs(R))))
            eP = max(eP, float(np.max(np.abs(Pq - P) / P)))
            moved += int(not np.array_equal(topn(P, n), topn(Pq, n)))
        pairs = slices * (slices + 1) // 2
        rows.append((slices, S, eR, eP, moved, pairs))
    return rows, (M, N, K, n, windows)


if __name__ == "__main__":
    rows, (M, N, K, n, windows) = study()
    print("config 5 (M = %d, N = %d, K = %d, n = %d), %d windows" % (M, N, K, n, windows))
    for slices, S, eR, eP, moved, pairs in rows:
        print("  %d slices (%2d bits): max |dR| / max|R| = %.1e, max rel dP = %.1e, windows whose top-%d bins move: %d / %d; slice-pair products kept (i + j < slices): %d"
              % (slices, S, eR, eP, n, moved, windows, pairs))
    real_in = 2 * M * N
    print("  operation counts per window: fp64 path %.2f M DFMA lane-ops = %.0f k cycles/SM at 64/clk;" % (2 * M * M * N / 1e6, 2 * M * M * N / 64 / 1e3))
    print("    slicing %d real inputs into s int8 slices at ~4 ALU ops per slice + 6 for the alignment: s = 6 -> %.2f M lane-ops = %.0f k cycles/SM at 128/clk;"
          % (real_in, real_in * 30 / 1e6, real_in * 30 / 128 / 1e3))
    print("    int8 MACs: (2M)^2 N per slice pair = %.1f M; 21 pairs = %.0f M = %.0f k cycles/SM at a dense 16 k MAC/clk/SM; HBM floor %.1f k cycles/SM"
          % (4 * M * M * N / 1e6, 21 * 4 * M * M * N / 1e6, 21 * 4 * M * M * N / 16384 / 1e3, 8 * M * N / 24.0 / 1e3))
