#!/usr/bin/env python
"""Generates tests/golden/*.npz - OUR golden vectors for the MUSIC DOA path.

The reference ships none, and it cannot be built or imported in this
image, so these fixtures are produced by the numpy/LAPACK restatement
(oracle/music_oracle.py, which follows /root/reference/lib/baz_music_doa.cc:72-161 and
/root/reference/python/music_doa_helper.py:29-46 line by line).  PARITY UNPINNED by the
reference; the fixtures pin *our* oracle (numpy + C), the synthetic generator and the CUDA
path against each other and against drift.

Run from the repo root:   python tests/golden/make_golden.py
Small cases store the input window; large ones store only (config, seed, window index)
plus a sha256 of the regenerated input so that generator drift is detected.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gr_baz_b200 import synth  # noqa: E402
from oracle import music_oracle as mo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def table_for(cfg):
    arr = mo.scaled_antenna_array(synth.SPACING, cfg["antenna_array"])
    return mo.steering_table_c64(arr, cfg["resolution"], synth.C_LIGHT / synth.FREQUENCY)


CASES = [
    # name, cfg overrides, windows, store_input
    ("c1_ula_x_snr20", dict(base=1), [0, 1, 2], True),
    ("c1_ula_x_snr0", dict(base=1, snr_db=0.0), [0, 1], True),
    ("c1_ula_x_snr40", dict(base=1, snr_db=40.0), [0, 1], True),
    ("c1_ula_y_snr20", dict(base=1, geometry="ula_y"), [0], True),
    ("c1_uca4_n2", dict(base=1, geometry="uca", n=2), [0, 1], True),
    ("grc_default", dict(base=1, snapshots=128), [0, 1], True),
    ("c2_ula_x", dict(base=2), [0, 7], False),
    ("c2_ula_x_snr40", dict(base=2, snr_db=40.0), [3], False),
    ("c3_uca8", dict(base=3), [0], False),
    ("c4_uca8", dict(base=4), [0, 5], False),
    ("c4_uca8_snr0", dict(base=4, snr_db=0.0), [1], False),
    ("c5_uca16_n2", dict(base=5), [0, 2], False),
    ("c5_uca16_n2_snr0", dict(base=5, snr_db=0.0), [1], False),
]


def main():
    for name, over, windows, store_input in CASES:
        over = dict(over)
        base = over.pop("base")
        cfg = synth.config(base, **over)
        seed = synth.BASE_SEED + base
        table = table_for(cfg)
        d = dict(
            m=cfg["m"], n=cfg["n"], snapshots=cfg["snapshots"], resolution=cfg["resolution"],
            geometry=cfg["geometry"], snr_db=cfg["snr_db"], base=base, seed=seed,
            windows=np.array(windows), table_sha256=sha(table),
        )
        if cfg["resolution"] <= 360:
            d["table"] = table
        for i, w in enumerate(windows):
            x = synth.gen_windows_numpy(cfg, seed, w, 1)[0]
            r = mo.work(x, cfg["m"], cfg["n"], table, want_spectrum=True, literal_pick=True, return_internals=True)
            d["in_sha256_%d" % i] = sha(x)
            if store_input:
                d["in_%d" % i] = x
            for key in ("R", "eigvals", "noise_projector", "P", "spectrum", "bins", "angles", "levels"):
                d["%s_%d" % (key, i)] = r[key]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "bins", [d["bins_%d" % i].tolist() for i in range(len(windows))])


if __name__ == "__main__":
    main()
