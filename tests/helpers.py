"""Shared helpers for the test-suite (oracle-side: allowed to import oracle/)."""
import glob
import os

import numpy as np

from gr_baz_b200 import synth
from oracle import music_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_table_cache = {}


def table_for(cfg):
    key = (cfg["geometry"], cfg["m"], cfg["resolution"])
    if key not in _table_cache:
        arr = mo.scaled_antenna_array(synth.SPACING, cfg["antenna_array"])
        _table_cache[key] = mo.steering_table_c64(arr, cfg["resolution"], synth.C_LIGHT / synth.FREQUENCY)
    return _table_cache[key]


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(path):
    """Returns (cfg, seed, table, list of per-window dicts incl. regenerated input)."""
    g = np.load(path)
    cfg = synth.config(int(g["base"]), m=int(g["m"]), n=int(g["n"]), snapshots=int(g["snapshots"]),
                       resolution=int(g["resolution"]), geometry=str(g["geometry"]), snr_db=float(g["snr_db"]))
    seed = int(g["seed"])
    table = g["table"] if "table" in g.files else table_for(cfg)
    wins = []
    for i, w in enumerate(g["windows"]):
        x = g["in_%d" % i] if ("in_%d" % i) in g.files else synth.gen_windows_numpy(cfg, seed, int(w), 1)[0]
        d = {k: g["%s_%d" % (k, i)] for k in ("R", "eigvals", "noise_projector", "P", "spectrum", "bins", "angles", "levels")}
        d["in"] = x
        d["in_sha256"] = str(g["in_sha256_%d" % i])
        d["w"] = int(w)
        wins.append(d)
    return cfg, seed, table, wins, str(g["table_sha256"])


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b)))
