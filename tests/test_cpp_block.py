"""The C++ gr::sync_block front (lib/baz_music_doa.{h,cc}): builds against the GNU Radio compile
shim, exports the reference's factory, and (GPU) produces the oracle's results through work()."""
import os
import struct
import subprocess

import numpy as np
import pytest

from gr_baz_b200 import build, synth
from oracle import c_oracle as co

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lib")


@pytest.fixture(scope="module")
def built():
    build.build_cuda()
    subprocess.check_call(["make", "-C", LIB, "-s"])
    return os.path.join(LIB, "_build")


def test_block_builds_and_exports_reference_factory(built):
    out = subprocess.check_output(["nm", "-DC", os.path.join(built, "libgnuradio-baz-music.so")], text=True)
    assert "baz_make_music_doa(unsigned int, unsigned int, unsigned int" in out
    assert "baz_music_doa::work(int," in out
    assert "baz_music_doa::set_array_response(" in out
    # the block links the CUDA C ABI, nothing from the oracle
    undefined = subprocess.check_output(["nm", "-D", "--undefined-only", os.path.join(built, "libgnuradio-baz-music.so")], text=True)
    assert "music_b200_process_host" in undefined and "music_oracle" not in undefined


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0", "all"])
def test_block_work_matches_oracle(built, tmp_path, devices):
    """devices: BAZ_MUSIC_DOA_DEVICES of the block's environment - None = the single-device handle, "0" = a
    multi-device handle over one GPU, "all" = every GPU of the box (windows dealt round-robin; same outputs)."""
    import torch
    env = dict(os.environ)
    if devices == "all":
        devices = ",".join(str(i) for i in range(torch.cuda.device_count()))
    if devices is not None:
        env["BAZ_MUSIC_DOA_DEVICES"] = devices
    cfg = synth.config(1)
    W = 13
    t1 = helpers.table_for(cfg)
    cfg2 = synth.config(1, geometry="ula_y")
    t2 = helpers.table_for(cfg2)
    x = synth.gen_windows_numpy(cfg, 31, 0, W)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<5I", cfg["m"], cfg["n"], cfg["nsamples"], cfg["resolution"], W))
        f.write(t1.view(np.float32).tobytes())
        f.write(t2.view(np.float32).tobytes())
        f.write(x.view(np.float32).tobytes())
    r = subprocess.run([os.path.join(built, "test_block"), str(fin), str(fout)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "MUSIC DOA: M: 4, N: 1, # samples: 4096, angular resolution: 360" in r.stderr  # banner, reference :52
    assert "Updating array response" in r.stderr  # reference :65
    raw = np.fromfile(fout, dtype=np.float32)
    n, K = cfg["n"], cfg["resolution"]
    o = 0
    ang3 = raw[o:o + W * n].reshape(W, n); o += W * n
    lvl3 = raw[o:o + W * n].reshape(W, n); o += W * n
    spec3 = raw[o:o + W * K].reshape(W, K); o += W * K
    bins = raw[o:o + W * n].view(np.int32).reshape(W, n); o += W * n
    ang1 = raw[o:o + W * n].reshape(W, n); o += W * n
    ang_t2 = raw[o:o + W * n].reshape(W, n); o += W * n
    lvl_t2 = raw[o:o + W * n].reshape(W, n); o += W * n
    assert o == raw.size
    ref = co.work_batch(x, cfg["m"], n, t1, want_spectrum=True)
    ref2 = co.work_batch(x, cfg["m"], n, t2)
    assert np.array_equal(bins, ref["bins"]) and np.array_equal(ang3, ref["angles"]) and np.array_equal(ang1, ref["angles"])
    assert helpers.rel_err(lvl3, ref["levels"]) <= 1e-5 and helpers.rel_err(spec3, ref["P"]) <= 1e-5
    assert np.array_equal(ang_t2, ref2["angles"]) and helpers.rel_err(lvl_t2, ref2["levels"]) <= 1e-5
