"""GPU: the device-side steering-table builder (music_b200_set_geometry)
against the literal Python helper (/root/reference/python/music_doa_helper.py:29-46 followed by the
complex64 rounding of swig/baz_swig.i:564): the table must be bit-identical, and the block's outputs
after a device retune must equal those after set_array_response() with the helper's table."""
import time

import numpy as np
import pytest

from gr_baz_b200 import synth
from gr_baz_b200._capi import MusicB200Error
from gr_baz_b200.music_doa import music_doa
from gr_baz_b200.music_doa_helper import calculate_antenna_array_response, music_doa_helper
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu

CASES = [  # geometry, m, n, K, wavelength scale
    ("ula_x", 4, 1, 360, 1.0), ("ula_x", 4, 1, 3600, 1.0), ("ula_x", 4, 1, 3600, 1.25), ("ula_y", 4, 2, 720, 0.8),
    ("uca", 8, 1, 7200, 1.0), ("uca", 8, 2, 3600, 0.9), ("uca", 16, 2, 3600, 1.0), ("uca", 16, 2, 3600, 3.7),
    ("ula_x", 6, 1, 1000, 1.0), ("uca", 5, 1, 777, 1.1),
]


def positions(geometry, m):
    return [[synth.SPACING * x, synth.SPACING * y] for x, y in synth.antenna_array(geometry, m)]


@pytest.mark.parametrize("geometry,m,n,K,scale", CASES)
def test_device_table_is_bit_identical_to_python_helper(geometry, m, n, K, scale):
    pos = positions(geometry, m)
    l0 = synth.C_LIGHT / synth.FREQUENCY
    l = l0 * scale
    t_first = calculate_antenna_array_response(pos, K, l0 * 1.5)  # something else to start from
    blk = music_doa(m, n, m * 256, t_first, K)
    t0 = time.perf_counter()
    ref = np.asarray(calculate_antenna_array_response(pos, K, l)).astype(np.complex64)
    t_py = time.perf_counter() - t0
    t0 = time.perf_counter()
    guarded = blk.set_array_geometry(pos, l)
    t_dev = time.perf_counter() - t0
    got = blk.array_response_c64()
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # the guard band is narrow: exact zeros of sin/cos and elements at the origin aside, ~3e-5 of the entries
    assert guarded <= 2 * K + 0.01 * 2 * m * K
    print("steer %s m=%d K=%d: python %.1f ms, device %.3f ms, guarded %d of %d" % (geometry, m, K, 1e3 * t_py, 1e3 * t_dev, guarded, 2 * m * K))


def test_outputs_after_device_retune_equal_set_array_response():
    cfg = synth.config(4, snapshots=512)
    pos = positions(cfg["geometry"], cfg["m"])
    l = synth.C_LIGHT / (synth.FREQUENCY * 0.93)
    x = synth.gen_windows_numpy(cfg, 4242, 0, 64)
    table = calculate_antenna_array_response(pos, cfg["resolution"], l)
    t_first = calculate_antenna_array_response(pos, cfg["resolution"], synth.C_LIGHT / synth.FREQUENCY)
    outs = {}
    for how in ("python", "device"):
        blk = music_doa(cfg["m"], cfg["n"], cfg["nsamples"], t_first, cfg["resolution"])
        if how == "python":
            blk.set_array_response(table)
        else:
            blk.set_array_geometry(pos, l)
        ang = np.zeros((64, cfg["n"]), np.float32)
        lvl = np.zeros((64, cfg["n"]), np.float32)
        spec = np.zeros((64, cfg["resolution"]), np.float32)
        assert blk.work(64, [x], [ang, lvl, spec]) == 64
        outs[how] = (ang, lvl, spec, blk.last_bins().copy())
    for a, b in zip(outs["python"], outs["device"]):
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b)
    ref = co.work_batch(x, cfg["m"], cfg["n"], np.asarray(table).astype(np.complex64))
    assert np.array_equal(outs["device"][3], ref["bins"])


def test_helper_device_table_retune():
    cfg = synth.config(1)
    x = synth.gen_windows_numpy(cfg, 99, 0, 12)
    res = []
    for dev in (False, True):
        hb = music_doa_helper(cfg["m"], cfg["n"], cfg["nsamples"], cfg["resolution"], synth.FREQUENCY, synth.SPACING,
                              cfg["antenna_array"], output_spectrum=False, device_table=dev)
        hb.set_frequency(synth.FREQUENCY * 1.1)
        ang = np.zeros((12, 1), np.float32)
        lvl = np.zeros((12, 1), np.float32)
        assert hb.work(12, [x], [ang, lvl]) == 12
        res.append((ang, lvl, hb.impl.last_bins().copy(), hb.impl.array_response_c64()))
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_set_geometry_argument_errors():
    cfg = synth.config(1)
    pos = positions(cfg["geometry"], cfg["m"])
    l = synth.C_LIGHT / synth.FREQUENCY
    blk = music_doa(cfg["m"], cfg["n"], cfg["nsamples"], calculate_antenna_array_response(pos, cfg["resolution"], l), cfg["resolution"])
    before = blk.array_response_c64()
    for bad_l in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(MusicB200Error):
            blk.set_array_geometry(pos, bad_l)
    with pytest.raises(MusicB200Error):
        blk.set_array_geometry([[0.0, float("nan")]] + pos[1:], l)
    with pytest.raises(ValueError):
        blk.set_array_geometry(pos[:-1], l)
    assert np.array_equal(before, blk.array_response_c64())  # a failed retune leaves the table alone
    # degenerate geometry: every element at the origin -> every entry is exactly (1, 0), all of them guarded
    blk.set_array_geometry([[0.0, 0.0]] * cfg["m"], l)
    t = blk.array_response_c64()
    assert np.all(t.real == 1.0) and np.all(t.imag == 0.0) and not np.any(np.signbit(t.imag))
