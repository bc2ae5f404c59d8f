"""GPU: the downstream reducers (music_b200_reduce_*) against
oracle/music_oracle.py::reduce_angles / reduce_spectrum.  No reference counterpart (the reference's consumers are
GUI sinks: /root/reference/python/doa_compass_control.py:102-108, /root/reference/python/plot_sink.py:38)."""
import numpy as np
import pytest
import torch

from gr_baz_b200 import synth
from gr_baz_b200._capi import MusicB200Error
from gr_baz_b200.music_doa import music_doa
from oracle import music_oracle as mo

import helpers

pytestmark = pytest.mark.gpu


def circ_close(a, b, tol):
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) % 360.0
    return np.all(np.minimum(d, 360.0 - d) <= tol)


def make_block(n=2, base=4):
    cfg = synth.config(base, snapshots=256, n=n)
    table = helpers.table_for(cfg)
    return cfg, table, music_doa(cfg["m"], n, cfg["nsamples"], table.tolist(), cfg["resolution"])


def test_circular_mean_across_the_wrap_and_unfilled_slots():
    cfg, table, blk = make_block(n=2)
    rng = np.random.default_rng(3)
    W = 5000
    ang = np.empty((W, 2), np.float32)
    ang[:, 0] = (rng.normal(0.0, 3.0, W) % 360.0).astype(np.float32)     # straddles 0 / 360: arithmetic mean ~180, circular ~0
    ang[:, 1] = (rng.normal(200.0, 10.0, W) % 360.0).astype(np.float32)
    lvl = rng.uniform(0.5, 50.0, (W, 2)).astype(np.float32)
    lvl[::7, 1] = 0.0   # unfilled slots are ignored
    ang[::7, 1] = 0.0
    for weighted in (False, True):
        got = blk.reduce_angles(ang, lvl, weighted=weighted)
        ref = mo.reduce_angles(ang, lvl, weighted=weighted)
        assert circ_close(got[0], ref[0], 2e-4)
        assert np.allclose(got[1], ref[1], atol=2e-6) and np.allclose(got[2], ref[2], rtol=1e-6)
    mean, res, wsum = blk.reduce_angles(ang, lvl)
    assert circ_close(mean[0], 0.0, 0.3) and abs(mean[1] - 200.0) < 1.0
    assert wsum[0] == W and wsum[1] == W - len(range(0, W, 7))
    assert res[0] > 0.99 and 0.9 < res[1] < 1.0
    # without levels every window counts, the zeros of the unfilled slots included
    got = blk.reduce_angles(ang)
    ref = mo.reduce_angles(ang)
    assert circ_close(got[0], ref[0], 2e-4) and np.allclose(got[1], ref[1], atol=2e-6) and np.all(got[2] == W)
    # nothing valid -> (0, 0, 0)
    z = blk.reduce_angles(ang[:10], np.zeros((10, 2), np.float32))
    assert np.all(z[0] == 0) and np.all(z[1] == 0) and np.all(z[2] == 0)
    with pytest.raises(MusicB200Error):
        blk.reduce_angles(ang, None, weighted=True)


def test_reducers_on_block_outputs_device_entry():
    """the whole chain on the device: process -> reduce, on config-4-shaped data with one fixed source"""
    cfg = synth.config(4, snapshots=256, fixed_sources=(123.4,))
    table = helpers.table_for(cfg)
    K, W = cfg["resolution"], 600
    blk = music_doa(cfg["m"], 1, cfg["nsamples"], table.tolist(), K)
    dev = torch.device("cuda:0")
    x = torch.from_numpy(synth.gen_windows_numpy(cfg, 99, 0, W).view(np.float32)).to(dev)
    d_ang = torch.zeros(W, 1, dtype=torch.float32, device=dev)
    d_lvl = torch.zeros(W, 1, dtype=torch.float32, device=dev)
    d_spec = torch.zeros(W, K, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    blk.process_device(x.data_ptr(), W, d_ang.data_ptr(), d_lvl.data_ptr(), d_spec.data_ptr(), None, stream=st)
    d_mean = torch.zeros(1, dtype=torch.float32, device=dev)
    d_res = torch.zeros(1, dtype=torch.float32, device=dev)
    d_ms = torch.zeros(K, dtype=torch.float32, device=dev)
    lib = blk._lib
    assert lib.music_b200_reduce_angles_device(blk._h, d_ang.data_ptr(), d_lvl.data_ptr(), W, 0, d_mean.data_ptr(), d_res.data_ptr(), None, st) == 0
    assert lib.music_b200_reduce_spectrum_device(blk._h, d_spec.data_ptr(), W, d_ms.data_ptr(), st) == 0
    torch.cuda.synchronize()
    ang, lvl, spec = d_ang.cpu().numpy(), d_lvl.cpu().numpy(), d_spec.cpu().numpy()
    ref = mo.reduce_angles(ang, lvl)
    assert circ_close(d_mean.cpu().numpy(), ref[0], 2e-4) and np.allclose(d_res.cpu().numpy(), ref[1], atol=2e-6)
    assert abs(float(d_mean.cpu().numpy()[0]) - 123.4) < 1.0 and float(d_res.cpu().numpy()[0]) > 0.99
    ms = d_ms.cpu().numpy()
    assert np.allclose(ms, mo.reduce_spectrum(spec), rtol=2e-6)
    assert abs(int(np.argmax(ms)) * 360.0 / K - 123.4) < 2.5  # the per-window source jitter spans 121.4 .. 125.5 degrees
    # host convenience gives the same numbers
    assert np.array_equal(blk.reduce_spectrum(spec), ms)
    m2, r2, _ = blk.reduce_angles(ang, lvl)
    assert np.array_equal(m2, d_mean.cpu().numpy()) and np.array_equal(r2, d_res.cpu().numpy())
