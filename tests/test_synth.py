import numpy as np
import pytest
import torch

from gr_baz_b200 import synth


@pytest.mark.parametrize("cid", [1, 2, 5])
def test_torch_cpu_generator_is_bit_identical_to_numpy(cid):
    cfg = synth.config(cid)
    a = synth.gen_windows_numpy(cfg, synth.BASE_SEED + cid, 5, 3)
    b = synth.gen_windows_torch(cfg, synth.BASE_SEED + cid, 5, 3, "cpu", chunk=2).numpy()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_windows_are_independent_of_batching():
    cfg = synth.config(1)
    a = synth.gen_windows_numpy(cfg, 99, 0, 6)
    b = synth.gen_windows_numpy(cfg, 99, 4, 1)
    assert np.array_equal(a[4], b[0])


def test_signal_statistics():
    cfg = synth.config(2)
    x = synth.gen_windows_numpy(cfg, 1, 0, 2)
    p = np.mean(np.abs(x) ** 2)
    assert abs(p - (1.0 + 10 ** (-cfg["snr_db"] / 10))) < 0.02


@pytest.mark.gpu
def test_device_generator_is_bit_identical_to_numpy():
    for cid in (1, 2, 5):
        cfg = synth.config(cid)
        a = synth.gen_windows_numpy(cfg, synth.BASE_SEED + cid, 11, 3)
        b = synth.gen_windows_torch(cfg, synth.BASE_SEED + cid, 11, 3, "cuda:0").cpu().numpy()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
