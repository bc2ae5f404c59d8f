"""CPU tests of the C-ABI boundary: the library loads, exports every symbol that
include/music_b200.h declares, validates arguments like the reference's constructor asserts,
and fails loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from gr_baz_b200 import _capi, build, synth
from gr_baz_b200.music_doa import music_doa
from gr_baz_b200.music_doa_helper import calculate_antenna_array_response, music_doa_helper

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_cuda()
    return _capi.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "music_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(music_b200_\w+)\s*\(", hdr)))
    assert declared == sorted(_capi.EXPORTS)
    raw = ctypes.CDLL(_capi.lib_path())
    for sym in declared:
        assert getattr(raw, sym) is not None
    assert lib.music_b200_version() == 2


def test_create_rejects_bad_parameters(lib):
    # the reference's asserts (lib/baz_music_doa.cc:45-50) are real errors here
    t = np.zeros((360, 4), np.complex64)
    h = ctypes.c_void_p()
    for m, n, ns, res in ((0, 1, 512, 360), (4, 0, 512, 360), (4, 4, 512, 360), (4, 5, 512, 360),
                          (4, 1, 0, 360), (4, 1, 510, 360), (4, 1, 512, 0), (17, 1, 17 * 4, 360)):
        rc = lib.music_b200_create(ctypes.byref(h), m, n, ns, res, t.ctypes.data, 0)
        assert rc == _capi.EINVAL and not h
        assert lib.music_b200_last_error(None)
    assert lib.music_b200_create(ctypes.byref(h), 4, 1, 512, 360, None, 0) == _capi.EINVAL


def test_python_block_raises_like_reference_convention():
    resp = calculate_antenna_array_response([[0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [1.5, 0.0]], 360, 1.0)
    with pytest.raises(ValueError):
        music_doa(4, 4, 512, resp, 360)
    with pytest.raises(ValueError):
        music_doa(4, 1, 512, resp[:-1], 360)  # array_response.size() != resolution
    with pytest.raises(Exception):
        music_doa_helper(4, 1, 510, 360, synth.FREQUENCY, 0.5, synth.antenna_array("ula_x", 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu(lib):
    t = np.zeros((360, 4), np.complex64)
    h = ctypes.c_void_p()
    rc = lib.music_b200_create(ctypes.byref(h), 4, 1, 512, 360, t.ctypes.data, 0)
    assert rc == _capi.ENODEVICE and not h
    resp = calculate_antenna_array_response([[0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [1.5, 0.0]], 360, 1.0)
    with pytest.raises(_capi.MusicB200Error):
        music_doa(4, 1, 512, resp, 360)
