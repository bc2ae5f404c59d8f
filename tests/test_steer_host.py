"""CPU: the host-libm evaluation used by music_b200_set_geometry() for its guarded entries
(gr-baz_b200/csrc/music_steer.cuh::steer_entry_host) reproduces the literal Python helper
(/root/reference/python/music_doa_helper.py:29-46 + complex64 rounding, swig/baz_swig.i:564) bit for bit."""
import ctypes

import numpy as np
import pytest

from gr_baz_b200 import _capi, synth
from gr_baz_b200.music_doa_helper import calculate_antenna_array_response


def host_table(L, pos, l, K):
    M = pos.shape[0]
    out = np.empty((K, M, 2), np.float32)
    v = (ctypes.c_float * 2)()
    for k in range(K):
        for a in range(M):
            L.music_b200_steer_entry_host(pos.ctypes.data, l, K, k, a, v)
            out[k, a] = v[0], v[1]
    return out


@pytest.mark.parametrize("geometry,m,K,scale", [
    ("ula_x", 4, 360, 1.0), ("ula_x", 4, 3600, 1.37), ("ula_y", 4, 720, 0.61),
    ("uca", 8, 1800, 1.0), ("uca", 16, 900, 2.2),
])
def test_host_entry_matches_python_helper(geometry, m, K, scale):
    L = _capi.load()
    arr = synth.antenna_array(geometry, m)
    pos = np.ascontiguousarray(np.asarray([[synth.SPACING * x, synth.SPACING * y] for x, y in arr], np.float64))
    l = synth.C_LIGHT / synth.FREQUENCY * scale
    ref = np.asarray(calculate_antenna_array_response(pos.tolist(), K, l)).astype(np.complex64)
    got = host_table(L, pos, l, K)
    assert np.array_equal(got.view(np.uint32).reshape(K, m, 2), ref.view(np.uint32).reshape(K, m, 2))
