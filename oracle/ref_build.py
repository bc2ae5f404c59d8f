"""TEST INFRASTRUCTURE ONLY - ctypes access to oracle/_ref/libbaz_music_ref.so: the reference's own
/root/reference/lib/baz_music_doa.cc compiled unmodified against stand-in headers (oracle/Makefile target `ref`,
oracle/ref_shim/armadillo, lib/gr_shim/).  Used by tests/test_ref_shim.py to check the oracle restatements against the
reference's own control flow.  It is built only where /root/reference exists (the build container); the prebuilt
.so travels to the GPU box with the repo snapshot.  Not the timed CPU baseline; parity stays formally unpinned."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libbaz_music_ref.so")
REFERENCE = "/root/reference/lib/baz_music_doa.cc"

_lib = None


def build():
    """(Re)build when the reference tree is present; returns True if the library exists afterwards."""
    if os.path.exists(REFERENCE):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(LIB)


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB)
        p = ctypes.c_void_p
        L.ref_music_work_batch.argtypes = [p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, p, ctypes.c_int, p, p, p]
        L.ref_music_work_batch.restype = ctypes.c_int
        _lib = L
    return _lib


def work_batch(in_c64, m, n, table_c64, want_spectrum=True):
    """in_c64 (W, nsamples) complex64, table (K, m) complex64 -> dict of float32 angles/levels (W, n), spectrum (W, K)."""
    x = np.ascontiguousarray(in_c64, dtype=np.complex64)
    t = np.ascontiguousarray(table_c64, dtype=np.complex64)
    W, nsamples = x.shape
    K = t.shape[0]
    ang = np.zeros((W, n), np.float32)
    lvl = np.zeros((W, n), np.float32)
    spec = np.zeros((W, K), np.float32) if want_spectrum else None
    rc = lib().ref_music_work_batch(x.ctypes.data, W, m, n, nsamples, t.ctypes.data, K, ang.ctypes.data, lvl.ctypes.data,
                                    None if spec is None else spec.ctypes.data)
    if rc != 0:
        raise ValueError("ref_music_work_batch failed: %d" % rc)
    return {"angles": ang, "levels": lvl, "spectrum": spec}
