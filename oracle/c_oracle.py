"""ctypes binding of the C oracle (oracle/music_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmusic_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with oracle/Makefile (gcc only, seconds)."""
    src = os.path.join(_HERE, "music_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        u = ctypes.c_uint
        L.music_oracle_work.argtypes = [fp, u, u, u, fp, u, fp, fp, fp, ip, dp, dp, dp, dp]
        L.music_oracle_work.restype = ctypes.c_int
        L.music_oracle_work_batch.argtypes = [fp, u, u, u, u, fp, u, fp, fp, fp, ip, dp]
        L.music_oracle_work_batch.restype = ctypes.c_int
        L.music_oracle_herm_eig.argtypes = [u, dp, dp, dp]
        L.music_oracle_herm_eig.restype = ctypes.c_int
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if a is not None else None


def herm_eig(A):
    """Eigen-decomposition of a Hermitian complex128 matrix with the oracle's Jacobi."""
    A = np.array(A, dtype=np.complex128, order="C")
    M = A.shape[0]
    V = np.zeros((M, M), np.complex128)
    w = np.zeros(M, np.float64)
    rc = lib().music_oracle_herm_eig(M, _dp(A.view(np.float64)), _dp(V.view(np.float64)), _dp(w))
    if rc != 0:
        raise RuntimeError("music_oracle_herm_eig failed: %d" % rc)
    return w, V


def work(in_c64, m, n, table_c64, want_spectrum=True, return_internals=False):
    in_c64 = np.ascontiguousarray(in_c64, dtype=np.complex64)
    table_c64 = np.ascontiguousarray(table_c64, dtype=np.complex64)
    K = table_c64.shape[0]
    nsamples = in_c64.shape[0]
    ang = np.zeros(n, np.float32)
    lvl = np.zeros(n, np.float32)
    bins = np.zeros(n, np.int32)
    spec = np.zeros(K, np.float32) if want_spectrum else None
    P = np.zeros(K, np.float64)
    R = np.zeros((m, m), np.complex128) if return_internals else None
    ev = np.zeros(m, np.float64) if return_internals else None
    V = np.zeros((m, m), np.complex128) if return_internals else None
    rc = lib().music_oracle_work(
        _fp(in_c64.view(np.float32)), m, n, nsamples, _fp(table_c64.view(np.float32)), K,
        _fp(ang), _fp(lvl), _fp(spec), _ip(bins), _dp(P),
        _dp(R.view(np.float64)) if R is not None else None, _dp(ev),
        _dp(V.view(np.float64)) if V is not None else None)
    if rc != 0:
        raise ValueError("music_oracle_work failed: %d" % rc)
    res = {"angles": ang, "levels": lvl, "bins": bins, "P": P}
    if want_spectrum:
        res["spectrum"] = spec
    if return_internals:
        G = V[:, : m - n]
        res.update(R=R, eigvals=ev, eigvec=V, noise_projector=G @ G.conj().T)
    return res


def work_batch(in_c64, m, n, table_c64, want_spectrum=False, want_P=True):
    in_c64 = np.ascontiguousarray(in_c64, dtype=np.complex64)
    table_c64 = np.ascontiguousarray(table_c64, dtype=np.complex64)
    W, nsamples = in_c64.shape
    K = table_c64.shape[0]
    ang = np.zeros((W, n), np.float32)
    lvl = np.zeros((W, n), np.float32)
    bins = np.zeros((W, n), np.int32)
    spec = np.zeros((W, K), np.float32) if want_spectrum else None
    P = np.zeros((W, K), np.float64) if want_P else None
    rc = lib().music_oracle_work_batch(
        _fp(in_c64.view(np.float32)), W, m, n, nsamples, _fp(table_c64.view(np.float32)), K,
        _fp(ang), _fp(lvl), _fp(spec), _ip(bins), _dp(P))
    if rc != 0:
        raise ValueError("music_oracle_work_batch failed: %d" % rc)
    return {"angles": ang, "levels": lvl, "bins": bins, "P": P, "spectrum": spec}
