"""ctypes binding for oracle/lloyd_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblloyd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lloyd_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # -march=native is avoided when building for another box: the .so travels to the GPU host.
        subprocess.check_call(
            ["make", "-C", _HERE, "CFLAGS=-O3 -mavx2 -mfma -fopenmp -fPIC -Wall -Wextra"]
            + (["-B"] if force else [])
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_lloyd_step_f32soa.argtypes = [fp, fp, fp, ctypes.c_int64, dp, dp, ctypes.c_int,
                                               ip, dp, dp, dp, ctypes.c_int]
        L.oracle_lloyd_step_f32soa.restype = ctypes.c_int
        L.oracle_lloyd_step_f64.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int, ip, dp, dp, ctypes.c_int]
        L.oracle_lloyd_step_f64.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def lloyd_step_f32soa(x, y, z, centers, mean=None, want_sums=True, want_inertia=False, n_threads=0):
    """E-step (+ sums) on float32 SoA points; centres are float64 in the mean-centred frame."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    z = np.ascontiguousarray(z, dtype=np.float32)
    c = np.ascontiguousarray(centers, dtype=np.float64)
    k = c.shape[0]
    n = x.shape[0]
    m = np.zeros(3) if mean is None else np.ascontiguousarray(mean, dtype=np.float64)
    labels = np.empty(n, dtype=np.int32)
    sums = np.zeros((k, 3))
    counts = np.zeros(k)
    inert = ctypes.c_double(0.0)
    rc = lib().oracle_lloyd_step_f32soa(
        _p(x, ctypes.c_float), _p(y, ctypes.c_float), _p(z, ctypes.c_float), n,
        _p(m, ctypes.c_double), _p(c, ctypes.c_double), k, _p(labels, ctypes.c_int32),
        _p(sums, ctypes.c_double) if want_sums else None,
        _p(counts, ctypes.c_double) if want_sums else None,
        ctypes.byref(inert) if want_inertia else None, n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle_lloyd_step_f32soa failed: {rc}")
    return labels, sums, counts, float(inert.value)


def lloyd_step_f64(X, centers, want_sums=True, n_threads=0):
    X = np.ascontiguousarray(X, dtype=np.float64)
    c = np.ascontiguousarray(centers, dtype=np.float64)
    k = c.shape[0]
    n = X.shape[0]
    labels = np.empty(n, dtype=np.int32)
    sums = np.zeros((k, 3))
    counts = np.zeros(k)
    rc = lib().oracle_lloyd_step_f64(
        _p(X, ctypes.c_double), n, _p(c, ctypes.c_double), k, _p(labels, ctypes.c_int32),
        _p(sums, ctypes.c_double) if want_sums else None,
        _p(counts, ctypes.c_double) if want_sums else None, n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle_lloyd_step_f64 failed: {rc}")
    return labels, sums, counts
