"""ORACLE (test infrastructure) -- ctypes binding of ``oracle/libstnref.so`` (built from ``stn_ref.c``
by ``oracle/Makefile``).  Parity unpinned; see ``stn_ref_numpy.py``."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libstnref.so")
    src = os.path.join(_HERE, "stn_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libstnref.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        i64, i = ctypes.c_int64, ctypes.c_int
        L.stn_ref_forward.argtypes = [fp, fp, fp, ip, i64, i, i, i, i, i, i]
        L.stn_ref_backward.argtypes = [fp, fp, fp, fp, fp, i64, i, i, i, i, i, i]
        L.stn_ref_max_threads.restype = i
        _LIB = L
    return _LIB


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def max_threads() -> int:
    return int(lib().stn_ref_max_threads())


def forward(U, theta, out_size, want_corners=False, nthreads=0):
    U = np.ascontiguousarray(U, np.float32)
    theta = np.ascontiguousarray(theta, np.float32).reshape(-1, 6)
    B, H, W, C = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    out = np.empty((B, Ho, Wo, C), np.float32)
    corners = np.empty((4, B, Ho * Wo), np.int32) if want_corners else None
    rc = lib().stn_ref_forward(_p(U), _p(theta), _p(out), _p(corners, ctypes.c_int32), B, H, W, C, Ho, Wo, nthreads)
    assert rc == 0, rc
    return (out, corners) if want_corners else out


def backward(U, theta, out_size, gout, need_dU=True, nthreads=0):
    U = np.ascontiguousarray(U, np.float32)
    theta = np.ascontiguousarray(theta, np.float32).reshape(-1, 6)
    gout = np.ascontiguousarray(gout, np.float32)
    B, H, W, C = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    dU = np.empty_like(U) if need_dU else None
    dtheta = np.empty((B, 2, 3), np.float32)
    rc = lib().stn_ref_backward(_p(U), _p(theta), _p(gout), _p(dU), _p(dtheta), B, H, W, C, Ho, Wo, nthreads)
    assert rc == 0, rc
    return dU, dtheta
