"""ORACLE (test infrastructure only). ctypes wrapper around oracle/scan_oracle.c (built by `make -C oracle`)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force: bool = False) -> None:
    names = ["libscan_oracle_f64.so", "libscan_oracle_f32.so"]
    if force or not all(os.path.exists(os.path.join(_HERE, n)) for n in names):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _lib(acc: str):
    if acc not in _LIBS:
        path = os.path.join(_HERE, f"libscan_oracle_{acc}.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.oracle_scan_fwd.argtypes = [fp] * 7 + [ctypes.c_int] * 6 + [fp, fp]
        lib.oracle_scan_fwd.restype = ctypes.c_int
        lib.oracle_scan_bwd.argtypes = [fp] * 8 + [ctypes.c_int] * 6 + [fp] * 7
        lib.oracle_scan_bwd.restype = ctypes.c_int
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIBS[acc] = lib
    return _LIBS[acc]


def num_threads() -> int:
    return int(_lib("f32").oracle_num_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def _c(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def scan_fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, acc="f64"):
    """numpy in / numpy out. B, C: (b, g, n, L). Returns (out, last_state)."""
    u, delta, A, B, C, D, delta_bias = map(_c, (u, delta, A, B, C, D, delta_bias))
    nb, nd, L = u.shape
    G, N = B.shape[1], B.shape[2]
    out = np.empty_like(u)
    last = np.empty((nb, nd, N), np.float32)
    rc = _lib(acc).oracle_scan_fwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias),
                                   int(delta_softplus), nb, nd, L, N, G, _p(out), _p(last))
    if rc:
        raise RuntimeError(f"oracle_scan_fwd rc={rc}")
    return out, last


def scan_bwd(u, delta, A, B, C, D, delta_bias, dout, delta_softplus=False, acc="f64"):
    u, delta, A, B, C, D, delta_bias, dout = map(_c, (u, delta, A, B, C, D, delta_bias, dout))
    nb, nd, L = u.shape
    G, N = B.shape[1], B.shape[2]
    du, ddelta = np.empty_like(u), np.empty_like(u)
    dA = np.empty_like(A)
    dB, dC = np.empty_like(B), np.empty_like(C)
    dD = np.zeros(nd, np.float32) if D is not None else None
    db = np.zeros(nd, np.float32) if delta_bias is not None else None
    rc = _lib(acc).oracle_scan_bwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), _p(dout),
                                   int(delta_softplus), nb, nd, L, N, G,
                                   _p(du), _p(ddelta), _p(dA), _p(dB), _p(dC), _p(dD), _p(db))
    if rc:
        raise RuntimeError(f"oracle_scan_bwd rc={rc}")
    return dict(du=du, ddelta=ddelta, dA=dA, dB=dB, dC=dC, dD=dD, ddelta_bias=db)
