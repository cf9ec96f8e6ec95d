"""ctypes front-end of the C chamfer oracle (oracle/chamfer_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.

Semantics restated (see the C file for the full citation block):
  /root/reference/chamfer_python.py:18-28   distChamfer output order, lowest-index ties
  /root/reference/global_optimization.py:292-294  ext.chamferDist() [3P, direct difference]
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfpv_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (idempotent)."""
    src = os.path.join(_HERE, "chamfer_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        L.fpvo_nn.argtypes = [f32p, i64, f32p, i64, f32p, i32p]
        L.fpvo_nn.restype = None
        L.fpvo_nn_naive.argtypes = [f32p, i64, f32p, i64, f32p, i32p]
        L.fpvo_nn_naive.restype = None
        L.fpvo_chamfer_fwd.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, f32p, i64p, i64p]
        L.fpvo_chamfer_fwd.restype = None
        L.fpvo_chamfer_bwd.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, f32p, i64p, i64p, f64p, f64p]
        L.fpvo_chamfer_bwd.restype = None
        L.fpvo_pack_key.argtypes = [ctypes.c_float, ctypes.c_uint32]
        L.fpvo_pack_key.restype = ctypes.c_uint64
        L.fpvo_num_threads.argtypes = []
        L.fpvo_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def nn(x, y, naive: bool = False):
    """x [N,3], y [M,3] -> (d [N] f32, idx [N] i32): canonical one-direction NN."""
    x, y = _f32(x), _f32(y)
    assert x.ndim == 2 and y.ndim == 2 and x.shape[1] == 3 and y.shape[1] == 3
    if y.shape[0] == 0:
        raise ValueError("nearest neighbour over an empty candidate set")
    d = np.empty(x.shape[0], np.float32)
    idx = np.empty(x.shape[0], np.int32)
    fn = lib().fpvo_nn_naive if naive else lib().fpvo_nn
    fn(_p(x, ctypes.c_float), x.shape[0], _p(y, ctypes.c_float), y.shape[0], _p(d, ctypes.c_float), _p(idx, ctypes.c_int32))
    return d, idx


def _shared_b(b, bs):
    """b may be [M,3] / [1,M,3] (shared by every batch) or [bs,M,3]."""
    b = _f32(b)
    if b.ndim == 2:
        return b, 0
    if b.shape[0] == 1 and bs > 1:
        return b[0], 0
    return b, b.shape[1] * 3


def dist_chamfer(a, b):
    """Reference return order (chamfer_python.py:28): (d_b2a, d_a2b, i_b2a, i_a2b)."""
    a = _f32(a)
    bs, N, _ = a.shape
    b2, stride = _shared_b(b, bs)
    M = b2.shape[-2]
    if N == 0 or M == 0:
        raise ValueError("distChamfer on an empty cloud")
    d_b2a = np.empty((bs, M), np.float32)
    d_a2b = np.empty((bs, N), np.float32)
    i_b2a = np.empty((bs, M), np.int64)
    i_a2b = np.empty((bs, N), np.int64)
    lib().fpvo_chamfer_fwd(_p(a, ctypes.c_float), _p(b2, ctypes.c_float), bs, N, M, stride,
                           _p(d_b2a, ctypes.c_float), _p(d_a2b, ctypes.c_float),
                           _p(i_b2a, ctypes.c_int64), _p(i_a2b, ctypes.c_int64))
    return d_b2a, d_a2b, i_b2a, i_a2b


def dist_chamfer_bwd(a, b, g_b2a, g_a2b, i_b2a, i_a2b):
    """float64 gradients (grad_a [bs,N,3], grad_b [bs,M,3]) of sum(g_b2a*d_b2a)+sum(g_a2b*d_a2b)."""
    a = _f32(a)
    bs, N, _ = a.shape
    b2, stride = _shared_b(b, bs)
    M = b2.shape[-2]
    g1 = _f32(g_b2a) if g_b2a is not None else None
    g2 = _f32(g_a2b) if g_a2b is not None else None
    i1 = np.ascontiguousarray(i_b2a, dtype=np.int64)
    i2 = np.ascontiguousarray(i_a2b, dtype=np.int64)
    ga = np.empty((bs, N, 3), np.float64)
    gb = np.empty((bs, M, 3), np.float64)
    lib().fpvo_chamfer_bwd(_p(a, ctypes.c_float), _p(b2, ctypes.c_float), bs, N, M, stride,
                           _p(g1, ctypes.c_float), _p(g2, ctypes.c_float),
                           _p(i1, ctypes.c_int64), _p(i2, ctypes.c_int64),
                           _p(ga, ctypes.c_double), _p(gb, ctypes.c_double))
    return ga, gb


def pack_key(d: float, idx: int) -> int:
    return int(lib().fpvo_pack_key(ctypes.c_float(d), ctypes.c_uint32(idx)))


def num_threads() -> int:
    return int(lib().fpvo_num_threads())
