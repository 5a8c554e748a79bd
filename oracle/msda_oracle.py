"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by ``trackformer_b200``).

ctypes/numpy front-end of ``oracle/msda_oracle.c`` -- the plain-C CPU
restatement of the reference's MSDeformAttn core
(``src/trackformer/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:24-378``,
definition ``ops/functions/ms_deform_attn_func.py:34-54``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.

Parity status: pinned against outputs of the reference's own
``ms_deform_attn_core_pytorch`` (+ autograd) captured in ``tests/golden/`` by
``tests/golden/make_golden.py`` -- the reference ships no golden vectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc -O2 -fopenmp). Building the checker is not using it."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        i = ctypes.c_int
        vp = ctypes.c_void_p
        for sfx in ("f32", "f64"):
            f = getattr(_lib, f"msda_oracle_fwd_{sfx}")
            f.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i]
            f.restype = None
            b = getattr(_lib, f"msda_oracle_bwd_{sfx}")
            b.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i]
            b.restype = None
        _lib.msda_oracle_num_threads.restype = ctypes.c_int
        _lib.msda_oracle_set_num_threads.argtypes = [ctypes.c_int]
    return _lib


def num_threads() -> int:
    return int(_load().msda_oracle_num_threads())


def set_num_threads(n: int) -> None:
    _load().msda_oracle_set_num_threads(int(n))


def _prep(value, shapes, loc, attn):
    value = np.ascontiguousarray(value)
    dt = value.dtype
    if dt not in (np.float32, np.float64):
        raise TypeError(f"oracle supports float32/float64, got {dt}")
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert M2 == M and two == 2 and shapes.shape == (L, 2)
    assert attn.shape == (N, Lq, M, L, P)
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) == S
    return value, shapes, loc, attn, (N, S, M, D, L, Lq, P), ("f32" if dt == np.float32 else "f64")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def msda_forward(value, shapes, loc, attn) -> np.ndarray:
    """out[N, Lq, M*D] for numpy inputs (float32 or float64)."""
    value, shapes, loc, attn, dims, sfx = _prep(value, shapes, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    getattr(_load(), f"msda_oracle_fwd_{sfx}")(_p(value), _p(shapes), _p(loc), _p(attn), _p(out), *dims)
    return out


def msda_backward(value, shapes, loc, attn, grad_out):
    """(grad_value, grad_loc, grad_attn) for numpy inputs."""
    value, shapes, loc, attn, dims, sfx = _prep(value, shapes, loc, attn)
    grad_out = np.ascontiguousarray(grad_out, dtype=value.dtype)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    getattr(_load(), f"msda_oracle_bwd_{sfx}")(
        _p(value), _p(shapes), _p(loc), _p(attn), _p(grad_out), _p(gv), _p(gl), _p(ga), *dims)
    return gv, gl, ga
