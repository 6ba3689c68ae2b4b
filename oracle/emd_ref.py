"""ctypes front-end of oracle/emd_recipe.c (TEST INFRASTRUCTURE ONLY): the approximate EMD (synchronous auction)
of metrics/emd/emd_cuda.cu:95-226 as called by Common/GAN_metrics.py:375-379, 406-407.  PARITY UNPINNED -- see the
header of emd_recipe.c for why and for the two hardware-dependent choices made explicit."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_emd.so")
_SRC = os.path.join(_HERE, "emd_recipe.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-march=x86-64-v3", "-o", _SO,
                               _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        lib.spgan_oracle_emd.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, fp, ip, ip]
        lib.spgan_oracle_emd.restype = ctypes.c_int
        lib.spgan_oracle_emd_grad.argtypes = [fp, fp, fp, ip, ctypes.c_int, ctypes.c_int, fp]
        lib.spgan_oracle_emd_grad.restype = None
        _lib = lib
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def emd(xyz1, xyz2, eps, iters, return_trace=False):
    """xyz1, xyz2 [B, n, 3] -> (dist [B, n] squared matched distances, assignment [B, n] int32), the outputs of
    emdModule.forward (metrics/emd/emd_module.py:33-60).  Unlike the reference n need not be a multiple of 1024."""
    a = np.ascontiguousarray(xyz1, np.float32)
    b = np.ascontiguousarray(xyz2, np.float32)
    assert a.ndim == 3 and a.shape == b.shape and a.shape[2] == 3
    B, n, _ = a.shape
    dist = np.empty((B, n), np.float32)
    ass = np.empty((B, n), np.int32)
    trace = np.zeros((B, iters), np.int32) if return_trace else None
    rc = _load().spgan_oracle_emd(_p(a, ctypes.c_float), _p(b, ctypes.c_float), B, n, float(eps), int(iters),
                                  _p(dist, ctypes.c_float), _p(ass, ctypes.c_int32),
                                  _p(trace, ctypes.c_int32) if trace is not None else None)
    if rc:
        raise ValueError("spgan_oracle_emd: bad arguments (need n >= 1, iters >= 1)")
    return (dist, ass, trace) if return_trace else (dist, ass)


def emd_grad(xyz1, xyz2, graddist, assignment):
    """Gradient w.r.t. xyz1 of sum(graddist * dist) with the assignment held fixed (emd_cuda.cu:283-300)."""
    a = np.ascontiguousarray(xyz1, np.float32)
    b = np.ascontiguousarray(xyz2, np.float32)
    g = np.ascontiguousarray(graddist, np.float32)
    s = np.ascontiguousarray(assignment, np.int32)
    B, n, _ = a.shape
    out = np.empty_like(a)
    _load().spgan_oracle_emd_grad(_p(a, ctypes.c_float), _p(b, ctypes.c_float), _p(g, ctypes.c_float),
                                  _p(s, ctypes.c_int32), B, n, _p(out, ctypes.c_float))
    return out
