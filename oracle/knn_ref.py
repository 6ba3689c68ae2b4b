"""ctypes front-end of oracle/knn_recipe.c (TEST INFRASTRUCTURE ONLY).

Restates Generation/modules.py:695-704 (pairwise distance -> full sort -> ranks 1..k)
with the exact fp32 rounding order of the reference's CPU run; see knn_recipe.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_knn.so")
_SRC = os.path.join(_HERE, "knn_recipe.c")
_lib = None


def build(force=False):
    """Compile knn_recipe.c -> liboracle_knn.so (gcc only, no GPU needed)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-march=x86-64-v3",
             "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.spgan_oracle_sqnorm.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        lib.spgan_oracle_sqnorm.restype = None
        lib.spgan_oracle_dist.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        lib.spgan_oracle_dist.restype = None
        lib.spgan_oracle_knn.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ip, fp]
        lib.spgan_oracle_knn.restype = ctypes.c_int
        lib.spgan_oracle_dist2.argtypes = [fp, fp, ctypes.c_int, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, fp]
        lib.spgan_oracle_dist2.restype = None
        lib.spgan_oracle_sqnorm_rows.argtypes = [fp, ctypes.c_int64, ctypes.c_int, fp]
        lib.spgan_oracle_sqnorm_rows.restype = None
        lib.spgan_oracle_topk_rows.argtypes = [fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
        lib.spgan_oracle_topk_rows.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f32(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    assert x.ndim == 3, "expected [B, C, N]"
    return x


def _ptr(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def sqnorm(x):
    """xs[b, n] = sum_c x[b, c, n]^2 in the reference's rounding order (modules.py:697)."""
    x = _f32(x)
    B, C, N = x.shape
    out = np.empty((B, N), np.float32)
    _load().spgan_oracle_sqnorm(_ptr(x, ctypes.c_float), B, C, N, _ptr(out, ctypes.c_float))
    return out


def dist(x):
    """The reference's [B, N, N] `dist` matrix, bit for bit (modules.py:695-699)."""
    x = _f32(x)
    B, C, N = x.shape
    out = np.empty((B, N, N), np.float32)
    _load().spgan_oracle_dist(_ptr(x, ctypes.c_float), B, C, N, _ptr(out, ctypes.c_float))
    return out


def knn(x, k, return_dist=False):
    """idx[b, n, r] = rank r+1 of row n in ascending (dist, j) order (modules.py:702-703)."""
    x = _f32(x)
    B, C, N = x.shape
    idx = np.empty((B, N, k), np.int32)
    kd = np.empty((B, N, k + 1), np.float32) if return_dist else None
    rc = _load().spgan_oracle_knn(_ptr(x, ctypes.c_float), B, C, N, k, _ptr(idx, ctypes.c_int32),
                                  _ptr(kd, ctypes.c_float) if kd is not None else None)
    if rc != 0:
        raise ValueError("spgan_oracle_knn: bad arguments (need 1 <= k < N)")
    return (idx, kd) if return_dist else idx


def dist2(xq, xsq, xc, xsc, cand_norm_first=False):
    """[B, Nq, Nc] distances between two channel-first clouds, (-2 q.c + first norm) + second norm, with the
    caller's squared norms (modules.py:629-646, pointnet_util.py:19-40)."""
    xq, xc = _f32(xq), _f32(xc)
    xsq = np.ascontiguousarray(xsq, np.float32)
    xsc = np.ascontiguousarray(xsc, np.float32)
    B, C, Nq = xq.shape
    Nc = xc.shape[2]
    out = np.empty((B, Nq, Nc), np.float32)
    _load().spgan_oracle_dist2(_ptr(xq, ctypes.c_float), _ptr(xsq, ctypes.c_float), Nq, _ptr(xc, ctypes.c_float),
                               _ptr(xsc, ctypes.c_float), Nc, B, C, int(bool(cand_norm_first)),
                               _ptr(out, ctypes.c_float))
    return out


def sqnorm_rows(p):
    """|p[b, n, :]|^2 of point-major rows [B, N, C], rounded squares added in channel order."""
    p = np.ascontiguousarray(p, np.float32)
    B, N, C = p.shape
    out = np.empty((B, N), np.float32)
    _load().spgan_oracle_sqnorm_rows(_ptr(p, ctypes.c_float), B * N, C, _ptr(out, ctypes.c_float))
    return out


def topk_rows(dist, k, first_rank=0):
    """Ranks first_rank..first_rank+k-1 of every row of dist [..., Nc] in ascending (dist, index) order."""
    dist = np.ascontiguousarray(dist, np.float32)
    Nc = dist.shape[-1]
    R = dist.size // Nc
    idx = np.empty(dist.shape[:-1] + (k,), np.int32)
    if _load().spgan_oracle_topk_rows(_ptr(dist, ctypes.c_float), R, Nc, k, first_rank, _ptr(idx, ctypes.c_int32)):
        raise ValueError("spgan_oracle_topk_rows: bad arguments")
    return idx


def idx_equal_up_to_ties(idx_a, idx_b, kdist):
    """Parity predicate of SURVEY 7.3-A: neighbour lists are identical wherever the fp32
    distances are strictly ordered; any permutation is accepted inside a group of
    bit-identical distances (including across the rank-0 / rank-k cuts).

    idx_a, idx_b: [B, N, k]; kdist: [B, N, k+1] ascending distances of ranks 0..k
    (from knn(..., return_dist=True)).  Returns (ok, n_rows_differing, n_rows_excused).
    """
    idx_a = np.asarray(idx_a).reshape(kdist.shape[0], kdist.shape[1], -1)
    idx_b = np.asarray(idx_b).reshape(idx_a.shape)
    k = idx_a.shape[-1]
    neq = idx_a != idx_b
    rows = np.argwhere(neq.any(-1))
    excused = 0
    ok = True
    for b, n in rows:
        d = kdist[b, n]
        # positions r (1..k) whose distance is unique among ranks 0..k must agree, and a
        # differing position is only excusable if its distance ties with a neighbour rank
        # (or with the cut: rank k may tie with the unseen rank k+1 -> treated as excusable
        # only when d[k] == d[k-1] or the caller passes a longer kdist).
        bad = False
        for r in range(k):
            if idx_a[b, n, r] == idx_b[b, n, r]:
                continue
            rr = r + 1
            tie = (d[rr] == d[rr - 1]) or (rr + 1 <= k and d[rr] == d[rr + 1]) or rr == k
            if not tie:
                bad = True
                break
        if bad:
            ok = False
        else:
            excused += 1
    return ok, int(len(rows)), excused
