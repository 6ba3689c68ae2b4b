"""Synthetic inputs of the benchmark workload: unit-sphere template, per-cloud latent,
"chair" clouds.  Host-side numpy; nothing here touches the GPU."""
import math
import os

import numpy as np


def normalize_cloud(pc):
    """Centre and scale to unit max radius (Generation/model.py:46-52)."""
    pc = pc - pc.mean(axis=-2, keepdims=True)
    r = np.sqrt((pc ** 2).sum(axis=-1)).max(axis=-1, keepdims=True)
    return pc / r[..., None]


def fibonacci_sphere(n):
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = math.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)


def sphere_template(n, root=None):
    """The reference's template/balls/<n>.xyz after pc_normalize when the fixture derived from it is
    present (tests/golden/sphere_<n>.npy), else a Fibonacci sphere of the same size."""
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "tests", "golden", "sphere_%d.npy" % n)
    if os.path.exists(path):
        return np.load(path).astype(np.float32), "template/balls/%d.xyz (fixture)" % n
    return normalize_cloud(fibonacci_sphere(n)).astype(np.float32), "fibonacci"


def latent_vectors(rng, B, nz, nv=0.2):
    """One N(0, nv) latent per cloud, [B, 1, nz] (Generation/model.py:128; tiled over N on the device)."""
    return rng.normal(0.0, nv, (B, 1, nz)).astype(np.float32)


def synthetic_chairs(rng, B, N):
    """[B, N, 3] fp32: points sampled uniformly by area from a union of boxes (seat, back, four legs)
    with per-cloud random proportions, normalised and shuffled (stand-in for the ShapeNet chair H5)."""
    out = np.empty((B, N, 3), np.float32)
    for b in range(B):
        w, d = rng.uniform(0.35, 0.5), rng.uniform(0.35, 0.5)
        seat_h, leg_t = rng.uniform(0.35, 0.5), rng.uniform(0.03, 0.06)
        back_h, slab = rng.uniform(0.4, 0.7), rng.uniform(0.04, 0.08)
        boxes = [((-w, seat_h, -d), (w, seat_h + slab, d)),
                 ((-w, seat_h + slab, -d), (w, seat_h + slab + back_h, -d + slab))]
        for sx in (-1, 1):
            for sz in (-1, 1):
                cx, cz = sx * (w - leg_t), sz * (d - leg_t)
                boxes.append(((cx - leg_t, 0.0, cz - leg_t), (cx + leg_t, seat_h, cz + leg_t)))
        lo = np.array([bx[0] for bx in boxes])                    # [nb, 3]
        ext = np.array([bx[1] for bx in boxes]) - lo
        # 6 faces per box: axis a fixed at lo or hi, spanning the other two axes
        f_lo, f_ext, f_axis, f_side, area = [], [], [], [], []
        for i in range(len(boxes)):
            for a in range(3):
                u, v = (a + 1) % 3, (a + 2) % 3
                for side in (0.0, 1.0):
                    f_lo.append(lo[i]); f_ext.append(ext[i]); f_axis.append(a); f_side.append(side)
                    area.append(ext[i][u] * ext[i][v])
        f_lo, f_ext = np.array(f_lo), np.array(f_ext)
        f_axis, f_side, area = np.array(f_axis), np.array(f_side), np.array(area)
        which = rng.choice(len(area), size=N, p=area / area.sum())
        t = rng.uniform(size=(N, 3))
        ax = f_axis[which]
        t[np.arange(N), ax] = f_side[which]
        pts = f_lo[which] + t * f_ext[which]
        pts = normalize_cloud(pts)
        out[b] = pts[rng.permutation(N)]
    return out
