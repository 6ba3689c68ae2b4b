"""CPU restatement of the reference's other kNN / grouping entry points.  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Distances come from oracle/knn_recipe.c (the exact fp32 rounding order of the reference's CPU run: FMA chain over
channels, two separately rounded additions of the squared norms); the gather / concat / top-k bookkeeping is
numpy.  Each function cites the reference lines it follows.  Pinned against tests/golden/graph_util.npz, which
tests/golden/make_golden_graph.py produced with the unmodified reference functions.
"""
import numpy as np

from . import knn_ref


def _bcn_of_rows(p):
    return np.ascontiguousarray(np.asarray(p, np.float32).transpose(0, 2, 1))


def knn_dist(x):
    """-pairwise_distance of Generation/modules.py:640-643 (Common/ops.py:129-133): x [B, C, N] -> [B, N, N],
    d[i, j] = (-2 x_i.x_j + |x_j|^2) + |x_i|^2 (candidate norm first), channel-first norm reduction."""
    x = np.ascontiguousarray(x, np.float32)
    xs = knn_ref.sqnorm(x)
    return knn_ref.dist2(x, xs, x, xs, cand_norm_first=True)


def knn(x, k):
    """modules.py:645: topk(k) of the negated distance = the k smallest distances, self included -> [B, N, k]."""
    return knn_ref.topk_rows(knn_dist(x), k, 0)


def get_graph_feature(x, k, idx=None):
    """modules.py:651-680: cat(neighbour - centre, centre) -> [B, 2C, N, k]."""
    x = np.asarray(x, np.float32)
    B, C, N = x.shape
    if idx is None:
        idx = knn(x, k)
    idx = np.asarray(idx).reshape(B, N, k)
    out = np.empty((B, 2 * C, N, k), np.float32)
    for b in range(B):
        nb = x[b][:, idx[b]]                    # [C, N, k]
        ctr = x[b][:, :, None]
        out[b, :C] = nb - ctr
        out[b, C:] = np.broadcast_to(ctr, (C, N, k))
    return out


def square_distance(src, dst):
    """Common/pointnet_util.py:19-40 (= pairwise_dist, modules.py:629-637): src [B, N, C], dst [B, M, C] ->
    [B, N, M] = (-2 src.dst + |src|^2) + |dst|^2, norms reduced over the contiguous last axis."""
    src, dst = np.asarray(src, np.float32), np.asarray(dst, np.float32)
    return knn_ref.dist2(_bcn_of_rows(src), knn_ref.sqnorm_rows(src), _bcn_of_rows(dst), knn_ref.sqnorm_rows(dst),
                         cand_norm_first=False)


pairwise_dist = square_distance


def knn_point(nsample, xyz, new_xyz):
    """Common/pointconv_util.py:107-118: the nsample smallest entries of square_distance(new_xyz, xyz) per query
    (reference order unspecified: sorted=False) -> [B, S, nsample] ascending."""
    return knn_ref.topk_rows(square_distance(new_xyz, xyz), nsample, 0)


def index_points(points, idx):
    """Common/pointnet_util.py:43-59: points [B, N, C], idx [B, S, ...] -> [B, S, ..., C]."""
    points, idx = np.asarray(points), np.asarray(idx)
    return np.stack([points[b][idx[b]] for b in range(points.shape[0])])


def group(x, idx, k):
    """modules.py:706-720: cat(centre, neighbour - centre) for a neighbour list [B, N, k]."""
    x = np.asarray(x, np.float32)
    B, C, N = x.shape
    idx = np.asarray(idx).reshape(B, N, k)
    out = np.empty((B, 2 * C, N, k), np.float32)
    for b in range(B):
        nb = x[b][:, idx[b]]
        ctr = x[b][:, :, None]
        out[b, :C] = np.broadcast_to(ctr, (C, N, k))
        out[b, C:] = nb - ctr
    return out


def get_edge_features_xyz(x, pc, k):
    """modules.py:727-776: ranks 1..k of the feature-space distance rows, applied to features and coordinates."""
    idx = knn_ref.knn(np.ascontiguousarray(x, np.float32), k)
    return group(x, idx, k), group(pc, idx, k)
