"""The reference's other kNN / grouping entry points on the hot path's kernels.

Same names, argument order and result layout as
  Generation/modules.py:629-680, 727-776   pairwise_dist, knn, get_graph_feature, get_edge_features_xyz
  Common/ops.py:129-162                    knn, get_graph_feature (duplicates of the above)
  Common/pointnet_util.py:19-59            square_distance, index_points
  Common/pointnet_util.py / Common/pointconv_util.py:107-118   knn_point
The distance arithmetic is the reference's CPU rounding order (see csrc/graph_util.cu); neighbour lists come out
sorted by (distance, index), which is one valid outcome of the reference's `topk` (whose order among exact ties,
and for `sorted=False`, is unspecified).  Distances and indices carry no gradient (the reference's `topk` indices
don't either; `square_distance` / `pairwise_dist` are forward-only here and raise if a gradient is requested).
"""
import torch

from . import ops


def _no_grad_input(*ts):
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        raise NotImplementedError("spgan_b200: square_distance / pairwise_dist are forward-only; detach the inputs "
                                  "(neighbour lists carry no gradient)")


def _bcn(x):
    """[B, C, N] fp32 CUDA, contiguous."""
    if x.dim() != 3:
        raise ValueError("expected a [B, C, N] tensor, got shape %s" % (tuple(x.shape),))
    return ops._c(x.detach())


def _rows_and_bcn(p):
    """point-major [B, N, C] (any strides) -> (contiguous rows [B, N, C], contiguous channel-first [B, C, N])."""
    if p.dim() != 3:
        raise ValueError("expected a [B, N, C] tensor, got shape %s" % (tuple(p.shape),))
    rows = ops._c(p.detach())
    return rows, ops.contiguous(rows.permute(0, 2, 1))


def knn(x, k):
    """x [B, C, N] -> idx int64 [B, N, k]: the k nearest points of every point, the point itself included
    (modules.py:640-646: `topk` of -|xi - xj|^2)."""
    x = _bcn(x)
    xs = ops.sqnorm_bcn(x)
    return ops.idx_to_int64(ops.knn_query(x, xs, x, xs, k, first_rank=0, cand_norm_first=True))


def get_graph_feature(x, k=20, idx=None):
    """x [B, dims, N] -> [B, 2*dims, N, k] = cat(neighbour - centre, centre) (modules.py:651-680; note the channel
    order is the opposite of get_edge_features).  `idx` [B, N, k] as returned by `knn`."""
    B, dims, N = x.shape
    if idx is None:
        idx = knn(x, k)
    idx32 = ops.idx_to_int32(idx.reshape(B, N, k)) if idx.dtype != torch.int32 else idx.reshape(B, N, k).contiguous()
    return ops.Group.apply(x, idx32, k, True)


def pairwise_dist(x, y):
    """x [B, N, C], y [B, M, C] -> [B, N, M]: -2 x.y + |x|^2 + |y|^2 (modules.py:629-637)."""
    _no_grad_input(x, y)
    xr, xb = _rows_and_bcn(x)
    yr, yb = (xr, xb) if y is x else _rows_and_bcn(y)
    return ops.pairwise_sqdist(xb, ops.sqnorm_rows(xr), yb, ops.sqnorm_rows(yr))


def square_distance(src, dst):
    """src [B, N, C], dst [B, M, C] -> [B, N, M] (pointnet_util.py:19-40; same arithmetic as pairwise_dist)."""
    return pairwise_dist(src, dst)


def knn_point(nsample, xyz, new_xyz):
    """xyz [B, N, C] all points, new_xyz [B, S, C] queries -> group_idx int64 [B, S, nsample]
    (pointconv_util.py:107-118: `topk(square_distance(new_xyz, xyz), nsample, largest=False, sorted=False)`;
    returned in ascending distance order)."""
    cr, cb = _rows_and_bcn(xyz)
    qr, qb = (cr, cb) if new_xyz is xyz else _rows_and_bcn(new_xyz)
    idx = ops.knn_query(qb, ops.sqnorm_rows(qr), cb, ops.sqnorm_rows(cr), nsample, first_rank=0, cand_norm_first=False)
    return ops.idx_to_int64(idx)


def index_points(points, idx):
    """points [B, N, C], idx [B, S] or [B, S, K] (int64 / int32) -> [B, S, C] or [B, S, K, C]
    (pointnet_util.py:43-59).  An out-of-range index raises IndexError like the reference's fancy indexing
    (checked with one host read-back, skipped while a CUDA graph is being captured)."""
    if points.dim() != 3 or idx.dim() < 2:
        raise ValueError("index_points: points [B, N, C] and idx [B, S, ...] expected")
    out, status = ops.GatherRows.apply(points, idx)
    if not torch.cuda.is_current_stream_capturing() and int(status):
        raise IndexError("index_points: index out of range for %d points" % points.shape[1])
    return out.view(*idx.shape, points.shape[2])


def get_edge_features_xyz(x, pc, k, num=-1):
    """x [B, dims, N] features, pc [B, 3, N] coordinates -> (e_fea [B, 2*dims, N, k], e_xyz [B, 6, N, k]): the
    neighbour list of feature space (ranks 1..k, modules.py:741-749) applied to both tensors (modules.py:727-776)."""
    idx32 = ops.knn_indices(x, k)
    return ops.Group.apply(x, idx32, k), ops.Group.apply(pc, idx32, k)
