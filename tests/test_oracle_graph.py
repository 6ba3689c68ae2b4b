"""The oracle's restatement of the reference's other kNN / grouping entry points against the golden
vectors the unmodified reference produced (tests/golden/make_golden_graph.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden
from oracle import pointnet_ref as P


def same_topk(dist_rows, idx_a, idx_b, ordered):
    """Two top-k answers over the same distance rows agree up to exact ties: the gathered distances are identical
    (rank by rank when `ordered`, as multisets otherwise) and neither list repeats an index."""
    da = np.take_along_axis(dist_rows, idx_a.astype(np.int64), -1)
    db = np.take_along_axis(dist_rows, idx_b.astype(np.int64), -1)
    if not ordered:
        da, db = np.sort(da, -1), np.sort(db, -1)
    distinct = all(len(set(r)) == len(r) for r in idx_a.reshape(-1, idx_a.shape[-1]))
    return np.array_equal(da, db) and distinct


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_knn_distance_matrix_bit_exact(tag):
    g = golden("graph_util")
    assert np.array_equal(-P.knn_dist(g["knn_%s_x" % tag]), g["knn_%s_negdist" % tag])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_knn_indices(tag):
    g = golden("graph_util")
    x, k, ref = g["knn_%s_x" % tag], int(g["knn_%s_k" % tag]), g["knn_%s_idx" % tag]
    ours = P.knn(x, k)
    assert same_topk(-g["knn_%s_negdist" % tag], ours, ref, ordered=True)
    if tag != "b":                                    # continuous data: no exact ties, indices identical
        assert np.array_equal(ours, ref)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_graph_feature(tag):
    g = golden("graph_util")
    x, k = g["knn_%s_x" % tag], int(g["knn_%s_k" % tag])
    assert np.array_equal(P.get_graph_feature(x, k, idx=g["knn_%s_idx" % tag]), g["ggf_%s" % tag])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_square_distance_and_knn_point(tag):
    g = golden("graph_util")
    xyz, new, ns = g["pt_%s_xyz" % tag], g["pt_%s_new" % tag], int(g["pt_%s_ns" % tag])
    d = P.square_distance(new, xyz)
    assert np.array_equal(g["pt_%s_sqdist" % tag], g["pt_%s_pairwise" % tag])      # same arithmetic in the reference
    if xyz.shape[2] <= 3:
        assert np.array_equal(d, g["pt_%s_sqdist" % tag])                          # xyz rows: bit for bit
    else:
        assert np.abs(d - g["pt_%s_sqdist" % tag]).max() <= 1e-5 * np.abs(d).max()  # wider rows: norm order differs
    assert same_topk(g["pt_%s_sqdist" % tag], P.knn_point(ns, xyz, new), g["pt_%s_knn_point" % tag],
                     ordered=False) or xyz.shape[2] > 3
    assert np.array_equal(P.index_points(xyz, g["pt_%s_knn_point" % tag]), g["pt_%s_index_points" % tag])


def test_get_edge_features_xyz():
    g = golden("graph_util")
    fea, xyz = P.get_edge_features_xyz(g["efx_x"], g["efx_pc"], int(g["efx_k"]))
    assert np.array_equal(fea, g["efx_fea"]) and np.array_equal(xyz, g["efx_xyz"])


# ---- the same restatements against torch CPU itself on fresh random shapes (not only the stored goldens)
@pytest.mark.parametrize("B,C,N", [(1, 3, 64), (2, 6, 33), (1, 64, 96), (2, 17, 100), (1, 128, 40)])
def test_knn_distance_matches_torch_cpu_bitwise(B, C, N):
    import torch
    x = torch.randn(B, C, N, generator=torch.Generator().manual_seed(B * 100 + C + N))
    inner = -2 * torch.matmul(x.transpose(2, 1), x)                  # Generation/modules.py:641-643
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    assert np.array_equal(-P.knn_dist(x.numpy()), pd.numpy())
    k = min(5, N)
    ref = pd.topk(k=k, dim=-1)[1].numpy()
    assert same_topk(-pd.numpy(), P.knn(x.numpy(), k), ref, ordered=True)


@pytest.mark.parametrize("B,N,M", [(1, 50, 70), (2, 33, 33), (1, 128, 5)])
def test_square_distance_xyz_matches_torch_cpu_bitwise(B, N, M):
    import torch
    g = torch.Generator().manual_seed(N * 7 + M)
    src, dst = torch.randn(B, N, 3, generator=g), torch.randn(B, M, 3, generator=g)
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))             # Common/pointnet_util.py:36-39
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    assert np.array_equal(P.square_distance(src.numpy(), dst.numpy()), dist.numpy())
