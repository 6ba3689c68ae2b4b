"""GPU parity of the reference's other kNN / grouping entry points (knn, get_graph_feature,
pairwise_dist, square_distance, knn_point, index_points, get_edge_features_xyz) against the golden vectors the
unmodified reference produced and against the C oracle.  Indices: identical wherever distances are strictly
ordered (bit-exact distance arithmetic); distances of xyz rows: bit for bit."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import knn_ref
from oracle import pointnet_ref as P
from test_oracle_graph import same_topk

pytestmark = pytest.mark.gpu


def _pkg():
    import spgan_b200
    return spgan_b200


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_knn_matches_reference(tag):
    g = golden("graph_util")
    x, k, ref = g["knn_%s_x" % tag], int(g["knn_%s_k" % tag]), g["knn_%s_idx" % tag]
    idx = _pkg().knn(_cu(x), k)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == ref.shape
    idx = idx.cpu().numpy()
    assert same_topk(-g["knn_%s_negdist" % tag], idx, ref, ordered=True)
    assert np.array_equal(idx, P.knn(x, k))                 # same (dist, index) tie order as the oracle
    if tag != "b":
        assert np.array_equal(idx, ref)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_get_graph_feature_matches_reference(tag):
    g = golden("graph_util")
    x, k = g["knn_%s_x" % tag], int(g["knn_%s_k" % tag])
    pkg = _pkg()
    out = pkg.get_graph_feature(_cu(x), k=k, idx=_cu(g["knn_%s_idx" % tag]))
    assert np.array_equal(out.cpu().numpy(), g["ggf_%s" % tag])
    if tag == "a":                                          # no exact ties: the built-in neighbour search agrees too
        assert np.array_equal(pkg.get_graph_feature(_cu(x), k=k).cpu().numpy(), g["ggf_a"])


def test_get_graph_feature_backward():
    g = golden("graph_util")
    x, k, idx = g["knn_a_x"], int(g["knn_a_k"]), g["knn_a_idx"]
    B, C, N = x.shape
    w = np.random.default_rng(3).standard_normal((B, 2 * C, N, k)).astype(np.float32)
    xt = _cu(x).requires_grad_()
    (_pkg().get_graph_feature(xt, k=k, idx=_cu(idx)) * _cu(w)).sum().backward()
    xr = torch.from_numpy(x).requires_grad_()                # torch restatement of modules.py:664-678 on the CPU
    it = torch.from_numpy(idx)
    nb = torch.stack([xr[b][:, it[b]] for b in range(B)])
    ctr = xr.unsqueeze(3).expand(B, C, N, k)
    (torch.cat([nb - ctr, ctr], 1) * torch.from_numpy(w)).sum().backward()
    err = (xt.grad.cpu() - xr.grad).abs().max() / xr.grad.abs().max()
    assert float(err) < 1e-5, float(err)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_square_distance_pairwise_dist_knn_point_index_points(tag):
    g = golden("graph_util")
    pkg = _pkg()
    xyz, new, ns = g["pt_%s_xyz" % tag], g["pt_%s_new" % tag], int(g["pt_%s_ns" % tag])
    ref_d = g["pt_%s_sqdist" % tag]
    d = pkg.square_distance(_cu(new), _cu(xyz)).cpu().numpy()
    d2 = pkg.pairwise_dist(_cu(new), _cu(xyz)).cpu().numpy()
    assert np.array_equal(d, d2)
    assert np.array_equal(d, P.square_distance(new, xyz))          # bit-exact against the oracle, any C
    if xyz.shape[2] <= 3:
        assert np.array_equal(d, ref_d)                             # xyz rows: bit for bit against the reference
    else:
        assert np.abs(d - ref_d).max() <= 1e-5 * np.abs(ref_d).max()
    gi = pkg.knn_point(ns, _cu(xyz), _cu(new))
    assert gi.dtype == torch.int64 and tuple(gi.shape) == g["pt_%s_knn_point" % tag].shape
    assert np.array_equal(gi.cpu().numpy(), P.knn_point(ns, xyz, new))
    if xyz.shape[2] <= 3:
        assert same_topk(ref_d, gi.cpu().numpy(), g["pt_%s_knn_point" % tag], ordered=False)
    out = pkg.index_points(_cu(xyz), _cu(g["pt_%s_knn_point" % tag]))
    assert np.array_equal(out.cpu().numpy(), g["pt_%s_index_points" % tag])


def test_strided_inputs_and_shared_cloud():
    pkg = _pkg()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 150)).astype(np.float32)        # channel-first storage, point-major VIEW
    view = _cu(x).permute(0, 2, 1)
    d = pkg.square_distance(view, view).cpu().numpy()
    assert np.array_equal(d, P.square_distance(x.transpose(0, 2, 1), x.transpose(0, 2, 1)))
    gi = pkg.knn_point(7, view, view).cpu().numpy()
    assert np.array_equal(gi, P.knn_point(7, x.transpose(0, 2, 1), x.transpose(0, 2, 1)))


def test_index_points_variants_and_errors():
    pkg = _pkg()
    rng = np.random.default_rng(6)
    pts = rng.standard_normal((3, 40, 8)).astype(np.float32)
    idx = rng.integers(0, 40, (3, 11))
    for dt in (torch.int64, torch.int32):
        out = pkg.index_points(_cu(pts), _cu(idx).to(dt))
        assert np.array_equal(out.cpu().numpy(), P.index_points(pts, idx))
    neg = idx.copy()
    neg[0, 0] = -1                                                  # wraps like fancy indexing
    assert np.array_equal(pkg.index_points(_cu(pts), _cu(neg)).cpu().numpy(), P.index_points(pts, neg))
    bad = idx.copy()
    bad[1, 3] = 40
    with pytest.raises(IndexError):
        pkg.index_points(_cu(pts), _cu(bad))
    with pytest.raises(RuntimeError):
        pkg.index_points(torch.from_numpy(pts), torch.from_numpy(idx))         # CPU tensors: no fallback
    # backward = scatter-add (duplicates accumulate)
    pt = _cu(pts).requires_grad_()
    idx2 = rng.integers(0, 40, (3, 9, 4))
    w = rng.standard_normal((3, 9, 4, 8)).astype(np.float32)
    (pkg.index_points(pt, _cu(idx2)) * _cu(w)).sum().backward()
    ref = np.zeros_like(pts)
    for b in range(3):
        np.add.at(ref[b], idx2[b].reshape(-1), w[b].reshape(-1, 8))
    assert np.abs(pt.grad.cpu().numpy() - ref).max() < 1e-5


def test_get_edge_features_xyz_matches_reference():
    g = golden("graph_util")
    fea, xyz = _pkg().get_edge_features_xyz(_cu(g["efx_x"]), _cu(g["efx_pc"]), int(g["efx_k"]))
    assert np.array_equal(fea.cpu().numpy(), g["efx_fea"])
    assert np.array_equal(xyz.cpu().numpy(), g["efx_xyz"])


def test_forward_only_functions_refuse_gradients():
    pkg = _pkg()
    p = torch.randn(1, 20, 3, device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError):
        pkg.square_distance(p, p)
    with torch.no_grad():
        pkg.square_distance(p, p)


def test_edge_shapes_and_argument_errors():
    pkg = _pkg()
    rng = np.random.default_rng(7)
    # Nq not a multiple of the 64-query tile, Nc not a multiple of the 128-candidate tile, C beyond one staged chunk
    xyz = rng.standard_normal((2, 131, 40)).astype(np.float32)
    new = rng.standard_normal((2, 67, 40)).astype(np.float32)
    assert np.array_equal(pkg.knn_point(31, _cu(xyz), _cu(new)).cpu().numpy(), P.knn_point(31, xyz, new))
    assert np.array_equal(pkg.square_distance(_cu(new), _cu(xyz)).cpu().numpy(), P.square_distance(new, xyz))
    x = rng.standard_normal((1, 5, 9)).astype(np.float32)
    assert np.array_equal(pkg.knn(_cu(x), 9).cpu().numpy(), P.knn(x, 9))         # k == N: every point, self included
    with pytest.raises(Exception):
        pkg.knn(_cu(x), 10)                                                        # k > N
    with pytest.raises(Exception):
        pkg.knn(_cu(rng.standard_normal((1, 4, 64)).astype(np.float32)), 33)       # beyond the 32-lane list


def test_full_size_query_kernel_agrees_with_hot_path_kernel():
    """B=8, N=2048, C=64: ranks 1..k from the general query kernel (query norm first) must equal the fused
    hot-path kernel's neighbour list bit for bit (same arithmetic, different tiling / pipeline), and `knn`
    must return the point itself at rank 0 wherever its distance row has a unique minimum at the diagonal."""
    pkg = _pkg()
    ops = pkg.ops
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(8, 64, 2048, device="cuda", generator=gen)
    k = 10
    hot = ops.knn_indices(x, k)
    xs = ops.sqnorm_bcn(x)
    gen_idx = ops.knn_query(x, xs, x, xs, k, first_rank=1, cand_norm_first=False)
    assert torch.equal(hot, gen_idx)
    full = pkg.knn(x, k + 1)
    self_first = full[:, :, 0] == torch.arange(2048, device="cuda").view(1, -1)
    assert float(self_first.float().mean()) > 0.999
    # on two clouds, the whole list against the C oracle
    xc = x[:2].cpu().numpy()
    assert np.array_equal(full[:2].cpu().numpy(), P.knn(xc, k + 1))
