"""GPU parity of the kNN(+group) kernels against the C oracle and the reference goldens.
Indices are compared bit-exactly (SURVEY 7.3-A tie rule where exact ties exist)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden, pack_like_golden
from oracle import knn_ref

pytestmark = pytest.mark.gpu


def _ops():
    import spgan_b200
    return spgan_b200.ops


def _knn_gpu(x, k, want_ee=False):
    ops = _ops()
    xt = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = ops.knn_indices(xt, k, want_ee=want_ee)
    torch.cuda.synchronize()
    if want_ee:
        return out[0].cpu().numpy(), out[1].cpu().numpy()
    return out.cpu().numpy()


def test_sqnorm_bitwise():
    ops = _ops()
    rng = np.random.default_rng(0)
    for (B, C, N) in [(2, 3, 2048), (3, 64, 256), (1, 128, 100), (2, 17, 33), (1, 300, 64), (1, 6, 5)]:
        x = rng.standard_normal((B, C, N)).astype(np.float32)
        xs = torch.empty((B, N), device="cuda")
        xt = torch.from_numpy(x).cuda()
        ops.L().sqnorm(xt.data_ptr(), B, C, N, -1, xs.data_ptr(), None)
        torch.cuda.synchronize()
        assert np.array_equal(xs.cpu().numpy(), knn_ref.sqnorm(x)), (B, C, N)


def test_sphere_matches_reference_golden(sphere2048):
    x = np.ascontiguousarray(sphere2048.T[None])
    idx = _knn_gpu(np.tile(x, (3, 1, 1)), 10)
    ref = golden("knn_sphere2048")["idx"].astype(np.int32)
    for b in range(3):
        assert np.array_equal(idx[b], ref[0])


def test_config1_matches_reference_golden_and_fused_group():
    g = golden("knn_config1")
    idx, ee = _knn_gpu(g["x"], 8, want_ee=True)
    assert np.array_equal(idx, g["idx"].astype(np.int32))
    # grouped edge features: exact (copies and one fp32 subtraction)
    assert np.array_equal(pack_like_golden(ee), g["ee_sub"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "knn_misc_*.npz")) +
                                        [os.path.join(GOLDEN, "knn_clustered.npz")]))
def test_golden_shapes_up_to_exact_ties(path):
    g = dict(np.load(path))
    k = g["idx"].shape[-1]
    idx = _knn_gpu(g["x"], k)
    oidx, kd = knn_ref.knn(g["x"], k, return_dist=True)
    assert np.array_equal(idx, oidx)                    # kernel == oracle recipe, always exact
    ok, ndiff, nexc = knn_ref.idx_equal_up_to_ties(idx, g["idx"].astype(np.int32), kd)
    assert ok, (ndiff, nexc)


@pytest.mark.parametrize("B,C,N,k", [(2, 3, 1000, 10), (2, 64, 300, 10), (1, 128, 257, 20), (3, 5, 64, 31),
                                     (1, 70, 129, 3), (2, 1, 40, 7), (1, 3, 2, 1)])
def test_random_shapes_bit_exact_vs_oracle(B, C, N, k):
    rng = np.random.default_rng(B * 7 + C * 3 + N)
    x = rng.standard_normal((B, C, N)).astype(np.float32)
    assert np.array_equal(_knn_gpu(x, k), knn_ref.knn(x, k))


def test_duplicate_points_tie_order():
    """Bit-identical distances: the kernel must order by (dist, index) like the oracle."""
    rng = np.random.default_rng(5)
    base = rng.standard_normal((1, 8, 32)).astype(np.float32)
    x = np.concatenate([base, base, base], axis=2)      # every point three times
    assert np.array_equal(_knn_gpu(x, 6), knn_ref.knn(x, 6))


def test_full_size_b64_n2048_c64():
    """BASELINE size: kernel on all 64 clouds, oracle on two of them; size-independent properties
    on the rest (indices in range, no self-duplicates, permutation-equivariance across clouds)."""
    rng = np.random.default_rng(11)
    one = rng.standard_normal((2, 64, 2048)).astype(np.float32)
    x = np.concatenate([one] * 32, axis=0)              # clouds repeat with period 2
    idx = _knn_gpu(x, 10)
    ref = knn_ref.knn(one, 10)
    assert np.array_equal(idx[:2], ref)
    assert np.array_equal(idx[62:], ref)
    assert (idx.reshape(32, 2, 2048, 10) == idx[:2][None]).all()
    assert idx.min() >= 0 and idx.max() < 2048


def test_group_with_given_idx_and_int64_api():
    import spgan_b200
    g = golden("knn_config1")
    x = torch.from_numpy(g["x"]).cuda()
    ee, idx = spgan_b200.get_edge_features(x, 8, return_idx=True)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (4, 256 * 8)
    assert np.array_equal(idx.view(4, 256, 8).cpu().numpy(), g["idx"].astype(np.int64))
    ee2 = spgan_b200.get_edge_features(x, 8, idx=idx)
    assert torch.equal(ee, ee2)
    assert tuple(ee.shape) == (4, 128, 256, 8)


def test_c_abi_rejects_bad_k():
    ops = _ops()
    import spgan_b200._lib as m
    x = torch.zeros((1, 3, 64), device="cuda")
    with pytest.raises(m.SpganError):
        ops.knn_indices(x, 40)
    with pytest.raises(m.SpganError):
        ops.knn_indices(torch.zeros((1, 3, 4), device="cuda"), 4)


# ------------------------------------------------------------------------------------------------------------
# tensor-core filter + exact refine on point-major rows (csrc/knn_tc.cu): bit-identical to the CUDA-core kernel
def _rows_of(x_bcn):
    B, C, N = x_bcn.shape
    return np.ascontiguousarray(x_bcn.transpose(0, 2, 1)).reshape(B * N, C)


def _tc_vs_exact(x_bcn, k, expect_no_fallback):
    import spgan_b200 as pkg
    ops = pkg.ops
    B, C, N = x_bcn.shape
    assert ops.L().knn_rows_workspace(B, (C + 3) // 4 * 4, N, k) > 0, "shape should be inside the tensor-core kernel's envelope"
    got = ops.knn_indices_rows(torch.from_numpy(_rows_of(x_bcn)).cuda(), B, N, k).cpu().numpy()
    fallbacks = int(ops.LAST_KNN_WORKSPACE[1])
    want = ops.knn_indices(torch.from_numpy(x_bcn).cuda(), k).cpu().numpy()
    assert np.array_equal(got, want), "%d of %d entries differ" % ((got != want).sum(), got.size)
    if expect_no_fallback:
        # the threshold is an upper bound (group minima): a rare query collects more than 16 candidates in one half
        # and is ranked by the exact scan instead -- same result, slower
        assert fallbacks <= max(2, B * N // 20000), fallbacks
    return fallbacks


@pytest.mark.parametrize("B,C,N,k", [(4, 64, 2048, 10), (2, 128, 1024, 15), (3, 16, 256, 5), (1, 256, 128, 8),
                                     (2, 100, 384, 10), (64, 64, 2048, 10), (3, 3, 2048, 10), (2, 6, 256, 4)])
def test_tc_filter_refine_matches_exact_kernel(B, C, N, k):
    rng = np.random.default_rng(B * 7 + C + N)
    x = rng.standard_normal((B, C, N)).astype(np.float32)
    _tc_vs_exact(x, k, expect_no_fallback=True)
    if B * N * N * C <= 3e8:                                     # and against the C oracle where it is affordable
        import spgan_b200 as pkg
        got = pkg.ops.knn_indices_rows(torch.from_numpy(_rows_of(x)).cuda(), B, N, k).cpu().numpy()
        assert np.array_equal(got, knn_ref.knn(x, k))


def test_tc_filter_refine_on_clustered_dense_and_degenerate_clouds():
    """The margin only decides how many candidates reach the exact stage, never the result: clustered features with a
    large common offset (cancellation, exact ties), a coarse grid (many equal distances), duplicated points (candidate
    lists overflow: the exact scan takes over), non-finite rows."""
    rng = np.random.default_rng(11)
    clustered = (1.0 + 0.05 * rng.standard_normal((2, 64, 1024))).astype(np.float32)
    _tc_vs_exact(clustered, 10, expect_no_fallback=False)
    grid = (rng.integers(0, 4, (2, 16, 512)) * 0.5).astype(np.float32)
    _tc_vs_exact(grid, 10, expect_no_fallback=False)
    dup = rng.standard_normal((1, 64, 1024)).astype(np.float32)
    dup[0, :, 512:] = dup[0, :, :512]                            # every point twice
    dup[0, :, :40] = dup[0, :, :1]                               # and one point 40 times: lists overflow
    fb = _tc_vs_exact(dup, 10, expect_no_fallback=False)
    assert fb > 0
    bad = rng.standard_normal((2, 64, 256)).astype(np.float32)
    bad[0, :, 5] = np.nan
    bad[1, 7, :] = np.inf
    _tc_vs_exact(bad, 10, expect_no_fallback=False)
    big = (1e3 * rng.standard_normal((1, 64, 512))).astype(np.float32)     # large magnitudes (fp16 range of the split)
    _tc_vs_exact(big, 10, expect_no_fallback=False)


def test_tc_path_on_the_sphere_template(sphere2048):
    """EdgeConv1's graph: the xyz sphere (C = 3, zero-padded to 4 channels for the TMA row pitch) through the
    tensor-core filter + exact refine equals the reference's own list (tests/golden/knn_sphere2048.npz)."""
    import spgan_b200 as pkg
    rows = torch.from_numpy(np.tile(sphere2048[None], (2, 1, 1)).astype(np.float32)).cuda().view(-1, 3)
    got = pkg.ops.knn_indices_rows(rows, 2, 2048, 10).cpu().numpy()
    want = golden("knn_sphere2048")["idx"].astype(np.int32)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[0])
