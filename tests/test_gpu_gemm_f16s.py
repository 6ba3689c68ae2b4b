"""Engine 3 of spgan_gemm (the FP16S mode of csrc/gemm_tc.cu): fp16 hi + 2^11-scaled fp16 residual, two TMEM
accumulators, kind::f16 MMA rate.  As accurate as the TF32x3 engine (both splits carry 11 + 11 significant bits);
validated on a B200 at the start of round 2 and the default engine since."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    import spgan_b200
    return spgan_b200.ops


def _rnd(*shape, seed=0, scale=1.0):
    rng = np.random.default_rng(seed + sum(shape))
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))


@pytest.mark.parametrize("tb", [True, False])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (4096, 128, 128), (1000, 70, 131), (2048, 1024, 256), (300, 16, 16),
                                   (5000, 130, 1280), (131072, 128, 64), (40000, 512, 200)])
def test_engine3_matches_fp64_like_engine1(M, N, K, tb):
    ops = _ops()
    A = _rnd(M, K, seed=1)
    B = _rnd(N, K, seed=2) if tb else _rnd(K, N, seed=2)
    bias = _rnd(N, seed=3)
    ref = A.double() @ (B.t() if tb else B).double() + bias.double()
    errs = {}
    for engine in (1, 3):
        ops.LAST_TC_WORKSPACE = None
        out = ops.gemm_raw(A.cuda(), B.cuda(), bias.cuda(), False, tb, engine=engine)
        torch.cuda.synchronize()
        assert int(ops.LAST_TC_WORKSPACE.view(torch.int32)[0]) == 0, "tcgen05 pipeline timed out"
        errs[engine] = rel_err(out.cpu().numpy(), ref.numpy())
    assert errs[3][0] < 1e-5 and errs[3][1] < 1e-5, errs
    assert errs[3][1] < 4 * errs[1][1] + 1e-7, errs            # same class of accuracy as TF32x3


def test_engine3_small_and_large_magnitudes():
    """The scaled residual keeps tiny operands exact to ~2^-22 relative (no fp16 subnormal loss) and operands up
    to a few thousand stay inside fp16's range."""
    ops = _ops()
    for sa, sb in [(1e-3, 1e-2), (300.0, 1e-2), (1e-4, 50.0)]:
        A, B = _rnd(2048, 256, seed=7, scale=sa), _rnd(128, 256, seed=8, scale=sb)
        out = ops.gemm_raw(A.cuda(), B.cuda(), None, False, True, engine=3)
        ref = A.double() @ B.t().double()
        emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
        assert el2 < 1e-5, (sa, sb, emax, el2)
