"""spgan_gemm_fused (csrc/gemm_ts.cu): TMA-staged A, operand resident in tensor memory, BatchNorm + LeakyReLU
prologue, column-statistics epilogue -- against fp64 products of the same operands."""
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


def _status(ops):
    torch.cuda.synchronize()
    return int(ops.LAST_TC_WORKSPACE.view(torch.int32)[0])


SHAPES = [(128, 64, 64), (4096, 128, 128), (1000, 72, 132), (2048, 1024, 256), (300, 16, 16), (131072, 128, 64),
          (40000, 512, 200), (129, 70, 36), (70000, 256, 128), (5000, 64, 32), (12800, 128, 192)]


@pytest.mark.parametrize("tb", [True, False])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_fused_plain_matches_fp64(M, N, K, tb):
    ops = _ops()
    A = _rnd(M, K, seed=1)
    B = _rnd(N, K, seed=2) if tb else _rnd(K, N, seed=2)
    bias = _rnd(N, seed=3)
    ref = A.double() @ (B.t() if tb else B).double() + bias.double()
    out = ops.gemm_fused_raw(A.cuda(), B.cuda(), bias.cuda(), tb=tb)
    assert out is not None, "shape should be inside the fused kernel's envelope"
    assert _status(ops) == 0
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    assert emax < 1e-5 and el2 < 2e-6, (emax, el2)


@pytest.mark.parametrize("M,N,K", [(4096, 128, 64), (131072, 256, 128), (3000, 1024, 256), (777, 130, 100)])
def test_fused_prologue_stats_accumulate(M, N, K):
    """pro(A) = lrelu(A * scale + shift); column sums of the output; C += on a second call."""
    ops = _ops()
    A = _rnd(M, K, seed=4)
    B = _rnd(N, K, seed=5)
    bias = _rnd(N, seed=6)
    sc, sh = _rnd(K, seed=7).abs() + 0.5, _rnd(K, seed=8)
    slope = 0.01
    pa = A.double() * sc.double() + sh.double()
    pa = torch.where(pa > 0, pa, pa * slope)
    ref = pa @ B.t().double() + bias.double()
    out, cs, cq = ops.gemm_fused_raw(A.cuda(), B.cuda(), bias.cuda(), tb=True, a_scale=sc.cuda(), a_shift=sh.cuda(),
                                     a_slope=slope, want_stats=True)
    assert _status(ops) == 0
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    assert emax < 1e-5 and el2 < 2e-6, (emax, el2)
    assert cs.shape[0] == 4 * ((M + 127) // 128)
    s1 = cs.double().sum(0).cpu().numpy()
    s2 = cq.double().sum(0).cpu().numpy()
    o64 = out.double().cpu()
    np.testing.assert_allclose(s1, o64.sum(0).numpy(), rtol=1e-5, atol=1e-4 * float(o64.abs().sum(0).max()) / M ** 0.5)
    np.testing.assert_allclose(s2, (o64 * o64).sum(0).numpy(), rtol=1e-5)
    # accumulate: C <- C + product (no prologue), deterministic
    out2 = out.clone()
    ops.gemm_fused_raw(A.cuda(), B.cuda(), None, tb=True, out=out2, accumulate=True)
    ref2 = ref + A.double() @ B.t().double()
    emax, el2 = rel_err(out2.cpu().numpy(), ref2.numpy())
    assert emax < 1e-5 and el2 < 2e-6, (emax, el2)
    out3 = out.clone()
    ops.gemm_fused_raw(A.cuda(), B.cuda(), None, tb=True, out=out3, accumulate=True)
    assert torch.equal(out2, out3)


def test_fused_strided_operands_and_unsupported_shapes():
    ops = _ops()
    big = _rnd(5000, 256, seed=9).cuda()
    A = big[:, 64:192]                                   # lda = 256, 16-byte aligned start
    B = _rnd(96, 128, seed=10).cuda()
    out = ops.gemm_fused_raw(A, B, None, tb=True)
    ref = A.double().cpu() @ B.double().cpu().t()
    assert rel_err(out.cpu().numpy(), ref.numpy())[1] < 2e-6
    assert ops.gemm_fused_raw(_rnd(500, 131, seed=1).cuda(), _rnd(64, 131, seed=2).cuda()) is None     # lda % 4 != 0
    assert ops.gemm_fused_raw(_rnd(500, 512, seed=1).cuda(), _rnd(64, 512, seed=2).cuda()) is None     # K > 256
    assert ops.gemm_fused_raw(_rnd(64, 64, seed=1).cuda(), _rnd(64, 64, seed=2).cuda()) is None        # M < 128
