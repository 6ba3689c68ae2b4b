"""tcgen05 engines of spgan_gemm (engine 1: TF32x3 split, engine 2: bf16x3 split; fp32 accumulate in
TMEM) against an fp64 reference and against the fp32 CUDA-core engine; the status word of the
workspace must stay 0 (no pipeline timeout).  Tolerances: 1e-5 relative for TF32x3 (fp32-faithful),
1e-4 for bf16x3 (~2^-16 per product)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    import spgan_b200
    return spgan_b200.ops


def _rnd(*shape, seed=0):
    rng = np.random.default_rng(seed + sum(shape))
    return torch.from_numpy(rng.standard_normal(shape).astype(np.float32))


def _status(ops):
    torch.cuda.synchronize()
    ws = ops.LAST_TC_WORKSPACE
    assert ws is not None, "tcgen05 engine was not selected"
    return int(ws.view(torch.int32)[0])


TOL = {1: 1e-5, 2: 1e-4}


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("tb", [True, False])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1024, 64, 64), (4096, 128, 128), (1000, 70, 131), (777, 256, 80),
                                   (2048, 1024, 256), (300, 16, 16), (5000, 130, 1280), (131072, 128, 64),
                                   (40000, 512, 200)])
def test_tc_gemm_matches_fp64(M, N, K, tb, engine):
    ops = _ops()
    A = _rnd(M, K, seed=1)
    B = _rnd(N, K, seed=2) if tb else _rnd(K, N, seed=2)
    bias = _rnd(N, seed=3)
    ops.LAST_TC_WORKSPACE = None
    out = ops.gemm_raw(A.cuda(), B.cuda(), bias.cuda(), False, tb, engine=engine)
    assert _status(ops) == 0, "tcgen05 pipeline timed out"
    ref = A.double() @ (B.t() if tb else B).double() + bias.double()
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    assert emax < TOL[engine] and el2 < TOL[engine], (emax, el2)
    out2 = ops.gemm_raw(A.cuda(), B.cuda(), None, False, tb, out=out, accumulate=True, engine=engine)
    assert _status(ops) == 0
    emax, el2 = rel_err(out2.cpu().numpy(), (2 * ref - bias.double()).numpy())
    assert emax < TOL[engine] and el2 < TOL[engine], (emax, el2)


@pytest.mark.parametrize("engine", [1, 2])
def test_tc_gemm_strided_and_unaligned(engine):
    ops = _ops()
    X = _rnd(1000, 80, seed=4).cuda()[:, 8:72]               # lda = 80, 32-byte offset
    W = _rnd(96, 134, seed=5).cuda()
    out = ops.gemm_raw(X, W[:, 3:67], None, False, True, engine=engine)        # unaligned weight slice
    assert _status(ops) == 0
    ref = X.cpu().double() @ W.cpu()[:, 3:67].double().t()
    assert max(rel_err(out.cpu().numpy(), ref.numpy())) < 1e-4
    Xu = _rnd(1000, 67, seed=6).cuda()                        # lda % 4 != 0 -> scalar load path
    out = ops.gemm_raw(Xu, W[:, :67], None, False, True, engine=engine)
    assert _status(ops) == 0
    ref = Xu.cpu().double() @ W.cpu()[:, :67].double().t()
    assert max(rel_err(out.cpu().numpy(), ref.numpy())) < 1e-4


def test_tc_engine_agrees_with_cuda_core_engine():
    ops = _ops()
    A, B = _rnd(8192, 128, seed=7).cuda(), _rnd(256, 128, seed=8).cuda()
    a = ops.gemm_raw(A, B, None, False, True, engine=0)
    b = ops.gemm_raw(A, B, None, False, True, engine=1)
    assert _status(ops) == 0
    assert max(rel_err(b.cpu().numpy(), a.cpu().numpy())) < 5e-6
    c = ops.gemm_raw(A, B, None, False, True, engine=2)
    assert _status(ops) == 0
    assert max(rel_err(c.cpu().numpy(), a.cpu().numpy())) < 5e-5


def test_small_shapes_fall_back_to_cuda_cores():
    ops = _ops()
    ops.LAST_TC_WORKSPACE = None
    A, B = _rnd(64, 3, seed=9).cuda(), _rnd(8, 3, seed=10).cuda()
    out = ops.gemm_raw(A, B, None, False, True, engine=1)
    assert ops.LAST_TC_WORKSPACE is None
    assert max(rel_err(out.cpu().numpy(), (A.cpu().double() @ B.cpu().double().t()).numpy())) < 1e-5


@pytest.mark.parametrize("Mo,No,K", [(128, 256, 8192), (1024, 256, 20000), (256, 128, 4096), (256, 128, 131072),
                                     (64, 512, 50000), (128, 1280, 10240), (72, 500, 9001), (64, 32, 50000)])
def test_tc_wgrad_tn_matches_fp64(Mo, No, K):
    """Weight-gradient form C = A^T B (A [K,Mo], B [K,No]) on the tensor cores (MN-major operands,
    atomic flush per 1024-row chunk)."""
    ops = _ops()
    A, B = _rnd(K, Mo, seed=11), _rnd(K, No, seed=12)
    ops.LAST_TC_WORKSPACE = None
    out = ops.gemm_raw(A.cuda(), B.cuda(), None, True, False, engine=1)
    assert _status(ops) == 0, "tcgen05 pipeline timed out"
    ref = A.double().t() @ B.double()
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    assert emax < 1e-5 and el2 < 1e-5, (emax, el2)
    out2 = ops.gemm_raw(A.cuda(), B.cuda(), None, True, False, out=out, accumulate=True, engine=1)
    assert _status(ops) == 0
    emax, el2 = rel_err(out2.cpu().numpy(), (2 * ref).numpy())
    assert emax < 1e-5 and el2 < 1e-5, (emax, el2)


def test_tc_wgrad_strided_operands():
    ops = _ops()
    G = _rnd(8192, 300, seed=13).cuda()[:, 17:273]      # lda = 300, unaligned start
    X = _rnd(8192, 160, seed=14).cuda()[:, 32:160]      # ldb = 160
    out = ops.gemm_raw(G, X, None, True, False, engine=1)
    assert _status(ops) == 0
    ref = G.cpu().double().t() @ X.cpu().double()
    assert max(rel_err(out.cpu().numpy(), ref.numpy())) < 1e-5
