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


@pytest.mark.parametrize("M,N,K", [(4096, 128, 64), (131072, 256, 128), (3000, 256, 256), (777, 130, 100), (40000, 64, 32)])
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
    assert cs.shape[0] == 4 * min((M + 127) // 128, 148)
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


@pytest.mark.parametrize("M,N,K", [(131072, 1024, 256), (1310720, 64, 128), (1310720, 128, 64), (131072, 256, 128),
                                   (131072, 192, 64), (1310720, 32, 64), (200000, 320, 256)])
def test_fused_pipeline_under_repetition_on_the_step_shapes(M, N, K):
    """The warp-specialised pipeline (TMA producers, converters, MMA issuer, epilogue) on the
    training step's own shapes -- many row tiles per CTA, 1 / 2 / 3 / 5 / 16 column tiles, every K-block count -- launched
    back to back: a protocol slip shows up as a pipeline time-out (trap) or a wrong tile, whatever the timing."""
    ops = _ops()
    A = _rnd(4096, K, seed=5).cuda().repeat(M // 4096 + 1, 1)[:M].contiguous()
    B = _rnd(N, K, seed=6).cuda()
    ref = (A[:4096].double() @ B.double().t()).float()
    out = torch.empty((M, N), device="cuda")
    for _ in range(60):
        assert ops.gemm_fused_raw(A, B, None, tb=True, out=out, wcache=False) is not None
    assert _status(ops) == 0
    for r0 in (0, (M // 4096 - 1) * 4096, (M // 8192) * 4096):
        emax, el2 = rel_err(out[r0:r0 + 4096].cpu().numpy(), ref.cpu().numpy())
        assert emax < 1e-5 and el2 < 2e-6, (r0, emax, el2)
    first = out.clone()
    ops.gemm_fused_raw(A, B, None, tb=True, out=out, wcache=False)
    assert torch.equal(first, out), "same operands, same result bit for bit"


@pytest.mark.parametrize("tb", [True, False])
@pytest.mark.parametrize("M,N,K", [(131072, 256, 1024), (4096, 128, 1280), (1000, 72, 512), (300, 64, 384), (5000, 192, 640),
                                   (129, 1280, 512), (2048, 16, 1536)])
def test_chunked_k_kernel_matches_fp64(M, N, K, tb):
    """K > 256 on engine 3: gemm_tsk_kernel (K walked in chunks of 128 with the accumulators of a column group resident
    in TMEM): fp64 parity incl. bias, C +=, odd column-tile counts and ragged edges; bit-identical repeats."""
    ops = _ops()
    A = _rnd(M, K, seed=41).cuda()
    B = (_rnd(N, K, seed=42) if tb else _rnd(K, N, seed=42)).cuda()
    bias = _rnd(N, seed=43).cuda()
    assert ops.L().gemm_bigk_route(M, N, K, A.data_ptr(), K) == 1
    ref = A.double().cpu() @ (B.t() if tb else B).double().cpu() + bias.double().cpu()
    out = ops.gemm_raw(A, B, bias, False, tb, engine=3)
    assert _status(ops) == 0
    # the tensor core adds into the fp32 TMEM accumulator with truncation: ~1.2e-9 relative per k of the chain
    tol2 = 2e-6 * max(1.0, K / 1024.0)
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    assert emax < 1e-5 and el2 < tol2, (emax, el2)
    out2 = ops.gemm_raw(A, B, bias, False, tb, engine=3)
    assert torch.equal(out, out2)
    ops.gemm_raw(A, B, None, False, tb, out=out2, accumulate=True, engine=3)
    emax, el2 = rel_err(out2.cpu().numpy(), (2 * ref - bias.double().cpu()).numpy())
    assert emax < 1e-5 and el2 < tol2, (emax, el2)
    for _ in range(30):                                   # the pipeline back to back
        ops.gemm_raw(A, B, bias, False, tb, out=out2, engine=3)
    assert _status(ops) == 0 and torch.equal(out, out2)


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


def test_bn_act_linear_chain_matches_unfused_modules():
    """conv -> BN(train) -> LeakyReLU -> conv -> BN -> LeakyReLU -> conv through BnActLinearTrain (statistics from the
    GEMM epilogue, BN + activation in the next GEMM's operand converter) against the same chain on torch modules in
    float64: outputs, input gradient, every parameter gradient, running statistics."""
    import torch.nn as nn
    ops = _ops()
    torch.manual_seed(0)
    R = 20000
    convs = [nn.Linear(16, 64), nn.Linear(64, 128), nn.Linear(128, 96)]
    bns = [nn.BatchNorm1d(64), nn.BatchNorm1d(128)]
    for bn in bns:
        nn.init.uniform_(bn.weight, 0.5, 1.5)
        nn.init.normal_(bn.bias, 0, 0.3)
    x0 = torch.randn(R, 16)
    r = torch.randn(R, 96)
    # reference in float64
    import copy
    cr, br = [copy.deepcopy(c).double() for c in convs], [copy.deepcopy(b).double().train() for b in bns]
    xr = x0.double().requires_grad_()
    h = cr[0](xr)
    h = nn.functional.leaky_relu(br[0](h), 0.01)
    h = cr[1](h)
    h = nn.functional.leaky_relu(br[1](h), 0.01)
    out_ref = cr[2](h)
    (out_ref * r.double()).sum().backward()
    # fused
    cg, bg = [c.cuda() for c in convs], [b.cuda().train() for b in bns]
    for p in list(cg[0].parameters()) + list(cg[1].parameters()) + list(cg[2].parameters()) + list(bg[0].parameters()) + list(bg[1].parameters()):
        p.grad = torch.zeros_like(p)                      # pre-allocated gradients: exercises the in-place accumulation
    xg = x0.cuda().requires_grad_()
    y = ops.linear(xg, cg[0].weight, cg[0].bias, zero_bias_grad=True)
    assert ops.fused_linear_ok(R, cg[1].weight, bg[0], bg[1]) and ops.fused_linear_ok(R, cg[2].weight, bg[1])
    y, st, _ = ops.bn_act_linear(y, ops.bn_train_stats(y, bg[0]), bg[0], 0.01, cg[1].weight, cg[1].bias, next_bn=bg[1],
                                 zero_bias_grad=True)
    out = ops.bn_act_linear(y, st, bg[1], 0.01, cg[2].weight, cg[2].bias)
    ops.MeanScale.apply(ops.Mul.apply(out, r.cuda()), float(r.numel())).backward()
    tol = 1e-3
    assert rel_err(out.detach().cpu().numpy(), out_ref.detach().numpy())[1] < 1e-5
    assert rel_err(xg.grad.cpu().numpy(), xr.grad.numpy())[1] < tol
    for a, b in zip(cg + bg, cr + br):
        for (n1, p1), (_, p2) in zip(a.named_parameters(), b.named_parameters()):
            if isinstance(a, nn.Linear) and n1 == "bias" and a is not cg[2]:
                continue                                   # bias before a train-mode BN: exactly zero, not returned
            assert rel_err(p1.grad.cpu().numpy(), p2.grad.numpy())[1] < tol, (type(a).__name__, n1)
    for a, b in zip(bg, br):
        assert rel_err(a.running_mean.cpu().numpy(), b.running_mean.numpy())[1] < 1e-5
        assert rel_err(a.running_var.cpu().numpy(), b.running_var.numpy())[1] < 1e-5
        assert int(a.num_batches_tracked) == 1


WG_SHAPES = [(1024, 256, 131072), (256, 128, 131072), (128, 1280, 40000), (128, 64, 400000), (64, 32, 50000),
             (72, 500, 9004), (100, 36, 5000), (64, 640, 20000), (16, 1024, 4096), (130, 70, 8200)]


@pytest.mark.parametrize("Mo,No,K", WG_SHAPES)
def test_wgrad_engine3_matches_fp64_and_is_deterministic(Mo, No, K):
    """Weight-gradient form C = A^T B on csrc/gemm_wg.cu (TMA-staged fp32 tiles, column-wise converters, A in TMEM,
    split-K partials added in a fixed order): fp64 parity, bit-identical repeats, C += and strided outputs."""
    ops = _ops()
    A, B = _rnd(K, Mo, seed=11).cuda(), _rnd(K, No, seed=12).cuda()
    ref = A.double().cpu().t() @ B.double().cpu()
    out = ops.gemm_raw(A, B, None, True, False, engine=3)
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    # accumulation chains of 4096 rows in the truncating fp32 TMEM accumulator: ~1.2e-9 relative per k
    assert emax < 4e-5 and el2 < 8e-6, (emax, el2)
    out2 = ops.gemm_raw(A, B, None, True, False, engine=3)
    if Mo % 4 == 0 and No % 4 == 0 and Mo >= 16 and No >= 16:       # inside gemm_wg.cu's envelope: deterministic
        assert torch.equal(out, out2)
    # accumulate into a column block of a wider matrix (the in-place parameter-gradient path of ops.Gemm.backward)
    wide = torch.ones(Mo, No + 8, device="cuda")
    ops.gemm_raw(A, B, None, True, False, out=wide[:, 4:4 + No], accumulate=True, engine=3)
    emax, el2 = rel_err((wide[:, 4:4 + No] - 1).cpu().numpy(), ref.numpy())
    assert emax < 8e-5 and el2 < 1.6e-5, (emax, el2)
    assert float(wide[:, :4].min()) == 1.0 and float(wide[:, 4 + No:].max()) == 1.0


@pytest.mark.parametrize("Mo,No,K", [(1024, 256, 131072), (128, 64, 400000), (256, 128, 40004), (64, 32, 50000),
                                     (72, 100, 9004), (16, 16, 4096)])
def test_wgrad_with_operand_prologue_matches_fp64(Mo, No, K):
    """spgan_gemm_wgrad_fused: C = dY^T lrelu(X * scale + shift), the activated operand formed inside the kernel's
    converter -- against the fp64 product of the explicitly activated operand, and through ops.wgrad_bn_act (both the
    in-kernel prologue and the norm_apply fallback give the same weight gradient)."""
    ops = _ops()
    L = ops.L()
    gz, x = _rnd(K, Mo, seed=21).cuda(), (_rnd(K, No, seed=22) * 1.3 + 0.2).cuda()
    sc, sh = (torch.rand(No) + 0.5).cuda(), (_rnd(No, seed=23) * 0.4).cuda()
    slope = 0.01
    act = torch.nn.functional.leaky_relu(x.double().cpu() * sc.double().cpu() + sh.double().cpu(), slope)
    ref = gz.double().cpu().t() @ act
    ws_bytes = L.gemm_wgrad_workspace(Mo, No, K)
    ws = torch.empty(ws_bytes // 4, device="cuda")
    out = torch.full((Mo, No), 3.0, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.gemm_wgrad_fused(Mo, No, K, gz.data_ptr(), Mo, x.data_ptr(), No, sc.data_ptr(), sh.data_ptr(), slope, out.data_ptr(), No, 0,
                       ws.data_ptr(), ws_bytes, st)
    torch.cuda.synchronize()                     # (a pipeline time-out traps: it would surface here)
    emax, el2 = rel_err(out.cpu().numpy(), ref.numpy())
    assert emax < 4e-5 and el2 < 8e-6, (emax, el2)
    out2 = torch.empty_like(out)
    L.gemm_wgrad_fused(Mo, No, K, gz.data_ptr(), Mo, x.data_ptr(), No, sc.data_ptr(), sh.data_ptr(), slope, out2.data_ptr(), No, 0,
                       ws.data_ptr(), ws_bytes, st)
    assert torch.equal(out, out2), "deterministic"
    L.gemm_wgrad_fused(Mo, No, K, gz.data_ptr(), Mo, x.data_ptr(), No, sc.data_ptr(), sh.data_ptr(), slope, out2.data_ptr(), No, 1,
                       ws.data_ptr(), ws_bytes, st)
    emax, _ = rel_err(out2.cpu().numpy(), 2 * ref.numpy())
    assert emax < 8e-5, emax
    # identity prologue == the plain weight gradient
    L.gemm_wgrad_fused(Mo, No, K, gz.data_ptr(), Mo, x.data_ptr(), No, None, None, 1.0, out2.data_ptr(), No, 0, ws.data_ptr(),
                       ws_bytes, st)
    assert torch.equal(out2, ops.gemm_raw(gz, x, None, True, False, engine=3))


@pytest.mark.parametrize("M,N,K,cols", [(8192, 128, 128, None), (8192, 64, 256, None), (300, 256, 128, None),
                                        (8192, 256, 128, (64, 192)), (100, 3, 64, None)])
def test_act_linear_matches_the_two_op_chain(M, N, K, cols):
    """ops.act_linear: LeakyReLU inside the GEMM's operand converter == LRelu followed by linear, forward and every
    gradient (the weight gradient re-forms the activated operand in gemm_wg's converter); outside the fused kernel's
    envelope (N = 3) it IS the two-op chain."""
    ops = _ops()
    Kw = K if cols is None else 256
    x0, W0, b0, r = _rnd(M, K, seed=31), _rnd(N, Kw, 1, seed=32) * 0.1, _rnd(N, seed=33), _rnd(M, N, seed=34)
    res = []
    for fused in (True, False):
        saved = ops.FUSE_ACT_LINEAR
        ops.FUSE_ACT_LINEAR = fused
        try:
            x = x0.cuda().requires_grad_()
            W = torch.nn.Parameter(W0.cuda())
            b = torch.nn.Parameter(b0.cuda())
            out = ops.act_linear(x, 0.01, W, b, cols=cols)
            ops.MeanScale.apply(ops.Mul.apply(out, r.cuda()), float(r.numel())).backward()
            res.append((out.detach(), x.grad, W.grad, b.grad))
        finally:
            ops.FUSE_ACT_LINEAR = saved
    ref = torch.nn.functional.leaky_relu(x0.double(), 0.01) @ (W0[:, :, 0] if cols is None else W0[:, cols[0]:cols[1], 0]).double().t() + b0.double()
    emax, el2 = rel_err(res[0][0].cpu().numpy(), ref.numpy())
    assert emax < 1e-5 and el2 < 2e-6, (emax, el2)
    for a, b_, what in zip(res[0], res[1], ("out", "dx", "dW", "db")):
        emax, el2 = rel_err(a.cpu().numpy(), b_.cpu().numpy())
        assert emax < 2e-5 and el2 < 1e-5, (what, emax, el2)


def test_split_weight_cache_follows_the_parameter():
    """Inside a weight_cache_scope (the trainer's) the split (fp16 hi / lo) of a Parameter is kept across products until
    the parameter changes: in-place updates bump its version, raw-pointer updates announce themselves through
    ops.weights_changed().  Outside a scope nothing is cached: a bare module stays correct even under writes that bump
    no version counter (`param.data.copy_()`)."""
    import torch.nn as nn
    ops = _ops()
    W = nn.Parameter(_rnd(96, 128, seed=1).cuda())
    b = nn.Parameter(_rnd(96, seed=2).cuda())
    x = _rnd(4096, 128, seed=3).cuda()
    ref = lambda: (x.double() @ W.detach().double().t() + b.detach().double()).cpu().numpy()
    # ---- outside a scope: never cached, `.data` writes are seen
    ops.linear(x, W, b)
    assert len(ops._WCACHE) == 0
    W.data.copy_(_rnd(96, 128, seed=4).cuda())             # no version bump
    assert rel_err(ops.linear(x, W, b).detach().cpu().numpy(), ref())[1] < 2e-6
    with ops.weight_cache_scope():
        y1 = ops.linear(x, W, b)
        n_entries = len(ops._WCACHE)
        assert n_entries >= 1
        y2 = ops.linear(x, W, b)                               # served from the cache
        assert len(ops._WCACHE) == n_entries and torch.equal(y1, y2)
        assert rel_err(y2.detach().cpu().numpy(), ref())[1] < 2e-6
        with torch.no_grad():
            W.mul_(0.5)                                        # version bump: the old split must not be used
        y3 = ops.linear(x, W, b)
        assert rel_err(y3.detach().cpu().numpy(), ref())[1] < 2e-6
        ops.L().axpby(2.0, W.data_ptr(), 0.0, None, W.data_ptr(), W.numel(), ops._stream())     # raw-pointer update
        ops.weights_changed()
        y4 = ops.linear(x, W, b)
        assert rel_err(y4.detach().cpu().numpy(), ref())[1] < 2e-6
        # a non-parameter operand is never cached
        before = len(ops._WCACHE)
        ops.gemm_raw(x, W.detach().clone(), None, False, True, wcache=True)
        assert len(ops._WCACHE) == before
    assert len(ops._WCACHE) == 0                               # leaving the scope drops everything
