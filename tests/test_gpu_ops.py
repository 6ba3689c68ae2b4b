"""GPU parity of every libspgan_b200 operator (called through the C ABI via spgan_b200.ops)
against plain torch CPU references, forward, backward and -- for the critic's operator set --
double backward.  Tolerance: 1e-3 relative (BASELINE.json), most ops are far tighter."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3
TIGHT = 2e-5


def _ops():
    import spgan_b200
    return spgan_b200.ops


def close(a, b, tol=TIGHT, what=""):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    emax, el2 = rel_err(a, b)
    assert emax <= tol and el2 <= tol, "%s: max %.3e l2 %.3e" % (what, emax, el2)


def rnd(*shape, seed=0, scale=1.0):
    rng = np.random.default_rng(seed + sum(shape))
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))


# ------------------------------------------------------------------------------- gemm
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("M,N,K", [(300, 70, 131), (129, 128, 64), (5, 1, 64), (1000, 3, 64), (257, 200, 3),
                                   (64, 1024, 9000), (128, 1280, 33),
                                   # the streaming kernels for thin products over many rows (csrc/gemm_skinny.cu)
                                   (5000, 64, 3), (8192, 3, 64), (4100, 130, 5), (4096, 4, 256), (6000, 2, 8),
                                   # ... and for the weight gradients of a Conv1d(3, C) (ta: C = A^T B over K rows)
                                   (64, 3, 20000), (200, 4, 16500), (20, 1, 40000)])
def test_gemm_all_layouts(ta, tb, M, N, K):
    ops = _ops()
    A = rnd(K, M, seed=1) if ta else rnd(M, K, seed=1)
    B = rnd(N, K, seed=2) if tb else rnd(K, N, seed=2)
    bias = rnd(N, seed=3)
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double() + bias.double()
    out = ops.gemm_raw(A.cuda(), B.cuda(), bias.cuda(), ta, tb)
    close(out, ref.float(), 1e-5, "gemm")
    out2 = ops.gemm_raw(A.cuda(), B.cuda(), None, ta, tb, out=out, accumulate=True)
    close(out2, (2 * ref - bias.double()).float(), 1e-5, "gemm accumulate")


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False)])
@pytest.mark.parametrize("M,N,K", [(64, 512, 1024), (64, 1, 64), (4, 256, 512), (128, 1024, 512), (64, 64, 256)])
def test_gemm_small_batch_split_k_is_deterministic(M, N, K, ta, tb):
    """Critic MLP shapes (M = batch rows): split-K through private partial tiles, fixed reduction order."""
    ops = _ops()
    A = rnd(K, M, seed=61) if ta else rnd(M, K, seed=61)
    B = rnd(N, K, seed=62) if tb else rnd(K, N, seed=62)
    bias = rnd(N, seed=63)
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double() + bias.double()
    o1 = ops.gemm_raw(A.cuda(), B.cuda(), bias.cuda(), ta, tb)
    o2 = ops.gemm_raw(A.cuda(), B.cuda(), bias.cuda(), ta, tb)
    close(o1, ref.float(), 1e-5, "small-batch gemm")
    assert torch.equal(o1, o2)


def test_gemm_strided_operands_and_unaligned_slices():
    ops = _ops()
    X = rnd(500, 64, seed=4).cuda()
    W = rnd(96, 134, seed=5).cuda()
    out = ops.gemm_raw(X, W[:, 3:67], None, False, True)            # ldb=134, base offset 12 bytes
    close(out, X.cpu().double() @ W.cpu()[:, 3:67].double().t(), 1e-5)
    Xs = rnd(500, 80, seed=6).cuda()[:, 8:72]                         # lda = 80
    out = ops.gemm_raw(Xs, W[:, 64:128], None, False, True)
    close(out, Xs.cpu().double() @ W.cpu()[:, 64:128].double().t(), 1e-5)


def test_gemm_autograd_first_and_second_order():
    ops = _ops()
    X0, W0, b0, r = rnd(200, 48, seed=7), rnd(40, 48, seed=8), rnd(40, seed=9), rnd(200, 40, seed=10)
    # reference: L = sum((dY/dX)^2) with Y = sum(tanh-free) -> use quadratic form to make it depend on W
    Xc, Wc, bc = X0.double().requires_grad_(), W0.double().requires_grad_(), b0.double().requires_grad_()
    Yc = Xc @ Wc.t() + bc
    (gXc,) = torch.autograd.grad((Yc * Yc * r.double()).sum(), Xc, create_graph=True)
    Lc = (gXc * gXc).mean()
    Lc.backward()
    X, W, b = X0.cuda().requires_grad_(), W0.cuda().requires_grad_(), b0.cuda().requires_grad_()
    Y = ops.linear(X, W, b)
    YY = ops.Mul.apply(ops.Mul.apply(Y, Y), r.cuda())
    (gX,) = torch.autograd.grad(ops.MeanScale.apply(YY, float(YY.numel())), X, create_graph=True)
    close(gX, gXc.float(), 1e-4, "first-order dX")
    Lg = ops.MeanScale.apply(ops.Mul.apply(gX, gX), 1.0)
    Lg.backward()
    close(Lg, Lc.float(), 1e-4, "L")
    close(X.grad, Xc.grad.float(), 1e-4, "ddX")
    close(W.grad, Wc.grad.float(), 1e-4, "ddW")
    close(b.grad, bc.grad.float(), 1e-4, "ddb")


# ------------------------------------------------------------------------------- reductions / broadcasts
@pytest.mark.parametrize("R,C,seg", [(4096, 64, 4096), (4096, 64, 256), (777, 3, 777), (60, 130, 20), (8, 1, 8)])
def test_colsum_bcast_addsegvec(R, C, seg):
    ops = _ops()
    x0, v0 = rnd(R, C, seed=11), rnd(R // seg, C, seed=12)
    x = x0.cuda().requires_grad_()
    s = ops.ColSum.apply(x, seg)
    close(s, x0.double().view(R // seg, seg, C).sum(1).float(), TIGHT)
    v = v0.cuda().requires_grad_()
    y = ops.AddSegVec.apply(x, v, seg)
    close(y, x0 + v0.repeat_interleave(seg, 0), TIGHT)
    w = rnd(R, C, seed=13)
    ops.MeanScale.apply(ops.Mul.apply(y, w.cuda()), float(R * C)).backward()
    close(x.grad, w, TIGHT)
    close(v.grad, w.double().view(R // seg, seg, C).sum(1).float(), TIGHT)
    b = ops.BcastSeg.apply(v0.cuda(), R, seg)
    close(b, v0.repeat_interleave(seg, 0), 0.0)


@pytest.mark.parametrize("R,C,seg", [(8192, 64, 8192), (2048, 128, 256), (100, 3, 100), (33, 1024, 33)])
def test_colstats(R, C, seg):
    ops = _ops()
    x0 = rnd(R, C, seed=14) * 0.3 + 5.0                     # large mean: exercises the shifted moments
    mean, rstd, var = ops.col_stats(x0.cuda(), seg, 1e-5)
    xs = x0.double().view(R // seg, seg, C)
    close(mean, xs.mean(1).float(), 1e-6)
    close(var, xs.var(1, unbiased=False).float(), 1e-4)
    close(rstd, (1 / torch.sqrt(xs.var(1, unbiased=False) + 1e-5)).float(), 1e-4)


# ------------------------------------------------------------------------------- batch norm
def _bn_ref(x, g, b):
    return F.batch_norm(x, None, None, g, b, True, 0.1, 1e-5)


@pytest.mark.parametrize("R,C", [(4096, 64), (1000, 3), (64, 512), (5000, 130)])
def test_batchnorm_train_first_and_second_order(R, C):
    ops = _ops()
    x0, g0, b0, r, r2 = rnd(R, C, seed=15), 1 + 0.2 * rnd(C, seed=16), rnd(C, seed=17), rnd(R, C, seed=18), rnd(R, C, seed=19)
    xc, gc, bc = x0.double().requires_grad_(), g0.double().requires_grad_(), b0.double().requires_grad_()
    yc = _bn_ref(xc, gc, bc)
    (gxc,) = torch.autograd.grad((yc * r.double()).sum(), xc, create_graph=True)
    Lc = (gxc * gxc * r2.double()).sum()
    Lc.backward()

    x, g, b = x0.cuda().requires_grad_(), g0.cuda().requires_grad_(), b0.cuda().requires_grad_()
    y, mean, var = ops.BatchNormTrain.apply(x, g, b, 1e-5)
    close(y, yc.float(), 1e-5, "y")
    close(mean.view(-1), x0.double().mean(0).float(), 1e-5, "mean")
    close(var.view(-1), x0.double().var(0, unbiased=False).float(), 1e-4, "var")
    (gx,) = torch.autograd.grad(ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(R * C)), x, create_graph=True)
    close(gx, gxc.float(), 1e-4, "dx")
    Lg = ops.MeanScale.apply(ops.Mul.apply(ops.Mul.apply(gx, gx), r2.cuda()), float(R * C))
    Lg.backward()
    close(Lg, Lc.float(), 1e-4, "L")
    close(x.grad, xc.grad.float(), 5e-4, "ddx")
    close(g.grad, gc.grad.float(), 5e-4, "ddgamma")
    assert b.grad is None or float(b.grad.abs().max()) == 0.0

    # plain first-order parameter grads
    x, g, b = x0.cuda().requires_grad_(), g0.cuda().requires_grad_(), b0.cuda().requires_grad_()
    y, _, _ = ops.BatchNormTrain.apply(x, g, b, 1e-5)
    ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(R * C)).backward()
    xc.grad = gc.grad = bc.grad = None
    (_bn_ref(xc, gc, bc) * r.double()).sum().backward()
    close(x.grad, xc.grad.float(), 1e-4, "dx1")
    close(g.grad, gc.grad.float(), 1e-4, "dgamma1")
    close(b.grad, bc.grad.float(), 1e-4, "dbeta1")


@pytest.mark.parametrize("R,C,slope", [(4096, 64, 0.01), (3000, 132, 0.2), (64, 512, 0.01)])
def test_batchnorm_act_fused_first_and_second_order(R, C, slope):
    """Twice-differentiable fused BN + LeakyReLU (the critic under the gradient penalty) against torch CPU double
    precision: y, d/dx with create_graph, and the gradients of a function of that gradient w.r.t. x, gamma and
    the upstream cotangent."""
    ops = _ops()
    x0, g0, b0 = rnd(R, C, seed=25), 1 + 0.2 * rnd(C, seed=26), rnd(C, seed=27)
    r, r2 = rnd(R, C, seed=28), rnd(R, C, seed=29)
    xc, gc, bc, rc = (t.double().requires_grad_() for t in (x0, g0, b0, r))
    yc = F.leaky_relu(_bn_ref(xc, gc, bc), slope)
    (gxc,) = torch.autograd.grad((yc * rc).sum(), xc, create_graph=True)
    Lc = (gxc * gxc * r2.double()).sum()
    Lc.backward()

    x, g, b, rr = (t.cuda().requires_grad_() for t in (x0, g0, b0, r))
    y, mean, var = ops.BatchNormActTrain2.apply(x, g, b, 1e-5, slope)
    close(y, yc.float(), 1e-5, "y")
    (gx,) = torch.autograd.grad(ops.MeanScale.apply(ops.Mul.apply(y, rr), float(R * C)), x, create_graph=True)
    close(gx, gxc.float(), 1e-4, "dx")
    Lg = ops.MeanScale.apply(ops.Mul.apply(ops.Mul.apply(gx, gx), r2.cuda()), float(R * C))
    Lg.backward()
    close(Lg, Lc.float(), 1e-4, "L")
    close(x.grad, xc.grad.float(), 5e-4, "ddx")
    close(g.grad, gc.grad.float(), 5e-4, "ddgamma")
    close(rr.grad, rc.grad.float(), 5e-4, "dd upstream cotangent")
    assert b.grad is None or float(b.grad.abs().max()) == 0.0


@pytest.mark.parametrize("slope", [0.01, 0.0, 0.2])
def test_batchnorm_act_fused(slope):
    ops = _ops()
    R, C = 3000, 96
    x0, g0, b0, r = rnd(R, C, seed=20), 1 + 0.2 * rnd(C, seed=21), rnd(C, seed=22), rnd(R, C, seed=23)
    xc, gc, bc = x0.double().requires_grad_(), g0.double().requires_grad_(), b0.double().requires_grad_()
    yc = F.leaky_relu(_bn_ref(xc, gc, bc), slope)
    (yc * r.double()).sum().backward()
    x, g, b = x0.cuda().requires_grad_(), g0.cuda().requires_grad_(), b0.cuda().requires_grad_()
    y, _, _ = ops.BatchNormActTrain.apply(x, g, b, 1e-5, slope)
    ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(R * C)).backward()
    close(y, yc.float(), 1e-5)
    close(x.grad, xc.grad.float(), 1e-4)
    close(g.grad, gc.grad.float(), 1e-4)
    close(b.grad, bc.grad.float(), 1e-4)


def test_batch_norm_act_module_semantics():
    """Running statistics / num_batches_tracked side effects and eval mode."""
    ops = _ops()
    R, C = 512, 40
    x0 = rnd(R, C, seed=24) * 2 + 1
    bn_ref = torch.nn.BatchNorm1d(C)
    bn = torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        bn_ref.weight.copy_(1 + 0.1 * rnd(C, seed=25)); bn_ref.bias.copy_(rnd(C, seed=26))
    bn.load_state_dict(bn_ref.state_dict())
    for _ in range(3):
        yr = F.leaky_relu(bn_ref(x0), 0.01)
        y = ops.batch_norm_act(x0.cuda(), bn, 0.01)
    close(y, yr, 1e-5)
    close(bn.running_mean, bn_ref.running_mean, 1e-5)
    close(bn.running_var, bn_ref.running_var, 1e-5)
    assert int(bn.num_batches_tracked) == 3
    bn.eval(); bn_ref.eval()
    xe = x0.cuda().requires_grad_()
    xr = x0.clone().requires_grad_()
    ye = ops.batch_norm_act(xe, bn, 0.01)
    yr = F.leaky_relu(bn_ref(xr), 0.01)
    close(ye, yr, 1e-5)
    r = rnd(R, C, seed=27)
    ops.MeanScale.apply(ops.Mul.apply(ye, r.cuda()), float(R * C)).backward()
    (yr * r).sum().backward()
    close(xe.grad, xr.grad, 1e-5)
    close(bn.weight.grad, bn_ref.weight.grad, 1e-4)
    close(bn.bias.grad, bn_ref.bias.grad, 1e-4)


# ------------------------------------------------------------------------------- activations / layout
def test_lrelu_tanh_mul_axpby():
    ops = _ops()
    x0, r = rnd(1000, 37, seed=28), rnd(1000, 37, seed=29)
    for slope in (0.01, 0.2):
        x = x0.cuda().requires_grad_()
        y = ops.LRelu.apply(x, slope)
        close(y, F.leaky_relu(x0, slope), 0.0)
        ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(x0.numel())).backward()
        close(x.grad, torch.where(x0 > 0, r, r * slope), 1e-6)
    x = x0.cuda().requires_grad_()
    y = ops.Tanh.apply(x)
    close(y, torch.tanh(x0), 1e-6)
    ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(x0.numel())).backward()
    close(x.grad, r * (1 - torch.tanh(x0) ** 2), 1e-5)
    close(ops.Axpby.apply(x0.cuda(), r.cuda(), 2.0, -3.0), 2 * x0 - 3 * r, 1e-6)
    close(ops.scale(x0.cuda(), 0.5), 0.5 * x0, 0.0)


def test_layout_roundtrip_and_strided_inputs():
    ops = _ops()
    pts = rnd(5, 300, 3, seed=30).cuda()
    xt = pts.transpose(2, 1)                                # [B,3,N] non-contiguous view (model.py:249)
    rows = ops.BcnToRows.apply(xt)
    assert torch.equal(rows, pts.reshape(-1, 3))
    x = rnd(3, 70, 45, seed=31).cuda()
    rows = ops.BcnToRows.apply(x)
    assert torch.equal(rows, x.permute(0, 2, 1).reshape(-1, 70))
    back = ops.RowsToBcn.apply(rows, 3, 70, 45)
    assert torch.equal(back, x)
    assert torch.equal(ops.contiguous(x.permute(0, 2, 1)), x.permute(0, 2, 1).contiguous())
    assert torch.equal(ops.contiguous(x[0].t()), x[0].t().contiguous())


def test_concat_cols_broadcast_latent():
    ops = _ops()
    B, N, nz = 3, 50, 16
    x = rnd(B * N, 3, seed=32).cuda()
    zv = rnd(B, nz, seed=33).cuda().requires_grad_()
    out = ops.ConcatCols.apply(x, zv, N, True)
    ref = torch.cat([x, zv.detach().repeat_interleave(N, 0)], 1)
    assert torch.equal(out, ref)
    zf = ref[:, 3:].contiguous().requires_grad_()
    out2 = ops.ConcatCols.apply(x, zf, N, False)
    assert torch.equal(out2, ref)
    r = rnd(B * N, 3 + nz, seed=34).cuda()
    ops.MeanScale.apply(ops.Mul.apply(out, r), float(r.numel())).backward()
    ops.MeanScale.apply(ops.Mul.apply(out2, r), float(r.numel())).backward()
    close(zf.grad, r[:, 3:], 1e-6)
    close(zv.grad, r[:, 3:].reshape(B, N, nz).sum(1), 1e-5)


# ------------------------------------------------------------------------------- pooling / softmax / edges
def test_segmax_first_argmax_and_double_backward_ops():
    ops = _ops()
    B, N, C = 4, 333, 70
    x0 = rnd(B * N, C, seed=35)
    x0[5] = x0[3]                                           # duplicate rows -> ties -> first index wins
    x = x0.cuda().requires_grad_()
    y = ops.SegMax.apply(x, N)
    ref, arg = x0.view(B, N, C).max(1)
    assert torch.equal(y.cpu(), ref)
    r = rnd(B, C, seed=36)
    ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(B * C)).backward()
    xr = x0.clone().requires_grad_()
    (xr.view(B, N, C).max(1)[0] * r).sum().backward()
    close(x.grad, xr.grad, 0.0)
    g = ops.SegMaxGather.apply(x.detach(), arg.int().cuda(), B * N, N)
    assert torch.equal(g.cpu(), ref)


def test_softmax_k_and_kmax():
    ops = _ops()
    P, k, C = 300, 10, 48
    x0, r = rnd(P * k, C, seed=37), rnd(P * k, C, seed=38)
    x = x0.cuda().requires_grad_()
    y = ops.SoftmaxK.apply(x, k)
    xr = x0.clone().requires_grad_()
    yr = torch.softmax(xr.view(P, k, C), 1).view(P * k, C)
    close(y, yr, 1e-5)
    ops.MeanScale.apply(ops.Mul.apply(y, r.cuda()), float(r.numel())).backward()
    (yr * r).sum().backward()
    close(x.grad, xr.grad, 1e-4)
    x = x0.cuda().requires_grad_()
    m = ops.KMax.apply(x, k)
    xr = x0.clone().requires_grad_()
    mr = xr.view(P, k, C).max(1)[0]
    assert torch.equal(m.cpu(), mr)
    r2 = rnd(P, C, seed=39)
    ops.MeanScale.apply(ops.Mul.apply(m, r2.cuda()), float(r2.numel())).backward()
    (mr * r2).sum().backward()
    close(x.grad, xr.grad, 0.0)


def test_softmax_mul_k_fused():
    ops = _ops()
    P, k, C = 200, 10, 64
    x0, y0, r = rnd(P * k, C, seed=71), rnd(P * k, C, seed=72), rnd(P * k, C, seed=73)
    xr, yr = x0.clone().requires_grad_(), y0.clone().requires_grad_()
    ref = yr * torch.softmax(xr.view(P, k, C), 1).view(P * k, C)
    (ref * r).sum().backward()
    x, y = x0.cuda().requires_grad_(), y0.cuda().requires_grad_()
    out = ops.SoftmaxMulK.apply(x, y, k)
    close(out, ref, 1e-5)
    ops.MeanScale.apply(ops.Mul.apply(out, r.cuda()), float(r.numel())).backward()
    close(x.grad, xr.grad, 1e-4)
    close(y.grad, yr.grad, 1e-5)


@pytest.mark.parametrize("C", [64, 3, 30])
def test_edge_combine_forward_backward(C):
    ops = _ops()
    B, N, k = 3, 100, 7
    P = B * N
    rng = np.random.default_rng(40)
    idx = torch.from_numpy(rng.integers(0, N, (B, N, k)).astype(np.int32))
    pc0, pn0, b0, r = rnd(P, C, seed=41), rnd(P, C, seed=42), rnd(C, seed=43), rnd(P * k, C, seed=44)
    pcr, pnr, br = pc0.clone().requires_grad_(), pn0.clone().requires_grad_(), b0.clone().requires_grad_()
    gidx = (idx.long() + (torch.arange(B) * N).view(B, 1, 1)).view(-1)
    ref = (pcr.repeat_interleave(k, 0) + pnr[gidx] - pnr.repeat_interleave(k, 0) + br)
    (ref * r).sum().backward()
    pc, pn, b = pc0.cuda().requires_grad_(), pn0.cuda().requires_grad_(), b0.cuda().requires_grad_()
    out = ops.EdgeCombine.apply(pc, pn, b, idx.cuda(), N, k)
    close(out, ref, 1e-6)
    ops.MeanScale.apply(ops.Mul.apply(out, r.cuda()), float(r.numel())).backward()
    close(pc.grad, pcr.grad, 1e-5)
    close(pn.grad, pnr.grad, 1e-5)
    close(b.grad, br.grad, 1e-5)
    out2 = ops.EdgeCombine.apply(None, pn0.cuda(), None, idx.cuda(), N, k)
    close(out2, pn0[gidx] - pn0.repeat_interleave(k, 0), 1e-6)


def test_adain_vs_instance_norm():
    ops = _ops()
    B, N, C = 3, 200, 64
    x0, s0, r = rnd(B * N, C, seed=45) * 2 + 1, rnd(B * N, 2 * C, seed=46), rnd(B * N, C, seed=47)
    xr, sr = x0.clone().requires_grad_(), s0.clone().requires_grad_()
    xn = F.instance_norm(xr.view(B, N, C).permute(0, 2, 1), eps=1e-5).permute(0, 2, 1).reshape(B * N, C)
    ref = sr[:, :C] * xn + sr[:, C:]
    (ref * r).sum().backward()
    x, s = x0.cuda().requires_grad_(), s0.cuda().requires_grad_()
    out = ops.AdaIN.apply(x, s, N, 1e-5)
    close(out, ref, 1e-5)
    ops.MeanScale.apply(ops.Mul.apply(out, r.cuda()), float(r.numel())).backward()
    close(x.grad, xr.grad, 1e-4)
    close(s.grad, sr.grad, 1e-5)


def test_permute_conv_out_weight():
    ops = _ops()
    w0 = rnd(12, 9, 1, 5, seed=48)
    w = w0.cuda().requires_grad_()
    m = ops.PermuteOCK.apply(w)
    ref = w0[:, :, 0, :].permute(0, 2, 1).reshape(12, 45)
    assert torch.equal(m.cpu(), ref)
    r = rnd(12, 45, seed=49)
    ops.MeanScale.apply(ops.Mul.apply(m, r.cuda()), float(r.numel())).backward()
    close(w.grad, r.view(12, 5, 9).permute(0, 2, 1).reshape(12, 9, 1, 5), 1e-6)


# ------------------------------------------------------------------------------- penalty / loss / optimizer
def test_gp_interp_penalty_and_mean():
    ops = _ops()
    B, N = 6, 128
    real = rnd(B, N, 3, seed=50).cuda().transpose(2, 1)
    fake = rnd(B + 2, 3, N, seed=51).cuda()
    alpha = torch.rand(B, generator=torch.Generator().manual_seed(1)).cuda()
    mix = ops.gp_interpolate(real, fake[:B], alpha)
    close(mix, real + alpha.view(B, 1, 1) * (fake[:B] - real), 1e-6)
    g0 = rnd(B, 3 * N, seed=52) * 0.2
    gr = g0.clone().requires_grad_()
    ref = (((gr.norm(2, dim=1) - 1.0) / 1.0) ** 2).mean() * 10
    ref.backward()
    g = g0.cuda().requires_grad_()
    out = ops.GradPenalty.apply(g, 1.0, 10.0)
    out.backward()
    close(out, ref, 1e-5)
    close(g.grad, gr.grad, 1e-5)
    v = rnd(B, 1, seed=53)
    close(ops.MeanScale.apply(v.cuda(), -1.0), -v.mean(), 1e-6)


def test_adam_matches_torch():
    ops = _ops()
    n = 10007
    p0 = rnd(n, seed=54)
    pr = p0.clone().requires_grad_()
    opt = torch.optim.Adam([pr], lr=1e-4, betas=(0.5, 0.99))
    p, m, v = p0.cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    for t in range(1, 4):
        g = rnd(n, seed=55 + t)
        pr.grad = g.clone()
        opt.step()
        gg = g.cuda()
        ops.L().adam_step(p.data_ptr(), gg.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.5, 0.99, 1e-8, t, 1.0, None)
    close(p, pr.detach(), 1e-6)


@pytest.mark.parametrize("Mo,No,K", [(64, 32, 20000), (64, 3, 16500), (33, 17, 40001), (1, 1, 16384)])
def test_gemm_tall_skinny_weight_gradient(Mo, No, K):
    """C = A^T B with a tiny output and K = #points (EdgeConv1 conv_w.3 / first-layer weight gradients)."""
    ops = _ops()
    A, B = rnd(K, Mo, seed=5), rnd(K, No, seed=6)
    ref = (A.double().t() @ B.double()).float()
    out = ops.gemm_raw(A.cuda(), B.cuda(), None, True, False)
    close(out, ref, 2e-5, "skinny TN")
    out2 = ops.gemm_raw(A.cuda(), B.cuda(), None, True, False, out=out, accumulate=True)
    close(out2, 2 * ref, 2e-5, "skinny TN accumulate")
    # strided operands (column slices of wider row-major matrices)
    Aw, Bw = rnd(K, Mo + 5, seed=7).cuda(), rnd(K, No + 3, seed=8).cuda()
    out3 = ops.gemm_raw(Aw[:, 2:2 + Mo], Bw[:, 1:1 + No], None, True, False)
    close(out3, (Aw[:, 2:2 + Mo].double().t() @ Bw[:, 1:1 + No].double()).float(), 2e-5, "skinny TN strided")


# ------------------------------------------------------------------------------- fused BN + LeakyReLU + max pool
@pytest.mark.parametrize("B,N,C", [(3, 50, 64), (2, 2048, 1024), (5, 333, 132)])
def test_bn_act_segmax_fused_matches_torch(B, N, C):
    """Discriminator.py:77-81,104: BatchNorm1d(train) -> LeakyReLU -> max over points, against torch CPU
    (values, arg-max routing of the gradient, BN parameter gradients); gamma has both signs and a zero."""
    ops = _ops()
    slope, eps = 0.01, 1e-5
    x = rnd(B * N, C, seed=11)
    gamma = rnd(C, seed=12)
    gamma[3] = 0.0
    beta = rnd(C, seed=13)
    gout = rnd(B, C, seed=14)
    xr = x.clone().double().requires_grad_()
    gr, br = gamma.clone().double().requires_grad_(), beta.clone().double().requires_grad_()
    y = F.leaky_relu(F.batch_norm(xr, None, None, gr, br, True, 0.0, eps), slope)
    ref = y.view(B, N, C).max(dim=1).values
    ref.backward(gout.double())

    xg = x.cuda().requires_grad_()
    gg, bg = gamma.cuda().requires_grad_(), beta.cuda().requires_grad_()
    pooled, mean, var = ops.BatchNormActSegMaxTrain.apply(xg, gg, bg, eps, slope, N)
    pooled.backward(gout.cuda())
    close(pooled, ref.float(), TIGHT, "pooled")
    close(mean.view(-1), x.double().mean(0).float(), TIGHT, "mean")
    close(var.view(-1), x.double().var(0, unbiased=False).float(), TIGHT, "var")
    close(gg.grad, gr.grad.float(), 1e-4, "dgamma")
    close(bg.grad, br.grad.float(), 1e-4, "dbeta")
    close(xg.grad, xr.grad.float(), 1e-4, "dx")


def test_bn_act_segmax_module_path_matches_unfused():
    """batch_norm_act_segmax (fused) == SegMax(batch_norm_act(.)) incl. the running-stat side effects."""
    import torch.nn as nn
    ops = _ops()
    B, N, C = 4, 256, 128
    x = rnd(B * N, C, seed=21)
    outs = []
    for fused in (True, False):
        bn = nn.BatchNorm1d(C).cuda().train()
        with torch.no_grad():
            bn.weight.copy_(rnd(C, seed=22).cuda())
            bn.bias.copy_(rnd(C, seed=23).cuda())
        xg = x.cuda().requires_grad_()
        prev = ops.FUSE_BN_POOL
        ops.FUSE_BN_POOL = fused
        try:
            p = ops.batch_norm_act_segmax(xg, bn, 0.01, N)
        finally:
            ops.FUSE_BN_POOL = prev
        p.backward(rnd(B, C, seed=24).cuda())
        outs.append((p.detach(), xg.grad, bn.weight.grad, bn.bias.grad, bn.running_mean.clone(), bn.running_var.clone(),
                     int(bn.num_batches_tracked)))
    for a, b, what in zip(outs[0][:6], outs[1][:6], ["pooled", "dx", "dgamma", "dbeta", "running_mean", "running_var"]):
        close(a, b, 1e-5, what)
    assert outs[0][6] == outs[1][6] == 1


# ------------------------------------------------------------------------------- fused EdgeBlock attention
@pytest.mark.parametrize("P,k,C", [(300, 10, 128), (64, 8, 64), (50, 16, 12)])
def test_bn_act_softmax_mul_k_fused_equals_unfused_chain(P, k, C):
    """Generator.py:78-82 with both BatchNorm2d(train) + LeakyReLU folded into the loads: bit-identical forward,
    gradients and running statistics equal to the generic chain; and the chain itself against torch CPU."""
    import torch.nn as nn
    ops = _ops()
    xw, xy, gout = rnd(P * k, C, seed=31), rnd(P * k, C, seed=32, scale=2.0), rnd(P * k, C, seed=33)
    res = []
    for fused in (True, False):
        bns = []
        for sd in (34, 35):
            bn = nn.BatchNorm2d(C).cuda().train()
            with torch.no_grad():
                bn.weight.copy_(rnd(C, seed=sd).cuda())
                bn.bias.copy_(rnd(C, seed=sd + 10).cuda())
            bns.append(bn)
        a, b = xw.cuda().requires_grad_(), xy.cuda().requires_grad_()
        prev = ops.FUSE_EDGE_ATTENTION
        ops.FUSE_EDGE_ATTENTION = fused
        try:
            out = ops.bn_act_softmax_mul_k(a, bns[0], b, bns[1], 0.01, k)
        finally:
            ops.FUSE_EDGE_ATTENTION = prev
        out.backward(gout.cuda())
        res.append(dict(out=out.detach(), dxw=a.grad, dxy=b.grad, gw=bns[0].weight.grad, bw=bns[0].bias.grad,
                        gy=bns[1].weight.grad, by=bns[1].bias.grad, rm=bns[0].running_mean.clone(),
                        rv=bns[1].running_var.clone()))
    assert torch.equal(res[0]["out"], res[1]["out"])
    for key in res[0]:
        close(res[0][key], res[1][key], 1e-5, key)
    # torch CPU reference of the whole expression
    a, b = xw.double().requires_grad_(), xy.double().requires_grad_()
    g1, b1 = rnd(C, seed=34).double(), rnd(C, seed=44).double()
    g2, b2 = rnd(C, seed=35).double(), rnd(C, seed=45).double()
    wa = F.leaky_relu(F.batch_norm(a, None, None, g1, b1, True, 0.0, 1e-5), 0.01)
    ya = F.leaky_relu(F.batch_norm(b, None, None, g2, b2, True, 0.0, 1e-5), 0.01)
    ref = ya * F.softmax(wa.view(P, k, C), dim=1).view(P * k, C)
    ref.backward(gout.double())
    close(res[0]["out"], ref.float(), 1e-5, "out vs torch")
    close(res[0]["dxw"], a.grad.float(), 1e-4, "dxw vs torch")
    close(res[0]["dxy"], b.grad.float(), 1e-4, "dxy vs torch")


def test_gradient_penalty_accumulation_handed_to_the_bn_backward():
    """Under the penalty's final backward the input of every BatchNorm gets two gradient terms (the layer's own backward
    and its double-backward node).  With the hand-over the second term is added inside the BatchNorm backward's apply
    kernel instead of by autograd's accumulation pass: same gradients, fewer kernels."""
    import spgan_b200 as pkg
    from oracle import spgan_ref as R
    ops = _ops()
    o = R.default_opts(np=256)
    real, fake = rnd(3, 3, 256, seed=70) * 0.5, rnd(3, 3, 256, seed=71) * 0.5
    alpha = torch.rand(3, 1, 1, generator=torch.Generator().manual_seed(5))
    res = {}
    for on in (True, False):
        saved = ops.FUSE_GP_ACCUMULATE
        ops.FUSE_GP_ACCUMULATE = on
        try:
            D = pkg.Discriminator(o)
            D.load_state_dict(R.synth_state(R.discriminator_spec(o), 9))
            D = D.cuda().train()
            n0 = ops.GP_HANDOVERS
            gp = pkg.GradientPenalty(10)(D, real.cuda(), fake.cuda(), alpha=alpha)
            gp.backward()
            assert (ops.GP_HANDOVERS - n0 >= 3) if on else (ops.GP_HANDOVERS == n0), (on, ops.GP_HANDOVERS - n0)   # one per BatchNorm
            res[on] = (float(gp.detach()), {n: p.grad.clone() for n, p in D.named_parameters() if p.grad is not None})
        finally:
            ops.FUSE_GP_ACCUMULATE = saved
    assert res[True][0] == res[False][0]
    assert len(res[False][1]) >= 16
    for n, g in res[False][1].items():
        close(res[True][1][n], g, 2e-5, n)
