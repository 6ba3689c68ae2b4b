"""GPU parity of the EdgeBlock passes with folded BatchNorm reductions (csrc/edge_fused.cu; Generator.py:75-88):
each fused entry point against the unfused chain it replaces (same kernels' arithmetic: values bit-identical, reductions
to rounding), the composed autograd Functions against a plain torch CPU restatement, and the whole EdgeBlock with the
fusions switched on / off."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
NEG = 0.01


def _ops():
    import spgan_b200
    return spgan_b200.ops


def rnd(*shape, seed=0, scale=1.0):
    rng = np.random.default_rng(seed + sum(shape))
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))


def close(a, b, tol, what=""):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    emax, el2 = rel_err(a, b)
    assert emax <= tol and el2 <= tol, "%s: max %.3e l2 %.3e" % (what, emax, el2)


def _graph(B, N, k, seed):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(0, N, (B, N, k)).astype(np.int32))


@pytest.mark.parametrize("B,N,k,C,with_pc", [(3, 100, 7, 64, True), (2, 256, 10, 128, True), (2, 300, 10, 32, False),
                                             (1, 64, 3, 16, False), (5, 2048, 10, 64, True)])
def test_edge_combine_stats_equals_gather_then_statistics(B, N, k, C, with_pc):
    ops = _ops()
    L = ops.L()
    P = B * N
    idx = _graph(B, N, k, 3).cuda()
    pn = (rnd(P, C, seed=1) + 0.3).cuda()
    pc = rnd(P, C, seed=2).cuda() if with_pc else None
    bias = rnd(C, seed=3).cuda()
    want = ops.EdgeCombine.apply(pc, pn, bias, idx, N, k)
    rows = L.edge_stats_rows(P, C)
    assert rows > 0
    out, cs, cq = ops._edge_combine_stats(pc, pn, bias, idx, N, k)
    assert torch.equal(out, want), "the gather itself must be bit-identical"
    E = P * k
    w64 = want.double().cpu()
    close(cs.double().sum(0).cpu(), w64.sum(0), 1e-6, "column sums")
    close(cq.double().sum(0).cpu(), (w64 * w64).sum(0), 1e-6, "column sums of squares")
    # through spgan_bn_finalize: the statistics a BatchNorm2d(train) computes, and its running buffers
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    mean, rstd, var, scale, shift = ops._bn_finalize(cs, cq, E, bn.weight, bn.bias, bn.eps, ops.bn_running(bn))
    ref = torch.nn.BatchNorm1d(C).train()
    ref(want.cpu())
    close(mean.view(-1), w64.mean(0), 1e-5, "mean")
    close(var.view(-1), w64.var(0, unbiased=False), 1e-5, "var")
    close(bn.running_mean, ref.running_mean, 1e-5, "running_mean")
    close(bn.running_var, ref.running_var, 1e-5, "running_var")
    assert int(bn.num_batches_tracked) == 1
    # deterministic: a second launch gives the same partial rows bit for bit
    out2, cs2, cq2 = ops._edge_combine_stats(pc, pn, bias, idx, N, k)
    assert torch.equal(cs, cs2) and torch.equal(cq, cq2)


def test_edge_stats_rows_envelope():
    L = _ops().L()
    assert L.edge_stats_rows(1000, 64) > 0 and L.edge_stats_rows(1000, 128) > 0 and L.edge_stats_rows(1000, 1024) > 0
    assert L.edge_stats_rows(1000, 30) == 0          # C % 4
    assert L.edge_stats_rows(1000, 96) == 0          # 256 % (C / 4)
    assert L.attn_bwd_rows(1000, 10, 128) > 0 and L.attn_bwd_rows(1000, 10, 96) == 0 and L.attn_bwd_rows(1000, 17, 64) == 0


@pytest.mark.parametrize("P,k,C", [(300, 10, 64), (2048, 10, 128), (100, 5, 32), (4096, 10, 256)])
def test_attention_backward_with_sums_equals_unfused_chain(P, k, C):
    ops = _ops()
    L = ops.L()
    E = P * k
    g, xw, xy = rnd(E, C, seed=5).cuda(), (rnd(E, C, seed=6) * 1.5 + 0.2).cuda(), (rnd(E, C, seed=7) - 0.1).cuda()
    gam_w, bet_w = (rnd(C, seed=8).abs() + 0.5).cuda(), rnd(C, seed=9).cuda() * 0.3
    gam_y, bet_y = -(rnd(C, seed=10).abs() + 0.5).cuda(), rnd(C, seed=11).cuda() * 0.3          # negative gammas too
    mw, rw, _ = ops.col_stats(xw, E, 1e-5)
    my, ry, _ = ops.col_stats(xy, E, 1e-5)
    w, prod = torch.empty_like(xw), torch.empty_like(xw)
    s = torch.cuda.current_stream().cuda_stream
    L.bn_softmax_mul_k(xw.data_ptr(), xy.data_ptr(), P, k, C, mw.data_ptr(), rw.data_ptr(), gam_w.data_ptr(), bet_w.data_ptr(),
                       my.data_ptr(), ry.data_ptr(), gam_y.data_ptr(), bet_y.data_ptr(), NEG, w.data_ptr(), prod.data_ptr(), s)
    # the unfused chain
    dwa0, dya0 = torch.empty_like(xw), torch.empty_like(xw)
    L.bn_softmax_mul_k_bwd(g.data_ptr(), xy.data_ptr(), w.data_ptr(), P, k, C, my.data_ptr(), ry.data_ptr(), gam_y.data_ptr(),
                           bet_y.data_ptr(), NEG, dwa0.data_ptr(), dya0.data_ptr(), s)
    dxw0, sgw0, sgxw0 = ops._norm_bwd(dwa0, xw, NEG, E, mw, rw, gam_w, bet_w)
    dxy0, sgy0, sgxy0 = ops._norm_bwd(dya0, xy, NEG, E, my, ry, gam_y, bet_y)
    # the fused pass
    rows = L.attn_bwd_rows(P, k, C)
    assert rows > 0
    part = torch.empty((rows, 4, C), device="cuda")
    dwa, dya = torch.empty_like(xw), torch.empty_like(xw)
    L.bn_softmax_mul_k_bwd_stats(g.data_ptr(), xw.data_ptr(), xy.data_ptr(), w.data_ptr(), P, k, C, mw.data_ptr(), rw.data_ptr(),
                                 gam_w.data_ptr(), bet_w.data_ptr(), my.data_ptr(), ry.data_ptr(), gam_y.data_ptr(),
                                 bet_y.data_ptr(), NEG, dwa.data_ptr(), dya.data_ptr(), part.data_ptr(), s)
    assert torch.equal(dwa, dwa0) and torch.equal(dya, dya0), "gradients w.r.t. the activated tensors: bit-identical"
    sums = torch.empty((4, C), device="cuda")
    acc = [torch.full((C,), 2.0, device="cuda") for _ in range(4)]
    L.partials_finalize(part.data_ptr(), rows, 4, C, sums.data_ptr(), acc[0].data_ptr(), acc[1].data_ptr(), acc[2].data_ptr(),
                        acc[3].data_ptr(), s)
    for v, ref in enumerate((sgw0, sgxw0, sgy0, sgxy0)):
        scale = float(ref.abs().max()) + 1e-6
        assert float((sums[v] - ref.view(-1)).abs().max()) <= 2e-5 * scale + 1e-5, v
        assert torch.allclose(acc[v], 2.0 + sums[v])
    # ... and the conv_x branch's BatchNorm backward inside the scatter
    B, N = 1, P
    idx = _graph(B, N, k, 12).cuda()
    dpc0, dpn0 = torch.empty((P, C), device="cuda"), torch.empty((P, C), device="cuda")
    L.edge_combine_bwd(dxy0.data_ptr(), idx.data_ptr(), P, N, k, C, dpc0.data_ptr(), dpn0.data_ptr(), s)
    dpc, dpn = torch.empty((P, C), device="cuda"), torch.empty((P, C), device="cuda")
    L.edge_combine_bwd_bn(dya.data_ptr(), xy.data_ptr(), idx.data_ptr(), P, N, k, C, my.data_ptr(), ry.data_ptr(),
                          gam_y.data_ptr(), bet_y.data_ptr(), sgy0.data_ptr(), sgxy0.data_ptr(), NEG, dpc.data_ptr(),
                          dpn.data_ptr(), s)
    close(dpc, dpc0, 1e-6, "dpc")
    close(dpn, dpn0, 1e-5, "dpn (atomic scatter order)")


def _edge_attention_reference(xw, gw, bw, a, d, bias, idx, N, k, gy, by):
    """prod = lrelu(bn(y)) * softmax_k(lrelu(bn(xw))), y = a[p] + d[nbr] - d[p] + bias (torch CPU, float64)."""
    P, C = d.shape
    B = P // N
    gidx = (idx.long() + (torch.arange(B) * N).view(B, 1, 1)).view(-1)
    y = a.repeat_interleave(k, 0) + d[gidx] - d.repeat_interleave(k, 0) + bias

    def bn_act(t, g_, b_):
        m, v = t.mean(0), t.var(0, unbiased=False)
        return F.leaky_relu((t - m) / torch.sqrt(v + 1e-5) * g_ + b_, NEG)

    wa = bn_act(xw, gw, bw).view(P, k, C)
    ya = bn_act(y, gy, by).view(P, k, C)
    return (ya * torch.softmax(wa, dim=1)).view(P * k, C)


@pytest.mark.parametrize("B,N,k,C", [(2, 128, 10, 64), (3, 96, 6, 128)])
def test_edge_attention_function_forward_backward(B, N, k, C):
    ops = _ops()
    P, E = B * N, B * N * k
    idx = _graph(B, N, k, 20)
    t = dict(xw=rnd(E, C, seed=21), gw=rnd(C, seed=22).abs() + 0.5, bw=rnd(C, seed=23) * 0.2, a=rnd(P, C, seed=24),
             d=rnd(P, C, seed=25), bias=rnd(C, seed=26), gy=rnd(C, seed=27).abs() + 0.5, by=rnd(C, seed=28) * 0.2)
    r = rnd(E, C, seed=29)
    ref_in = {kk: v.double().requires_grad_() for kk, v in t.items()}
    ref = _edge_attention_reference(ref_in["xw"], ref_in["gw"], ref_in["bw"], ref_in["a"], ref_in["d"], ref_in["bias"], idx, N,
                                    k, ref_in["gy"], ref_in["by"])
    (ref * r.double()).sum().backward()

    bn_w, bn_y = torch.nn.BatchNorm2d(C).cuda().train(), torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn_w.weight.copy_(t["gw"]); bn_w.bias.copy_(t["bw"]); bn_y.weight.copy_(t["gy"]); bn_y.bias.copy_(t["by"])
    for direct in (False, True):
        g = {kk: t[kk].cuda().requires_grad_() for kk in ("xw", "a", "d", "bias")}
        for p in (bn_w.weight, bn_w.bias, bn_y.weight, bn_y.bias):
            p.grad = torch.zeros_like(p) if direct else None          # .grad present: accumulated in place by the kernel
        assert ops.edge_attention_stats_fusable(P, C, k, bn_w, bn_y)
        out = ops.edge_attention(g["xw"], None, bn_w, g["a"], g["d"], g["bias"], idx.cuda(), N, k, bn_y, NEG)
        close(out, ref, 2e-5, "prod")
        ops.MeanScale.apply(ops.Mul.apply(out, r.cuda()), float(r.numel())).backward()
        close(g["xw"].grad, ref_in["xw"].grad, 2e-4, "d xw")
        close(g["a"].grad, ref_in["a"].grad, 2e-4, "d a")
        close(g["d"].grad, ref_in["d"].grad, 2e-4, "d d")
        close(bn_w.weight.grad, ref_in["gw"].grad, 2e-4, "d gamma_w")
        close(bn_w.bias.grad, ref_in["bw"].grad, 2e-4, "d beta_w")
        close(bn_y.weight.grad, ref_in["gy"].grad, 2e-4, "d gamma_y")
        close(bn_y.bias.grad, ref_in["by"].grad, 2e-4, "d beta_y")
        assert g["bias"].grad is None or float(g["bias"].grad.abs().max()) == 0.0     # feeds a train-mode BN: exactly zero
    assert int(bn_y.num_batches_tracked) == 2 and int(bn_w.num_batches_tracked) == 2


def test_edge_block_fused_equals_unfused():
    """The whole EdgeBlock (train mode), forward and every gradient, with the gather-side fusions on and off."""
    import spgan_b200 as pkg
    ops = pkg.ops
    B, Cin, F_, N, k = 2, 64, 128, 256, 10
    torch.manual_seed(0)
    blk = pkg.EdgeBlock(Cin, F_, k).cuda().train()
    x0 = rnd(B, Cin, N, seed=30).cuda()
    idx = _graph(B, N, k, 31).cuda()
    r = rnd(B, F_, N, seed=32).cuda()
    res = {}
    saved = (ops.FUSE_EDGE_STATS,)
    try:
        for fused in (True, False):
            ops.FUSE_EDGE_STATS = fused
            blk.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_()
            before = ops.L().launches
            out = blk(x, idx)
            fwd_launches = ops.L().launches - before
            (out * r).sum().backward()
            res[fused] = (out.detach(), x.grad.clone(), {n: p.grad.clone() for n, p in blk.named_parameters() if p.grad is not None},
                          fwd_launches)
    finally:
        (ops.FUSE_EDGE_STATS,) = saved
    with torch.no_grad():                       # the critic phase's generator pass: no softmax weights are written
        out_ng = blk(x0.clone(), idx)
    assert torch.equal(out_ng, res[True][0]), "no-grad forward must equal the recorded one bit for bit"
    close(res[True][0], res[False][0], 2e-5, "forward")
    close(res[True][1], res[False][1], 5e-4, "input gradient")
    scale = max(float(v.abs().max()) for v in res[False][2].values())
    for n, gref in res[False][2].items():
        gf = res[True][2][n]
        assert float((gf - gref).abs().max()) <= 5e-4 * max(float(gref.abs().max()), 1e-3 * scale), n
    assert res[True][3] < res[False][3], "the fused path must launch fewer kernels (%d vs %d)" % (res[True][3], res[False][3])
