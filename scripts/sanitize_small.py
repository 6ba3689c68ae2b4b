"""Small-shape pass over the hand-synchronised kernels for compute-sanitizer (memcheck / racecheck / synccheck):
the TMEM-resident-A GEMMs (TMA + mbarrier + tcgen05 pipelines; K <= 256 and chunked K), the streaming SS GEMM, the
weight-gradient GEMMs (with operand prologue), the thin streaming products, the fused EdgeBlock passes, the kNN kernels
(CUDA-core and tensor-core filter + refine), the one-launch auction EMD, the pooled BatchNorm and the Chamfer kernel.

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402
from spgan_b200 import emd as emd_mod  # noqa: E402

ops = pkg.ops
rng = np.random.default_rng(0)
t = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).cuda()


def check(name, got, ref, tol=1e-4):
    err = float((got.double().cpu() - ref).abs().max() / (ref.abs().max() + 1e-30))
    print("%-28s rel err %.2e" % (name, err), flush=True)
    assert err < tol, name


A, B, bias = t(300, 100), t(70, 100), t(70)
sc, sh = t(100).abs() + 0.5, t(100)
out, cs, cq = ops.gemm_fused_raw(A, B, bias, tb=True, a_scale=sc, a_shift=sh, a_slope=0.01, want_stats=True)
pa = A.double().cpu() * sc.double().cpu() + sh.double().cpu()
pa = torch.where(pa > 0, pa, pa * 0.01)
check("gemm_ts (fused)", out, pa @ B.double().cpu().t() + bias.double().cpu())
A2, B2 = t(400, 256), t(200, 256)
check("gemm_ts (K=256, 4 n-tiles)", ops.gemm_fused_raw(A2, B2, None, tb=True), A2.double().cpu() @ B2.double().cpu().t())
A3, B3 = t(260, 320), t(96, 320)
for eng in (1, 3):
    check("gemm_tc SS engine %d" % eng, ops.gemm_raw(A3, B3, None, False, True, engine=eng), A3.double().cpu() @ B3.double().cpu().t())
A4, B4 = t(4096, 64), t(4096, 48)
check("gemm_tc_tn", ops.gemm_raw(A4, B4, None, True, False, engine=1), A4.double().cpu().t() @ B4.double().cpu())
A5, B5 = t(300, 512), t(200, 512)
assert ops.L().gemm_bigk_route(300, 200, 512, A5.data_ptr(), 512) == 1
check("gemm_tsk (K=512, chunked)", ops.gemm_raw(A5, B5, bias.new_zeros(200), False, True, engine=3), A5.double().cpu() @ B5.double().cpu().t())
gz, xp = t(4096, 32), t(4096, 48)
scw, shw = t(48).abs() + 0.5, t(48)
ws_b = ops.L().gemm_wgrad_workspace(32, 48, 4096)
wsw = torch.empty(ws_b // 4, device="cuda")
dW = torch.empty(32, 48, device="cuda")
ops.L().gemm_wgrad_fused(32, 48, 4096, gz.data_ptr(), 32, xp.data_ptr(), 48, scw.data_ptr(), shw.data_ptr(), 0.01, dW.data_ptr(), 48, 0,
                         wsw.data_ptr(), ws_b, ops._stream())
pw = xp.double().cpu() * scw.double().cpu() + shw.double().cpu()
check("gemm_wg (prologue)", dW, gz.double().cpu().t() @ torch.where(pw > 0, pw, pw * 0.01))
At, Bt = t(4100, 3), t(64, 3)
check("gemm_thin_k", ops.gemm_raw(At, Bt, None, False, True), At.double().cpu() @ Bt.double().cpu().t())
An, Bn = t(4100, 64), t(64, 3)
check("gemm_thin_n", ops.gemm_raw(An, Bn, None, False, False), An.double().cpu() @ Bn.double().cpu())
Aw, Bw = t(16500, 64), t(16500, 3)
check("gemm_tn_thin", ops.gemm_raw(Aw, Bw, None, True, False), Aw.double().cpu().t() @ Bw.double().cpu())
# EdgeBlock with every fused pass (gather + statistics, attention forward / backward with sums, BN backward in the scatter)
blk = pkg.EdgeBlock(16, 32, 5).cuda().train()
for p_ in blk.parameters():
    p_.grad = torch.zeros_like(p_)
xe = t(2 * 128, 16).requires_grad_()
ie = ops.knn_indices_rows(xe.detach(), 2, 128, 5)
oe = blk.forward_rows(xe, ie, 2, 128)
oe.backward(torch.ones_like(oe))
print("EdgeBlock fused fwd+bwd ok", float(oe.sum()), flush=True)
rows = t(2 * 256, 64)
i_tc = ops.knn_indices_rows(rows, 2, 256, 10)
i_cc = ops.knn_indices(rows.view(2, 256, 64).permute(0, 2, 1).contiguous(), 10)
assert torch.equal(i_tc, i_cc)
print("knn_tc (filter + refine) == knn_group_fast", flush=True)
x = t(2, 16, 256)
idx = ops.knn_indices(x, 10)
from oracle import knn_ref  # noqa: E402
assert np.array_equal(idx.cpu().numpy(), knn_ref.knn(x.cpu().numpy(), 10))
print("knn_group_fast bit-exact", flush=True)
x3 = t(1, 3, 132)
assert np.array_equal(ops.knn_indices(x3, 5, want_ee=True)[0].cpu().numpy(), knn_ref.knn(x3.cpu().numpy(), 5))
print("knn_group (general + ee) bit-exact", flush=True)
a, b = torch.rand(2, 128, 3).cuda(), torch.rand(2, 128, 3).cuda()
d, s = emd_mod.emdModule()(a, b, 0.005, 20)
print("emd_auction ok", float(d.sum()), flush=True)
cd = pkg.pairwise_CD(torch.rand(3, 256, 3).cuda(), torch.rand(2, 256, 3).cuda())
print("pairwise_chamfer ok", float(cd.sum()), flush=True)
D = pkg.Discriminator(type("O", (), {"small_d": False})()).cuda().train()
o = D(t(2, 3, 256).requires_grad_())
o.sum().backward()
print("critic fwd+bwd (bn_pool, colreduce) ok", float(o.sum()), flush=True)
torch.cuda.synchronize()
print("sanitize_small: all ok")
