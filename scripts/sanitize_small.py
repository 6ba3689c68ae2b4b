"""Small-shape pass over the hand-synchronised kernels for compute-sanitizer (memcheck / racecheck / synccheck):
the TMEM-resident-A GEMM (TMA + mbarrier + tcgen05 pipelines), the streaming SS GEMM, the weight-gradient GEMM, the
fused kNN kernels, the one-launch auction EMD, the pooled BatchNorm and the Chamfer kernel.

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
