"""One spgan_gemm_fused launch per shape for ncu: python scripts/prof_fused_shape.py M N K [pro]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
pro = len(sys.argv) > 4
A = torch.randn(M, K, device="cuda")
B = torch.randn(N, K, device="cuda")
out = torch.empty(M, N, device="cuda")
sc, sh = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda")
for _ in range(4):
    if pro:
        pkg.ops.gemm_fused_raw(A, B, None, tb=True, out=out, a_scale=sc, a_shift=sh, a_slope=0.01, want_stats=N <= 256)
    else:
        pkg.ops.gemm_fused_raw(A, B, None, tb=True, out=out)
torch.cuda.synchronize()
