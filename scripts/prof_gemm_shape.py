"""One spgan_gemm shape, a few launches (for `ncu -k regex:gemm_tc_kernel`):  python scripts/prof_gemm_shape.py M N K"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spgan_b200 as pkg
M, N, K = (int(v) for v in sys.argv[1:4])
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); out = torch.empty(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
for _ in range(3):
    pkg.ops.gemm_raw(A, B, bias, False, True, out=out)
torch.cuda.synchronize()
