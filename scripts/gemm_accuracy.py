"""Diagnostic (GPU): relative error of the three spgan_gemm engines vs fp64 for the step's K sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spgan_b200 as pkg
ops = pkg.ops
torch.manual_seed(0)
for K in (64, 128, 256, 1024, 1280):
    for dist in ("randn", "positive"):
        A = torch.randn(4096, K); B = torch.randn(128, K)
        if dist == "positive": A = A.abs()          # post-activation-like operand: products do not cancel
        ref = A.double() @ B.double().t()
        row = []
        for eng in (0, 1, 2):
            out = ops.gemm_raw(A.cuda(), B.cuda(), None, False, True, engine=eng).cpu().double()
            e = (out - ref)
            row.append("eng%d max %.2e rms %.2e bias %.2e" % (eng, e.abs().max() / ref.abs().max(), e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt(), (e.mean() / ref.abs().mean()).abs()))
        print("K=%5d %-8s | " % (K, dist) + " | ".join(row))
