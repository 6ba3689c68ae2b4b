"""Microbenchmark (GPU): kNN(+group) kernel at the BASELINE sizes (CUDA events, median of 7)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spgan_b200 as pkg
ops = pkg.ops
for (B, C, N, k, ee) in [(64, 64, 2048, 10, False), (64, 3, 2048, 10, False), (64, 128, 2048, 10, False), (64, 64, 2048, 10, True),
                         (64, 3, 2048, 10, True)]:
    x = torch.randn(B, C, N, device="cuda")
    ts = []
    for it in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.knn_indices(x, k, want_ee=ee); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[3]
    flops = 2.0 * B * N * N * C
    byts = B * C * N * 4 + B * N * k * 4 + (B * 2 * C * N * k * 4 if ee else 0)
    print("B=%d C=%3d N=%d k=%d ee=%d: %7.3f ms  %6.2f TFLOP/s (fp32 FFMA)  %7.1f GB/s algorithmic" % (
        B, C, N, k, ee, t, flops / t / 1e9, byts / t / 1e6))
