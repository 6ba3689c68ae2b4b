"""K > 256 products of the step on engine 3 (L2 flushed): run once with SPGAN_TSK=0 (gemm_tc.cu, register-converting A
producers) and once without (gemm_tsk_kernel, TMA + TMEM-resident A, K in chunks of 128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

ops = pkg.ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
print("SPGAN_TSK =", os.environ.get("SPGAN_TSK", "1"))
for M, N, K, tb, name in [(131072, 256, 1024, True, "fc2 dgrad (NT)"), (131072, 256, 1024, False, "fc2 dgrad (NN)"),
                          (131072, 128, 1280, True, "conv_out"), (131072, 64, 640, True, "EdgeConv1 conv_out"),
                          (131072, 1280, 512, True, "wide")]:
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") if tb else torch.randn(K, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ref = (A[:512].double() @ (B.t() if tb else B).double()).float()
    ts = []
    for it in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_raw(A, B, None, False, tb, out=out, engine=3); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    err = float((out[:512] - ref).abs().max() / ref.abs().max())
    print("%-20s M=%7d N=%5d K=%5d: %7.3f ms  %6.1f TFLOP/s   max rel err %.2e" % (name, M, N, K, t, 2.0 * M * N * K / t / 1e9, err), flush=True)
