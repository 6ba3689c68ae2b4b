"""Microbenchmark (GPU): spgan_gemm engines on the step's big shapes (CUDA events, L2-cold inputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spgan_b200 as pkg
ops = pkg.ops
shapes = [(131072, 1024, 256, "D fc2 fwd"), (131072, 256, 1024, "D fc2 dgrad"), (131072, 256, 128, "D mlps.6"),
          (1310720, 128, 64, "EdgeConv2 conv_w.3"), (131072, 128, 1280, "EdgeConv2 conv_out"),
          (131072, 128, 128, "head.2 / adain1.style"), (131072, 256, 128, "tail.0 / adain2.style")]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for M, N, K, name in shapes:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); out = torch.empty(M, N, device="cuda")
    for eng in (0, 1, 2):
        ts = []
        for it in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm_raw(A, B, None, False, True, out=out, engine=eng); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        st = int(ops.LAST_TC_WORKSPACE.view(torch.int32)[0]) if eng >= 1 else 0
        print("%-26s M=%8d N=%5d K=%5d engine %d: %8.3f ms  %7.1f TFLOP/s  out %6.0f GB/s  status %d" % (
            name, M, N, K, eng, t, 2.0 * M * N * K / t / 1e9, M * N * 4 / t / 1e6, st))

print("--- weight gradients C = A^T B")
for K, Mo, No, name in [(131072, 1024, 256, "D fc2 wgrad"), (131072, 256, 128, "D mlps.6 wgrad"),
                        (131072, 128, 1280, "EdgeConv2 conv_out wgrad"), (1310720, 128, 64, "EdgeConv2 conv_w.3 wgrad")]:
    A = torch.randn(K, Mo, device="cuda"); B = torch.randn(K, No, device="cuda"); out = torch.empty(Mo, No, device="cuda")
    for eng in (0, 1):
        ts = []
        for it in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm_raw(A, B, None, True, False, out=out, engine=eng); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print("%-26s Mo=%5d No=%5d K=%8d engine %d: %8.3f ms  %7.1f TFLOP/s  in %6.0f GB/s" % (
            name, Mo, No, K, eng, t, 2.0 * Mo * No * K / t / 1e9, K * (Mo + No) * 4 / t / 1e6))
