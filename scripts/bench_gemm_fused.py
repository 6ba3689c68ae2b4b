"""Times spgan_gemm (engines 1, 3) and spgan_gemm_fused on the K <= 256 shapes of the step, L2 flushed between calls."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

ops = pkg.ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def med(fn, n=7):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2]


for M, N, K in [(131072, 1024, 256), (131072, 256, 128), (131072, 128, 256), (131072, 128, 128), (131072, 128, 64),
                (131072, 64, 128), (1310720, 128, 64), (1310720, 64, 32), (1310720, 64, 128), (131072, 1280, 128)]:
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    out = torch.empty(M, N, device="cuda")
    sc, sh = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda")
    os.environ["X"] = "1"
    t1 = med(lambda: ops.gemm_raw(A, B, None, False, True, out=out, engine=1))
    t3 = med(lambda: ops.gemm_raw(A, B, None, False, True, out=out, engine=3))
    tf = med(lambda: ops.gemm_fused_raw(A, B, None, tb=True, out=out))
    tfs = med(lambda: ops.gemm_fused_raw(A, B, None, tb=True, out=out, a_scale=sc, a_shift=sh, a_slope=0.01, want_stats=N <= 256))
    fl = 2.0 * M * N * K / 1e9
    byts = 4.0 * M * (N + K) / 1e6
    print("M=%-8d N=%-5d K=%-4d  tf32x3 %.3f ms (%5.1f TF)  engine3 %.3f ms (%5.1f TF)  fused %.3f ms (%5.1f TF, %5.2f TB/s)  "
          "fused+pro+stats %.3f ms" % (M, N, K, t1, fl / t1, t3, fl / t3, tf, fl / tf, byts / tf / 1e3, tfs), flush=True)
