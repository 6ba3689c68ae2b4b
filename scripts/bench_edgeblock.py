"""EdgeConv1 / EdgeConv2 (EdgeBlock(3,64) / EdgeBlock(64,128), k=10) at B=64, N=2048: forward (no-grad and recorded) and
forward + backward, CUDA events, L2 flushed between repetitions; per-entry-point device times of one instrumented pass."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

ops = pkg.ops
B, N, k = 64, 2048, 10
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def med(fn, n=9):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2]


torch.manual_seed(0)
for cin, cout in ((3, 64), (64, 128)):
    blk = pkg.EdgeBlock(cin, cout, k).cuda().train()
    for p in blk.parameters():
        p.grad = torch.zeros_like(p)
    x = torch.randn(B * N, cin, device="cuda", requires_grad=True)
    feat = torch.randn(B * N, max(cin, 4), device="cuda")
    idx = ops.knn_indices_rows(feat, B, N, k)
    g = torch.randn(B * N, cout, device="cuda")

    def fwd_nograd():
        with torch.no_grad():
            blk.forward_rows(x, idx, B, N)

    def fwd():
        blk.forward_rows(x, idx, B, N)

    def fwd_bwd():
        blk.forward_rows(x, idx, B, N).backward(g)

    t0, t1, t2 = med(fwd_nograd), med(fwd), med(fwd_bwd)
    print("EdgeBlock(%d,%d): forward no-grad %.3f ms, forward %.3f ms, forward+backward %.3f ms" % (cin, cout, t0, t1, t2), flush=True)
    ops.L().profile = []
    fwd_bwd()
    torch.cuda.synchronize()
    agg = {}
    for name, _, e0, e1 in ops.L().profile:
        agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1)
    ops.L().profile = None
    print("   " + ", ".join("%s %.3f" % (n.replace("spgan_", ""), t) for n, t in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
