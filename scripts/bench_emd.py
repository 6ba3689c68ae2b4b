"""Microbenchmark (GPU): auction-EMD kernel (csrc/emd.cu) at the evaluation sizes -- cloud pairs per second and
iterations actually executed are not observable from outside, so the oracle's trace on one pair is printed next to it.
    python scripts/bench_emd.py [B] [n] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import spgan_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 300
rng = np.random.default_rng(0)
a = torch.from_numpy(rng.random((B, n, 3), dtype=np.float32)).cuda()
b = torch.from_numpy(rng.random((B, n, 3), dtype=np.float32)).cuda()
mod = pkg.emdModule()
ts = []
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dist, ass = mod(a, b, 0.005, iters); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
print("EMD auction B=%d n=%d eps=0.005 iters<=%d: %.3f ms  %.1f pairs/s  mean sqrt(dist) %.5f  bijective %.3f" % (
    B, n, iters, t, B / t * 1e3, float(dist.sqrt().mean()),
    float(np.mean([len(set(r)) == n for r in ass.cpu().numpy()]))))
