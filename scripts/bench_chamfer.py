"""Microbenchmark (GPU): pairwise Chamfer distance, BASELINE configs[4] per-pair shape (N = 2048).
8 fp32 instructions per point pair => roofline = issue rate of the CUDA cores (148 SM x 128 lanes x clock)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import spgan_b200 as pkg
from spgan_b200 import synthetic

S = R = int(os.environ.get("CHAMFER_CLOUDS", "96"))
N = 2048
rng = np.random.default_rng(123)
a = torch.from_numpy(synthetic.synthetic_chairs(rng, S, N)).cuda()
b = torch.from_numpy(synthetic.synthetic_chairs(rng, R, N)).cuda()
ts = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cd = pkg.pairwise_CD(a, b); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts[1:])[len(ts[1:]) // 2]
pairs = S * R
ops = 8.0 * pairs * N * N
peak = 148 * 128 * 1.965e9
print(json.dumps({"bench": "pairwise_chamfer", "clouds": [S, R], "points": N, "ms": t, "cloud_pairs_per_s": pairs / (t / 1e3),
                  "fp32_inst_per_s": ops / (t / 1e3), "frac_of_fp32_issue_peak": ops / (t / 1e3) / peak,
                  "est_5000x5000_s_1gpu": 25e6 / (pairs / (t / 1e3)), "mmd_cov": pkg.lgan_mmd_cov(cd)}))
