"""One weight-gradient product (C = A^T B) per launch for ncu: python scripts/prof_wgrad_shape.py Mo No K"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

Mo, No, K = (int(v) for v in sys.argv[1:4])
A = torch.randn(K, Mo, device="cuda")
B = torch.randn(K, No, device="cuda")
out = torch.empty(Mo, No, device="cuda")
for _ in range(4):
    pkg.ops.gemm_raw(A, B, None, True, False, out=out, engine=3)
torch.cuda.synchronize()
