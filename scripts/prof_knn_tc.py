"""A few spgan_knn_rows launches at the generator's EdgeConv2 shape for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

B, C, N, k = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 2048, 10
rows = torch.randn(B * N, C, device="cuda")
for _ in range(3):
    pkg.ops.knn_indices_rows(rows, B, N, k)
torch.cuda.synchronize()
