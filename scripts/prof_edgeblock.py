"""EdgeConv2 (EdgeBlock(64, 128, k=10)) forward + backward at the BASELINE size (B=64, N=2048) plus the critic's fused
BN + LeakyReLU + max-pool, for one `ncu --set full` pass over the memory-bound kernels of the step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

ops = pkg.ops
B, N, k = 64, 2048, 10
torch.manual_seed(0)
blk = pkg.EdgeBlock(64, 128, k).cuda().train()
for p in blk.parameters():
    p.grad = torch.zeros_like(p)
x = torch.randn(B * N, 64, device="cuda", requires_grad=True)
idx = ops.knn_indices_rows(x.detach(), B, N, k)
for _ in range(2):
    out = blk.forward_rows(x, idx, B, N)
    out.backward(torch.randn_like(out))
y = torch.randn(B * N, 1024, device="cuda", requires_grad=True)
g, b = (torch.rand(1024, device="cuda") + 0.5).requires_grad_(), torch.randn(1024, device="cuda").requires_grad_()
pooled, _, _ = ops.BatchNormActSegMaxTrain.apply(y, g, b, 1e-5, 0.01, N)
pooled.backward(torch.randn_like(pooled))
torch.cuda.synchronize()
