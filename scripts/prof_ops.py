"""A few launches of every non-GEMM hot kernel at the BASELINE sizes, for one `ncu --set full` pass:
kNN (C = 64), the critic's fused BN+LeakyReLU+max-pool (forward + backward), the fused EdgeBlock attention
(forward + backward) and the pairwise Chamfer kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import spgan_b200 as pkg
from spgan_b200 import synthetic
ops = pkg.ops
B, N, k = 64, 2048, 10
P, E = B * N, B * N * k
torch.manual_seed(0)
x = torch.randn(B, 64, N, device="cuda")
ops.knn_indices(x, k)
# critic tail
y = torch.randn(P, 1024, device="cuda", requires_grad=True)
g, b = torch.rand(1024, device="cuda") + 0.5, torch.randn(1024, device="cuda")
g.requires_grad_(); b.requires_grad_()
pooled, _, _ = ops.BatchNormActSegMaxTrain.apply(y, g, b, 1e-5, 0.01, N)
pooled.backward(torch.randn_like(pooled))
del y, pooled
# EdgeBlock2 attention
xw = torch.randn(E, 128, device="cuda", requires_grad=True)
xy = torch.randn(E, 128, device="cuda", requires_grad=True)
gw, bw, gy, by = (torch.rand(128, device="cuda").requires_grad_() for _ in range(4))
prod = ops.BnActSoftmaxMulKTrain.apply(xw, gw, bw, xy, gy, by, 1e-5, 1e-5, 0.01, k)
prod.backward(torch.randn_like(prod))
del xw, xy, prod
# pairwise Chamfer
rng = np.random.default_rng(1)
a = torch.from_numpy(synthetic.synthetic_chairs(rng, 48, N)).cuda()
c = torch.from_numpy(synthetic.synthetic_chairs(rng, 48, N)).cuda()
pkg.pairwise_CD(a, c)
torch.cuda.synchronize()
