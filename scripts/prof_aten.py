"""Which stock-ATen kernels still run inside one eager WGAN-GP step, and which autograd nodes launch them
(torch.profiler; the product's own kernels go through ctypes and show up as plain CUDA launches)."""
import os
import sys
from collections import Counter

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402
from spgan_b200 import synthetic  # noqa: E402

B, N, NZ = 64, 2048, 128
dev = torch.device("cuda", 0)
torch.manual_seed(123)
opts = type("Opts", (), dict(np=N, nk=20, nz=NZ, softmax=True, off=False, attn=False, use_head=False, eql=False,
                             z_norm=False, small_d=False))()
G, D = pkg.Generator(opts).to(dev).train(), pkg.Discriminator(opts).to(dev).train()
tr = pkg.WGANGPTrainer(G, D)
rng = np.random.default_rng(0)
ball, _ = synthetic.sphere_template(N)
x = torch.from_numpy(np.tile(ball[None], (B, 1, 1))).to(dev)
real = torch.from_numpy(synthetic.synthetic_chairs(rng, B, N)).to(dev).transpose(2, 1)
z = torch.from_numpy(synthetic.latent_vectors(rng, B, NZ)).to(dev).expand(B, N, NZ)
alpha = torch.rand(B, 1, 1).to(dev)
for _ in range(2):
    tr.step(x, z, z, real, alpha=alpha)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.step(x, z, z, real, alpha=alpha)
    torch.cuda.synchronize()
ev = prof.events()
kern = Counter()
ktime = Counter()
for e in ev:
    if e.device_type == torch.autograd.DeviceType.CUDA and ("at::" in e.name or "at_cuda" in e.name or "elementwise" in e.name):
        kern[e.name[:90]] += 1
        ktime[e.name[:90]] += e.device_time
print("== stock kernels in one step")
for k, c in kern.most_common(20):
    print("%5d  %8.1f us  %s" % (c, ktime[k], k))
cpu = Counter()
for e in ev:
    if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and e.cpu_parent is not None:
        p = e.cpu_parent
        while p.cpu_parent is not None and p.name.startswith("aten::"):
            p = p.cpu_parent
        if e.name in ("aten::add_", "aten::add", "aten::fill_", "aten::zero_", "aten::copy_", "aten::zeros", "aten::ones_like",
                      "aten::mul", "aten::sum", "aten::clone"):
            cpu[(e.name, p.name[:70])] += 1
print("== aten ops by enclosing node")
for (n, p), c in cpu.most_common(40):
    print("%5d  %-16s <- %s" % (c, n, p))
