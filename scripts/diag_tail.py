"""Diagnostic (GPU): the folded tail (per-cloud bias) vs torch conv1d on cat(global, x2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.nn.functional as F
from conftest import rel_err
import spgan_b200 as pkg
ops = pkg.ops
rng = np.random.default_rng(0)
B, N = 4, 256
f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
x2, g, W0, b0, W2, b2, W4, b4, r = f(B*N,128), f(B,512), f(256,640)/25, f(256)*.1, f(64,256)/16, f(64)*.1, f(3,64)/8, f(3)*.1, f(B*N,3)
def ref(dt):
    xs=[t.detach().clone().to(dt).requires_grad_() for t in (x2,g,W0,b0,W2,b2,W4,b4)]
    x2_,g_,W0_,b0_,W2_,b2_,W4_,b4_=xs
    feat=torch.cat([g_.repeat_interleave(N,0), x2_],1)
    t=F.leaky_relu(feat@W0_.t()+b0_,0.01); t2=F.leaky_relu(t@W2_.t()+b2_,0.01); o=torch.tanh(t2@W4_.t()+b4_)
    (o*r.to(dt)).sum().backward()
    return o,[t.grad for t in xs]
o64,g64=ref(torch.float64); o32,g32=ref(torch.float32)
xs=[t.detach().clone().cuda().requires_grad_() for t in (x2,g,W0,b0,W2,b2,W4,b4)]
x2_,g_,W0_,b0_,W2_,b2_,W4_,b4_=xs
gb=ops.linear(g_, W0_[:, :512], b0_)
t=ops.LRelu.apply(ops.AddSegVec.apply(ops.linear(x2_, W0_[:, 512:]), gb, N),0.01)
t2=ops.LRelu.apply(ops.linear(t,W2_,b2_),0.01)
o=ops.Tanh.apply(ops.linear(t2,W4_,b4_))
ops.MeanScale.apply(ops.Mul.apply(o,r.cuda()),float(r.numel())).backward()
print('out', rel_err(o.detach().cpu().numpy(), o64.detach().numpy()))
for n,a,b,c in zip("x2 g W0 b0 W2 b2 W4 b4".split(), xs, g64, g32):
    print(n, 'gpu-vs-64', rel_err(a.grad.cpu().numpy(), b.numpy()), 'cpu32-vs-64', rel_err(c.numpy(), b.numpy()))
