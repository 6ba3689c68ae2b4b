"""Diagnostic (GPU): per-parameter gradient error of the generator vs the reference golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import golden, rel_err, pack_like_golden
from oracle import spgan_ref as R
import spgan_b200 as pkg

g = golden("generator_default")
sphere256 = np.load(os.path.join(ROOT, "tests/golden/sphere_256.npy"))
o = R.default_opts(np=256)
G = pkg.Generator(o); G.load_state_dict(R.synth_state(R.generator_spec(o), 51)); G = G.cuda().train()
Bg = g["out_train"].shape[0]
x = torch.from_numpy(np.tile(sphere256[None], (Bg, 1, 1))).cuda()
z = torch.from_numpy(np.tile(g["z"], (1, 256, 1))).cuda()
G.debug_idx = (None, torch.from_numpy(g["idx2"].astype(np.int32)).cuda())
out = G(x, z)
r = torch.from_numpy(g["r"]).cuda()
pkg.ops.MeanScale.apply(pkg.ops.Mul.apply(out, r), float(r.numel())).backward()
print("out", rel_err(out.detach().cpu().numpy(), g["out_train"]))
print("x1", rel_err(G._last_x1.view(Bg, 256, 64).permute(0, 2, 1).cpu().numpy(), g["x1"]))
for k, p in G.named_parameters():
    if "grad." + k in g:
        ref = g["grad." + k]
        e = rel_err(pack_like_golden(p.grad), ref)
        print("%-32s max %.2e l2 %.2e  |ref| %.3g noise %.1e" % (k, e[0], e[1], np.abs(ref).max(), float(g.get("noise.grad." + k, 0))))
