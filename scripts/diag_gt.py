"""Diagnostic (GPU): gradient at the tail[0] pre-activation, ours vs oracle fp32/fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import torch.nn.functional as F
from conftest import golden, rel_err
from oracle import spgan_ref as R
import spgan_b200 as pkg
ops = pkg.ops
g = golden("generator_default")
sphere256 = np.load(os.path.join(ROOT, "tests/golden/sphere_256.npy"))
o = R.default_opts(np=256)
Bg = g["out_train"].shape[0]
cap = {}
orig_apply = ops.AddSegVec.apply
def spy(x, v, seg):
    out = orig_apply(x, v, seg)
    out.register_hook(lambda gr: cap.__setitem__("gpu", gr.detach().cpu()))
    cap["gpu_pre"] = out.detach().cpu()
    return out
ops.AddSegVec.apply = spy
G = pkg.Generator(o); G.load_state_dict(R.synth_state(R.generator_spec(o), 51)); G = G.cuda().train()
x = torch.from_numpy(np.tile(sphere256[None], (Bg, 1, 1))).cuda()
z = torch.from_numpy(np.tile(g["z"], (1, 256, 1))).cuda()
G.debug_idx = (None, torch.from_numpy(g["idx2"].astype(np.int32)).cuda())
out = G(x, z)
r = torch.from_numpy(g["r"]).cuda()
ops.MeanScale.apply(ops.Mul.apply(out, r), float(r.numel())).backward()
def run_oracle(dt, key):
    orig = F.conv1d
    def spy_conv(inp, w, b=None, *a, **k):
        y = orig(inp, w, b, *a, **k)
        if w.shape[0] == 256 and w.shape[1] == 640:
            y.register_hook(lambda gr: cap.__setitem__(key, gr.detach()))
            cap[key + "_pre"] = y.detach()
        return y
    R.F.conv1d = spy_conv
    sd = {}
    for k_, v in R.synth_state(R.generator_spec(o), 51).items():
        v = v.to(dt) if v.is_floating_point() else v.clone()
        if v.is_floating_point() and "running" not in k_: v.requires_grad_(True)
        sd[k_] = v
    idx1 = torch.from_numpy(g["idx1"].astype(np.int64)).view(Bg, -1); idx2 = torch.from_numpy(g["idx2"].astype(np.int64)).view(Bg, -1)
    oo = R.generator_forward(sd, x.cpu().to(dt), z.cpu().to(dt), o, training=True, idx1=idx1, idx2=idx2)
    (oo * r.cpu().to(dt)).sum().backward()
    R.F.conv1d = orig
run_oracle(torch.float32, "c32"); run_oracle(torch.float64, "c64")
tr = lambda t: t.permute(0, 2, 1).reshape(-1, 256).numpy()      # [B,256,N] -> rows
print("pre  gpu-vs-64", rel_err(cap["gpu_pre"].numpy(), tr(cap["c64_pre"])), "c32-vs-64", rel_err(tr(cap["c32_pre"]), tr(cap["c64_pre"])))
print("grad gpu-vs-64", rel_err(cap["gpu"].numpy(), tr(cap["c64"])), "c32-vs-64", rel_err(tr(cap["c32"]), tr(cap["c64"])))
d = np.abs(cap["gpu"].numpy() - tr(cap["c64"]))
i = np.unravel_index(np.argmax(d), d.shape)
print("worst", i, cap["gpu"].numpy()[i], tr(cap["c64"])[i], tr(cap["c32"])[i], "pre", cap["gpu_pre"].numpy()[i], tr(cap["c64_pre"])[i])
print("n elements with err>1e-3*max:", int((d > 1e-3 * np.abs(tr(cap["c64"])).max()).sum()), "of", d.size)
