"""Diagnostic (GPU): global_conv BatchNorm1d (B rows) intermediates vs oracle fp32 / fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import golden, rel_err
from oracle import spgan_ref as R
import spgan_b200 as pkg
ops = pkg.ops
g = golden("generator_default")
sphere256 = np.load(os.path.join(ROOT, "tests/golden/sphere_256.npy"))
o = R.default_opts(np=256)
Bg = g["out_train"].shape[0]
rec_gpu = []
orig = ops.batch_norm_act
def spy(y, bn, slope):
    z = orig(y, bn, slope)
    if y.shape[0] == Bg: rec_gpu.append((y.detach().cpu(), z.detach().cpu()))
    return z
ops.batch_norm_act = spy
G = pkg.Generator(o); G.load_state_dict(R.synth_state(R.generator_spec(o), 51)); G = G.cuda().train()
x = torch.from_numpy(np.tile(sphere256[None], (Bg, 1, 1))).cuda()
z = torch.from_numpy(np.tile(g["z"], (1, 256, 1))).cuda()
G.debug_idx = (None, torch.from_numpy(g["idx2"].astype(np.int32)).cuda())
out = G(x, z)
def run_oracle(dt):
    rec = []
    orig_bn = R._batch_norm
    def spy_bn(xx, sd, p, training):
        yy = orig_bn(xx, sd, p, training)
        if xx.dim() == 2: rec.append((xx.detach(), yy.detach()))
        return yy
    R._batch_norm = spy_bn
    sd = {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in R.synth_state(R.generator_spec(o), 51).items()}
    idx1 = torch.from_numpy(g["idx1"].astype(np.int64)).view(Bg, -1); idx2 = torch.from_numpy(g["idx2"].astype(np.int64)).view(Bg, -1)
    R.generator_forward(sd, x.cpu().to(dt), z.cpu().to(dt), o, training=True, idx1=idx1, idx2=idx2)
    R._batch_norm = orig_bn
    return rec
r32, r64 = run_oracle(torch.float32), run_oracle(torch.float64)
for i in range(2):
    print("BN", i, "pre: gpu-vs-64", rel_err(rec_gpu[i][0].numpy(), r64[i][0].numpy()), "cpu32-vs-64", rel_err(r32[i][0].numpy(), r64[i][0].numpy()))
    pre = r64[i][0].numpy(); print("   |mean|/std over batch: median", np.median(np.abs(pre.mean(0)) / (pre.std(0) + 1e-30)))
    # compare post-activation via lrelu of oracle BN output
    post64 = torch.nn.functional.leaky_relu(r64[i][1], 0.01).numpy(); post32 = torch.nn.functional.leaky_relu(r32[i][1], 0.01).numpy()
    print("   post: gpu-vs-64", rel_err(rec_gpu[i][1].numpy(), post64), "cpu32-vs-64", rel_err(post32, post64))
