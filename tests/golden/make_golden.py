#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference (liruihui/SP-GAN) owns no tests / golden vectors for this path, so the
fixtures written here are the pins for the oracle (oracle/) and for the CUDA product.
Weights are synthetic and regenerated from a seed by oracle.spgan_ref.synth_state, so
only inputs that cannot be regenerated and the reference OUTPUTS are stored.  Large
gradient tensors are stored as a strided subsample (see `pack`).
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

REF = os.environ.get("SPGAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
sys.dont_write_bytecode = True

from oracle import spgan_ref as R  # noqa: E402

SUBSAMPLE_ABOVE = 32768
SUBSAMPLE_STRIDE = 16


def pack(t):
    """fp32 numpy copy; tensors above SUBSAMPLE_ABOVE elements keep every 16th flat element."""
    a = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    if a.size > SUBSAMPLE_ABOVE:
        a = np.ascontiguousarray(a.reshape(-1)[::SUBSAMPLE_STRIDE])
    return a


ONLY = [t for t in os.environ.get("SPGAN_GOLDEN_ONLY", "").split(",") if t]     # e.g. generator_eql_attn


def save(name, **arrays):
    if ONLY and name not in ONLY:
        print("skipped %s (SPGAN_GOLDEN_ONLY)" % name)
        return
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %8.1f KB  keys=%d" % (name + ".npz", os.path.getsize(path) / 1024, len(arrays)))


def load_state(module, sd):
    module.load_state_dict(OrderedDict((k, v.clone()) for k, v in sd.items()), strict=True)


def grads_of(module, prefix="grad."):
    return {prefix + k: pack(p.grad) for k, p in module.named_parameters() if p.grad is not None}


def buffers_of(module, prefix="buf."):
    return {prefix + k: pack(b) for k, b in module.named_buffers()}


def _to64(sd):
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point():
            v = v.double()
            if "running_" not in k:
                v.requires_grad_(True)
        out[k] = v
    return out


def _rel(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    d = np.abs(a - b)
    return max(d.max() / (np.abs(b).max() + 1e-30), np.sqrt((d ** 2).sum()) / (np.sqrt((b ** 2).sum()) + 1e-30))


def noise_floor_generator(g_sd, o, xg, zg, rg, idx1, idx2, arrs):
    """fp32 rounding noise of the REFERENCE itself: relative distance of its fp32 gradients from an
    fp64 evaluation of the same graph (same neighbour lists).  Tests use max(1e-3, 3 * noise)."""
    sd = _to64(g_sd)
    out = R.generator_forward(sd, xg.double(), zg.double(), o, training=True, idx1=idx1, idx2=idx2)
    (out * rg.double()).sum().backward()
    res = {"noise.out_train": np.float32(_rel(arrs["out_train"], out.detach().numpy()))}
    for k, v in sd.items():
        if v.requires_grad and v.grad is not None and "grad." + k in arrs:
            res["noise.grad." + k] = np.float32(_rel(arrs["grad." + k], pack(v.grad)))
    return res


def _perturbed(sd, rel, seed):
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            out[k] = v * (1.0 + rel * torch.from_numpy(rng.standard_normal(tuple(v.shape)).astype(np.float32)))
        else:
            out[k] = v.clone()
    return out


def sensitivity_generator(g_sd, o, xg, zg, rg, idx1, idx2, rel=1e-6):
    """Conditioning of the reference path itself: relative L2 change of every parameter gradient when
    ALL weights are perturbed by `rel` (1e-6, i.e. ~16 fp32 ulps) with the neighbour lists held fixed.
    LeakyReLU(0.01) masks and max-pool arg-maxes flip under such perturbations, so this is far above
    `rel`; parity tests accept max(1e-3, 3 x this) per tensor (DESIGN.md "Parity")."""
    def grads(sd):
        leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
                for k, v in sd.items()}
        out = R.generator_forward(leaf, xg, zg, o, training=True, idx1=idx1, idx2=idx2)
        (out * rg).sum().backward()
        return {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}
    g0 = grads(g_sd)
    res = {}
    for seed in (101, 102, 103, 104):
        g1 = grads(_perturbed(g_sd, rel, seed))
        for k in g0:
            d = float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-30))
            res["sens.grad." + k] = np.float32(max(d, float(res.get("sens.grad." + k, 0.0))))
    vals = sorted(float(v) for v in res.values() if float(v) < 0.5)
    res["sens.global"] = np.float32(vals[len(vals) // 2])          # median over tensors of the worst draw
    return res


def sensitivity_train_step(o, xg, arrs, Nt, rel=1e-6):
    """Same for the first composed train step: D-phase gradients of D, G-phase gradients of G."""
    tile = lambda a: torch.from_numpy(np.tile(a, (1, Nt, 1)))
    i2d = torch.from_numpy(arrs["s0.idx2_d"].astype(np.int64)).reshape(xg.shape[0], -1)
    i2g = torch.from_numpy(arrs["s0.idx2_g"].astype(np.int64)).reshape(xg.shape[0], -1)

    def grads(gs, ds):
        st = R.TrainState(gs, ds, o)
        R.wgan_gp_train_step(st, xg, tile(arrs["s0.z_d"]), tile(arrs["s0.z_g"]),
                             torch.from_numpy(arrs["s0.data"]).transpose(2, 1), torch.from_numpy(arrs["s0.alpha"]),
                             idx2_d=i2d, idx2_g=i2g)
        return ({k: st.d[k].grad for k in st.d_params if st.d[k].grad is not None},
                {k: st.g[k].grad for k in st.g_params if st.g[k].grad is not None})
    gs, ds = R.synth_state(R.generator_spec(o), 61), R.synth_state(R.discriminator_spec(o), 62)
    d0, g0 = grads(gs, ds)
    res = {}
    for seed in (201, 202, 203, 204):
        d1, g1 = grads(_perturbed(gs, rel, seed), _perturbed(ds, rel, seed + 10))
        for pre, a, b in (("sens.s0.gradD.", d0, d1), ("sens.s0.gradG.", g0, g1)):
            for k in a:
                dd = float((b[k] - a[k]).norm() / (a[k].norm() + 1e-30))
                res[pre + k] = np.float32(max(dd, float(res.get(pre + k, 0.0))))
    vals = sorted(float(v) for v in res.values() if float(v) < 0.5)
    res["sens.global"] = np.float32(vals[len(vals) // 2])
    return res


def sphere(n):
    ball = np.loadtxt(os.path.join(REF, "template/balls/%d.xyz" % n))[:, :3]
    return R.normalize_cloud(ball)       # model.py:46-52 restated


def main():
    torch.set_num_threads(8)
    # importing Generation.modules consumes torch RNG (modules.py:1647-1654): import first, seed after
    from Generation.Generator import Generator, EdgeBlock, AdaptivePointNorm
    from Generation.Discriminator import Discriminator
    from Generation.modules import get_edge_features, edgeConv
    from Common.gradient_penalty import GradientPenalty
    torch.Tensor.cuda = lambda self, *a, **k: self          # gradient_penalty.py:24,32 call .cuda()

    # ------------------------------------------------------------------ kNN fixtures
    ball = sphere(2048)
    np.save(os.path.join(HERE, "sphere_2048.npy"), ball.astype(np.float32))
    xs = torch.Tensor(ball)[None].transpose(2, 1).contiguous()          # [1,3,2048]
    _, idx = get_edge_features(xs, 10, return_idx=True)
    save("knn_sphere2048", idx=idx.view(1, 2048, 10).numpy().astype(np.int16))

    rng = np.random.default_rng(0)
    x_c1 = torch.from_numpy(rng.standard_normal((4, 64, 256)).astype(np.float32))
    ee, idx = get_edge_features(x_c1, 8, return_idx=True)
    save("knn_config1", x=x_c1.numpy(), idx=idx.view(4, 256, 8).numpy().astype(np.int16),
         ee_sub=pack(ee))

    rng = np.random.default_rng(1)
    x_cl = (1.0 + 0.05 * rng.standard_normal((2, 64, 256))).astype(np.float32)   # adversarial: ties
    _, idx = get_edge_features(torch.from_numpy(x_cl), 10, return_idx=True)
    save("knn_clustered", x=x_cl, idx=idx.view(2, 256, 10).numpy().astype(np.int16))

    rng = np.random.default_rng(2)
    for (B, C, N, k) in [(2, 3, 100, 5), (1, 128, 320, 10), (3, 6, 33, 4), (1, 17, 64, 20)]:
        x = torch.from_numpy(rng.standard_normal((B, C, N)).astype(np.float32))
        _, idx = get_edge_features(x, k, return_idx=True)
        save("knn_misc_B%d_C%d_N%d_k%d" % (B, C, N, k), x=x.numpy(),
             idx=idx.view(B, N, k).numpy().astype(np.int16))

    # ------------------------------------------------------------------ config 1 blocks
    rng = np.random.default_rng(10)
    r_out = torch.from_numpy(rng.standard_normal((4, 64, 256)).astype(np.float32))
    for name, ctor, spec in (("edgeblock", lambda: EdgeBlock(64, 64, 8), R.edge_block_spec("", 64, 64, 8)),
                             ("edgeconv", lambda: edgeConv(64, 64, 8), R.edge_conv_spec("", 64, 64))):
        m = ctor()
        load_state(m, R.synth_state(spec, 11))
        m.train()
        x = x_c1.clone().requires_grad_(True)
        out = m(x)
        (out * r_out).sum().backward()
        arrs = {"out_train": out.detach().numpy(), "grad_x": x.grad.numpy(), "r_out": r_out.numpy()}
        arrs.update(grads_of(m))
        arrs.update(buffers_of(m))
        m.eval()
        with torch.no_grad():
            arrs["out_eval"] = m(x_c1).numpy()
        save("config1_" + name, **arrs)

    # ------------------------------------------------------------------ AdaptivePointNorm
    rng = np.random.default_rng(20)
    m = AdaptivePointNorm(64, 128)
    sd = R.synth_state(R._conv("style", (128, 128, 1)), 21)
    load_state(m, sd)
    xa = torch.from_numpy(rng.standard_normal((3, 64, 200)).astype(np.float32)).requires_grad_(True)
    sa = torch.from_numpy(rng.standard_normal((3, 128, 200)).astype(np.float32)).requires_grad_(True)
    ra = torch.from_numpy(rng.standard_normal((3, 64, 200)).astype(np.float32))
    out = m(xa, sa)
    (out * ra).sum().backward()
    arrs = {"x": xa.detach().numpy(), "style": sa.detach().numpy(), "r": ra.numpy(),
            "out": out.detach().numpy(), "grad_x": xa.grad.numpy(), "grad_style": sa.grad.numpy()}
    arrs.update(grads_of(m))
    save("adain", **arrs)

    # ------------------------------------------------------------------ Discriminator
    opts = R.default_opts()
    rng = np.random.default_rng(30)
    B, N = 4, 256
    d_sd = R.synth_state(R.discriminator_spec(opts), 31)
    D = Discriminator(opts)
    load_state(D, d_sd)
    D.train()
    pts = (0.5 * rng.standard_normal((B, N, 3))).astype(np.float32)
    xd = torch.from_numpy(pts).transpose(2, 1).requires_grad_(True)          # non-contiguous [B,3,N]
    rd = torch.from_numpy(rng.standard_normal((B, 1)).astype(np.float32))
    out = D(xd)
    (out * rd).sum().backward()
    arrs = {"pts": pts, "r": rd.numpy(), "out_train": out.detach().numpy(), "grad_x": xd.grad.numpy()}
    arrs.update(grads_of(D))
    arrs.update(buffers_of(D))
    D.eval()
    with torch.no_grad():
        arrs["out_eval"] = D(torch.from_numpy(pts).transpose(2, 1)).numpy()
    save("discriminator", **arrs)

    # small_d variant: forward only
    opts_s = R.default_opts(small_d=True)
    Ds = Discriminator(opts_s)
    load_state(Ds, R.synth_state(R.discriminator_spec(opts_s), 32))
    Ds.train()
    save("discriminator_small", out_train=Ds(torch.from_numpy(pts).transpose(2, 1)).detach().numpy())

    # ------------------------------------------------------------------ GradientPenalty
    D = Discriminator(opts)
    load_state(D, d_sd)
    D.train()
    rng = np.random.default_rng(40)
    real = torch.from_numpy((0.5 * rng.standard_normal((B, 3, N))).astype(np.float32)).requires_grad_(True)
    fake = torch.from_numpy((0.5 * rng.standard_normal((B + 2, 3, N))).astype(np.float32))   # fake[:B] slice
    torch.manual_seed(41)
    alpha = torch.rand(B, 1, 1)
    torch.manual_seed(41)
    gp = GradientPenalty(10, gamma=1)(D, real, fake)
    gp.backward()
    arrs = {"real": real.detach().numpy(), "fake": fake.numpy(), "alpha": alpha.numpy(),
            "gp": np.float32(gp.item()), "grad_real": real.grad.numpy()}
    arrs.update(grads_of(D))
    arrs.update(buffers_of(D))
    save("gradient_penalty", **arrs)

    # ------------------------------------------------------------------ Generator
    ball256 = sphere(256).astype(np.float32)
    for tag, kw in (("default", {}), ("off_znorm", {"off": True, "z_norm": True}), ("use_head", {"use_head": True}),
                    ("eql_attn", {"eql": True, "attn": True})):
        o = R.default_opts(np=256, **kw)
        rng = np.random.default_rng(50)
        Bg = 4
        g_sd = R.synth_state(R.generator_spec(o), 51)
        G = Generator(o)
        load_state(G, g_sd)
        G.train()
        xg = torch.from_numpy(np.tile(ball256[None], (Bg, 1, 1)))
        zg = torch.from_numpy(R.latent_noise(rng, Bg, 256, o.nz))
        rg = torch.from_numpy(rng.standard_normal((Bg, 3, 256)).astype(np.float32))
        feats = {}
        h1 = G.adain1.register_forward_hook(lambda m, i, out: feats.__setitem__("x1", out.detach()))
        h2 = G.EdgeConv2.register_forward_hook(lambda m, i, out: feats.__setitem__("e2", out.detach()))
        out = G(xg, zg)
        h1.remove(); h2.remove()
        (out * rg).sum().backward()
        _, idx1 = get_edge_features(xg.transpose(2, 1).contiguous(), o.nk // 2, return_idx=True)
        arrs = {"z": zg.numpy()[:, :1].copy(), "r": rg.numpy(), "out_train": out.detach().numpy(),
                "x1": feats["x1"].numpy(), "edgeconv2_out": pack(feats["e2"])}
        if tag == "eql_attn":
            # non-default flags: forward + parameter gradients, EdgeConv2's list for injection
            _, idx2 = get_edge_features(feats["x1"], o.nk // 2, return_idx=True)
            arrs["idx2"] = idx2.view(Bg, 256, -1).numpy().astype(np.int16)
            arrs.update(grads_of(G))
        if tag == "default":
            _, idx2 = get_edge_features(feats["x1"], o.nk // 2, return_idx=True)
            arrs["idx1"] = idx1.view(Bg, 256, -1).numpy().astype(np.int16)
            arrs["idx2"] = idx2.view(Bg, 256, -1).numpy().astype(np.int16)
            arrs.update(grads_of(G))
            arrs.update(buffers_of(G))
            arrs.update(noise_floor_generator(g_sd, o, xg, zg, rg, idx1, idx2, arrs))
            arrs.update(sensitivity_generator(g_sd, o, xg, zg, rg, idx1, idx2))
            G.eval()
            with torch.no_grad():
                arrs["out_eval"] = G(xg, zg).numpy()
                sel = np.zeros(256, np.int64); sel[64:160] = 1
                z2 = torch.from_numpy(R.latent_noise(rng, Bg, 256, o.nz))
                arrs["z2"] = z2.numpy()[:, :1].copy()
                arrs["selection"] = sel
                arrs["interp_z"] = G.interpolate(xg, zg.clone(), z2, torch.from_numpy(sel), 0.3).numpy()
                arrs["interp_latent"] = G.interpolate(xg, zg.clone(), z2, torch.from_numpy(sel), 0.3,
                                                      use_latent=True).numpy()
        save("generator_" + tag, **arrs)

    # ------------------------------------------------------------------ composed WGAN-GP train step
    from Common.network_utils import requires_grad
    o = R.default_opts(np=256)
    Bt, Nt = 4, 256
    G, D = Generator(o), Discriminator(o)
    load_state(G, R.synth_state(R.generator_spec(o), 61))
    load_state(D, R.synth_state(R.discriminator_spec(o), 62))
    G.train(); D.train()
    optG = torch.optim.Adam(filter(lambda p: p.requires_grad, G.parameters()), lr=1e-4, betas=(0.5, 0.99))
    optD = torch.optim.Adam(filter(lambda p: p.requires_grad, D.parameters()), lr=1e-4, betas=(0.5, 0.99))
    rng = np.random.default_rng(60)
    xg = torch.from_numpy(np.tile(ball256[None], (Bt, 1, 1)))
    gp_fn = GradientPenalty(10, gamma=1)
    arrs = {}
    x1_log = []
    hook = G.adain1.register_forward_hook(lambda m, i, out: x1_log.append(out.detach()))
    for step in range(2):
        data = torch.from_numpy(R.synthetic_chairs(rng, Bt, Nt))
        z_d = torch.from_numpy(R.latent_noise(rng, Bt, Nt, o.nz))
        z_g = torch.from_numpy(R.latent_noise(rng, Bt, Nt, o.nz))
        torch.manual_seed(600 + step)
        alpha = torch.rand(Bt, 1, 1)
        # ---- model.py:239-260 (+ GP) ----
        requires_grad(G, False); requires_grad(D, True)
        optD.zero_grad()
        real = data.clone().requires_grad_(True)
        fake = G(xg, z_d)
        real = real.transpose(2, 1)
        fake = fake.detach()
        d_real, d_fake = D(real), D(fake)
        torch.manual_seed(600 + step)
        gp = gp_fn(D, real, fake)
        lossD = (d_fake.mean() - d_real.mean()) + gp
        lossD.backward()
        if step == 0:
            arrs.update(grads_of(D, "s0.gradD."))
        optD.step()
        # ---- model.py:264-279 ----
        requires_grad(G, True); requires_grad(D, False)
        optG.zero_grad()
        fake = G(xg, z_g)
        _ = D(real)
        lossG = -D(fake).mean()
        lossG.backward()
        if step == 0:
            arrs.update(grads_of(G, "s0.gradG."))
        optG.step()
        for name, x1 in zip(("d", "g"), x1_log[-2:]):      # EdgeConv2 neighbour lists of the two G forwards
            _, i2 = get_edge_features(x1, o.nk // 2, return_idx=True)
            arrs["s%d.idx2_%s" % (step, name)] = i2.view(Bt, Nt, -1).numpy().astype(np.int16)
        arrs["s%d.data" % step] = data.numpy()
        arrs["s%d.z_d" % step] = z_d.numpy()[:, :1].copy()
        arrs["s%d.z_g" % step] = z_g.numpy()[:, :1].copy()
        arrs["s%d.alpha" % step] = alpha.numpy()
        arrs["s%d.loss_d" % step] = np.float32(lossD.item())
        arrs["s%d.gp" % step] = np.float32(gp.item())
        arrs["s%d.loss_g" % step] = np.float32(lossG.item())
        print("train step", step, lossD.item(), gp.item(), lossG.item())
    hook.remove()
    arrs.update(sensitivity_train_step(o, xg, arrs, Nt))
    arrs.update(buffers_of(G, "end.bufG."))
    arrs.update(buffers_of(D, "end.bufD."))
    arrs.update({"end.G." + k: pack(p) for k, p in G.named_parameters()})
    arrs.update({"end.D." + k: pack(p) for k, p in D.named_parameters()})
    save("train_step", **arrs)
    np.save(os.path.join(HERE, "sphere_256.npy"), ball256)


if __name__ == "__main__":
    main()
