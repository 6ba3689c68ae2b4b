"""GPU parity of the drop-in modules (Generator, Discriminator, EdgeBlock, edgeConv,
AdaptivePointNorm, GradientPenalty, composed train step) against the golden vectors produced
by the unmodified reference (tests/golden/make_golden.py) and against the CPU oracle run live.
Bar (BASELINE.json): kNN indices bit-exact, features / losses / gradients within 1e-3 relative."""
import numpy as np
import pytest
import torch

from conftest import assert_rel, golden
from oracle import spgan_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _pkg():
    import spgan_b200
    return spgan_b200


def _load(module, spec, seed):
    module.load_state_dict(R.synth_state(spec, seed), strict=True)
    return module.cuda()


def _check_grads(module, g, prefix="grad.", tol=TOL, max_tol=None):
    """Per-parameter relative check.  Where the golden file carries `sens.<key>` -- the relative change
    of the REFERENCE's own gradient under a 1e-6 relative perturbation of the weights (LeakyReLU masks
    and max-pool arg-maxes flip; measured by tests/golden/make_golden.py) -- the bar for that tensor is
    max(tol, 3 x max(sens, sens.global)): parity cannot be tighter than the reference's own conditioning
    (single mask flips are heavy-tailed, hence the file-wide median `sens.global` as a floor).  Gradients that are mathematically zero (a conv bias feeding a
    train-mode BatchNorm) are rounding noise in the reference too (~1e-6 of the layer's scale): they
    are compared against the largest gradient magnitude of the module instead of their own."""
    from conftest import pack_like_golden, rel_err
    scale = max(float(np.abs(v).max()) for kk, v in g.items() if kk.startswith(prefix))
    n = 0
    for k, p in module.named_parameters():
        if prefix + k not in g:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        ref = g[prefix + k]
        if p.grad is None:
            # a bias feeding a train-mode BatchNorm: mathematically zero gradient, the product returns none at all
            assert float(np.abs(ref).max()) < 1e-4 * scale, (k, "missing gradient")
            n += 1
            continue
        if float(np.abs(ref).max()) < 1e-4 * scale:
            got = pack_like_golden(p.grad)
            assert float(np.abs(got - ref).max()) <= tol * 1e-2 * scale, (k, "noise-level gradient too large")
        else:
            sens = max(float(g.get("sens." + prefix + k, 0.0)), float(g.get("sens.global", 0.0)))
            t = min(max(tol, 3.0 * sens), 0.25)
            assert_rel(p.grad, ref, t, prefix + k, max_tol=max(10 * t, max_tol or 0.0))
        n += 1
    assert n > 0


def _check_bufs(module, g, prefix="buf.", tol=TOL):
    for k, b in module.named_buffers():
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(g[prefix + k]), k
        else:
            assert_rel(b, g[prefix + k], tol, prefix + k)


def _scalar_loss(out, r):
    ops = _pkg().ops
    return ops.MeanScale.apply(ops.Mul.apply(out, r), float(r.numel()))


@pytest.mark.parametrize("name", ["edgeblock", "edgeconv"])
def test_config1_blocks(name):
    """BASELINE config 1: single EdgeConv block, B=4 N=256 k=8 C=64."""
    pkg = _pkg()
    g = golden("config1_" + name)
    x0 = torch.from_numpy(golden("knn_config1")["x"])
    if name == "edgeblock":
        m = _load(pkg.EdgeBlock(64, 64, 8), R.edge_block_spec("", 64, 64, 8), 11)
    else:
        m = _load(pkg.edgeConv(64, 64, 8), R.edge_conv_spec("", 64, 64), 11)
    m.train()
    x = x0.cuda().requires_grad_()
    out = m(x)
    assert tuple(out.shape) == (4, 64, 256)
    _scalar_loss(out, torch.from_numpy(g["r_out"]).cuda()).backward()
    assert_rel(out, g["out_train"], TOL, "out_train")
    assert_rel(x.grad, g["grad_x"], TOL, "grad_x")
    _check_grads(m, g)
    _check_bufs(m, g)
    m.eval()
    with torch.no_grad():
        assert_rel(m(x0.cuda()), g["out_eval"], TOL, "out_eval")


def test_adain():
    pkg = _pkg()
    g = golden("adain")
    m = _load(pkg.AdaptivePointNorm(64, 128), R._conv("style", (128, 128, 1)), 21)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_()
    s = torch.from_numpy(g["style"]).cuda().requires_grad_()
    out = m(x, s)
    _scalar_loss(out, torch.from_numpy(g["r"]).cuda()).backward()
    assert_rel(out, g["out"], TOL)
    assert_rel(x.grad, g["grad_x"], TOL)
    assert_rel(s.grad, g["grad_style"], TOL)
    _check_grads(m, g)


def test_discriminator():
    pkg = _pkg()
    g = golden("discriminator")
    D = _load(pkg.Discriminator(R.default_opts()), R.discriminator_spec(R.default_opts()), 31)
    D.train()
    x = torch.from_numpy(g["pts"]).cuda().transpose(2, 1).requires_grad_()       # strided view, as model.py:249
    out = D(x)
    assert tuple(out.shape) == (4, 1)
    _scalar_loss(out, torch.from_numpy(g["r"]).cuda()).backward()
    assert_rel(out, g["out_train"], TOL, "out_train")
    assert_rel(x.grad, g["grad_x"], TOL, "grad_x")
    _check_grads(D, g)
    _check_bufs(D, g)
    D.eval()
    with torch.no_grad():
        assert_rel(D(x.detach()), g["out_eval"], TOL, "out_eval")
    Ds = _load(pkg.Discriminator(R.default_opts(small_d=True)), R.discriminator_spec(R.default_opts(small_d=True)), 32)
    Ds.train()
    assert_rel(Ds(x.detach()), golden("discriminator_small")["out_train"], TOL)


def test_gradient_penalty_double_backward():
    pkg = _pkg()
    g = golden("gradient_penalty")
    D = _load(pkg.Discriminator(R.default_opts()), R.discriminator_spec(R.default_opts()), 31)
    D.train()
    real = torch.from_numpy(g["real"]).cuda()
    fake = torch.from_numpy(g["fake"]).cuda()                     # B+2 clouds: exercises fake[:B]
    gp = pkg.GradientPenalty(10, gamma=1)(D, real, fake, alpha=torch.from_numpy(g["alpha"]))
    gp.backward()
    ref = float(g["gp"])
    assert abs(float(gp) - ref) <= TOL * abs(ref), (float(gp), ref)
    _check_grads(D, g)
    _check_bufs(D, g)


def test_gradient_penalty_on_an_eval_mode_critic():
    """GradientPenalty(D.eval(), ...) (frozen BatchNorm statistics): the reference supports it through plain autograd;
    here the eval-mode BatchNorm under the double backward is composed from the twice-differentiable operator set.
    Checked against the live oracle (torch CPU), value and every parameter gradient."""
    pkg = _pkg()
    o = R.default_opts(np=256)
    spec = R.discriminator_spec(o)
    state = R.synth_state(spec, 33)
    for k in state:                                              # non-trivial running statistics
        if k.endswith("running_mean"):
            state[k] = 0.1 * torch.randn_like(state[k])
        if k.endswith("running_var"):
            state[k] = 0.5 + torch.rand_like(state[k])
    D = pkg.Discriminator(o)
    D.load_state_dict(state)
    D = D.cuda().eval()
    rng = np.random.default_rng(5)
    real = torch.from_numpy((0.5 * rng.standard_normal((2, 3, 256))).astype(np.float32))
    fake = torch.from_numpy((0.5 * rng.standard_normal((2, 3, 256))).astype(np.float32))
    alpha = torch.rand(2, 1, 1, generator=torch.Generator().manual_seed(1))
    gp = pkg.GradientPenalty(10)(D, real.cuda(), fake.cuda(), alpha=alpha)
    gp.backward()
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in state.items()}
    gp_ref = R.gradient_penalty(lambda t: R.discriminator_forward(sd, t, False), real, fake, alpha)
    gp_ref.backward()
    assert abs(float(gp) - float(gp_ref)) <= TOL * abs(float(gp_ref)), (float(gp), float(gp_ref))
    scale = max(float(v.grad.abs().max()) for v in sd.values() if v.requires_grad and v.grad is not None)
    n = 0
    for name, p in D.named_parameters():
        ref = sd[name].grad
        if ref is None or p.grad is None:
            continue
        err = float((p.grad.cpu() - ref.reshape(p.grad.shape)).abs().max())
        assert err <= TOL * max(float(ref.abs().max()), 1e-2 * scale), (name, err)
        n += 1
    assert n >= 10
    for name, b in D.named_buffers():                            # eval mode: the buffers do not move
        assert torch.equal(b.cpu(), state[name]), name


@pytest.mark.parametrize("tag,kw", [("default", {}), ("off_znorm", {"off": True, "z_norm": True}),
                                    ("use_head", {"use_head": True}), ("eql_attn", {"eql": True, "attn": True})])
def test_generator(tag, kw, sphere256):
    pkg = _pkg()
    g = golden("generator_" + tag)
    o = R.default_opts(np=256, **kw)
    G = _load(pkg.Generator(o), R.generator_spec(o), 51)
    G.train()
    Bg = g["out_train"].shape[0]
    x = torch.from_numpy(np.tile(sphere256[None], (Bg, 1, 1))).cuda()
    z = torch.from_numpy(np.tile(g["z"], (1, 256, 1))).cuda()
    if tag == "eql_attn":
        G.debug_idx = (None, torch.from_numpy(g["idx2"].astype(np.int32)).cuda())
    if tag == "default":
        # kNN indices through the generator must be bit-exact where the input features are: EdgeConv1
        # sees the sphere itself; for EdgeConv2 inject the reference's own list so that feature parity is
        # not polluted by near-tie flips from ~1e-6 feature differences.
        G.debug_idx = (None, torch.from_numpy(g["idx2"].astype(np.int32)).cuda())
    out = G(x, z)
    assert tuple(out.shape) == (Bg, 3, 256)
    if tag == "default":
        idx1 = G._graph_cache[1].cpu().numpy()
        assert np.array_equal(idx1, g["idx1"].astype(np.int32))
        assert_rel(G._last_x1.view(Bg, 256, 64).permute(0, 2, 1), g["x1"], TOL, "x1")
    assert_rel(out, g["out_train"], TOL, "out_train")
    if tag == "eql_attn":
        # non-default flags: same state_dict keys (head.N.conv.weight_orig, attn.*), forward and
        # parameter gradients against the reference; max_tol as for the default generator (mask flips)
        assert "head.0.conv.weight_orig" in G.state_dict() and "attn.gamma" in G.state_dict()
        _scalar_loss(out, torch.from_numpy(g["r"]).cuda()).backward()
        _check_grads(G, g, tol=5e-3)
    if tag != "default":
        return
    _scalar_loss(out, torch.from_numpy(g["r"]).cuda()).backward()
    _check_grads(G, g)
    _check_bufs(G, g)
    # free-running kNN on our own features: flips only at near ties
    G.debug_idx = None
    G2 = _load(pkg.Generator(o), R.generator_spec(o), 51)
    G2.train()
    out2 = G2(x, z)
    flips = float((out2 - out).abs().max())
    assert flips < 0.2, flips
    G.eval()
    gx = golden("generator_extra")                     # the reference's EdgeConv2 list of ITS eval-mode forward
    G.debug_idx = (None, torch.from_numpy(gx["idx2_eval"].astype(np.int32)).cuda())
    with torch.no_grad():
        out_eval = G(x, z)
    assert_rel(G._last_x1.view(Bg, 256, 64).permute(0, 2, 1), gx["x1_eval"], TOL, "x1_eval")
    assert_rel(out_eval, g["out_eval"], TOL, "out_eval")
    # op-level: our kNN on the reference's own eval-mode features reproduces its list bit for bit
    idx_own = pkg.ops.knn_indices(torch.from_numpy(gx["x1_eval"]).cuda(), G.nk).cpu().numpy()
    assert np.array_equal(idx_own, gx["idx2_eval"].astype(np.int32))
    G.debug_idx = (None, None)                         # free-running: own kNN on own features, near-tie flips only
    with torch.no_grad():
        out_free = G(x, z)
    assert float(np.abs(out_free.cpu().numpy() - g["out_eval"]).max()) < 0.05


def test_generator_broadcast_latent_equals_tiled(sphere256):
    pkg = _pkg()
    o = R.default_opts(np=256)
    G = _load(pkg.Generator(o), R.generator_spec(o), 51)
    G.eval()
    x = torch.from_numpy(np.tile(sphere256[None], (2, 1, 1))).cuda()
    zv = torch.from_numpy(golden("generator_default")["z"]).cuda()[:2]      # [2,1,128]
    with torch.no_grad():
        a = G(x, zv.expand(2, 256, 128))
        b = G(x, zv.repeat(1, 256, 1))
    assert torch.equal(a, b)


def test_generator_interpolate(sphere256):
    pkg = _pkg()
    g = golden("generator_default")
    o = R.default_opts(np=256)
    G = _load(pkg.Generator(o), R.generator_spec(o), 51)
    Bg = g["out_train"].shape[0]
    x = torch.from_numpy(np.tile(sphere256[None], (Bg, 1, 1))).cuda()
    z = torch.from_numpy(np.tile(g["z"], (1, 256, 1))).cuda()
    z2 = torch.from_numpy(np.tile(g["z2"], (1, 256, 1))).cuda()
    G.train()
    with torch.no_grad():
        G(x, z)                          # the golden was taken after one train-mode forward (BN buffers)
    G.eval()
    sel = torch.from_numpy(g["selection"]).cuda()
    gx = golden("generator_extra")
    # the reference's own EdgeConv2 neighbour lists of these two calls are injected, so that a blending
    # bug cannot hide behind near-tie flips: features and output at the 1e-3 bar
    with torch.no_grad():
        G.debug_idx = (None, torch.from_numpy(gx["idx2_interp_z"].astype(np.int32)).cuda())
        z1 = z.clone()
        a = G.interpolate(x, z1, z2, sel, 0.3)
        assert_rel(G._last_x1.view(Bg, 256, 64).permute(0, 2, 1), gx["x1_interp_z"], TOL, "x1 (interp z)")
        assert_rel(a, g["interp_z"], TOL, "interp_z")
        # like the reference, the latent branch writes the blend into its z1 argument (Generator.py:205-206)
        want = z.clone()
        want[:, sel == 1] = z[:, sel == 1] * 0.7 + z2[:, sel == 1] * 0.3
        assert torch.allclose(z1, want, rtol=0, atol=1e-6)
        G.debug_idx = (None, torch.from_numpy(gx["idx2_interp_latent"].astype(np.int32)).cuda())
        b = G.interpolate(x, z.clone(), z2, sel, 0.3, use_latent=True)
        assert_rel(G._last_x1.view(Bg, 256, 64).permute(0, 2, 1), gx["x1_interp_latent"], TOL, "x1 (interp latent)")
        assert_rel(b, g["interp_latent"], TOL, "interp_latent")
        G.debug_idx = None                             # free-running kNN stays within the near-tie flip bound
        c = G.interpolate(x, z.clone(), z2, sel, 0.3)
    assert float(np.abs(c.cpu().numpy() - g["interp_z"]).max()) < 0.05


def test_train_step_against_reference_golden(sphere256):
    """Two composed WGAN-GP steps (B=4, N=256): losses, first-step gradients, BN buffers."""
    pkg = _pkg()
    g = golden("train_step")
    o = R.default_opts(np=256)
    G = _load(pkg.Generator(o), R.generator_spec(o), 61)
    D = _load(pkg.Discriminator(o), R.discriminator_spec(o), 62)
    G.train(); D.train()
    tr = pkg.WGANGPTrainer(G, D)
    x = torch.from_numpy(np.tile(sphere256[None], (4, 1, 1))).cuda()
    for step in range(2):
        tile = lambda a: torch.from_numpy(np.tile(a, (1, 256, 1))).cuda()
        real = torch.from_numpy(g["s%d.data" % step]).cuda().transpose(2, 1)
        alpha = torch.from_numpy(g["s%d.alpha" % step])
        # inject the reference's EdgeConv2 neighbour lists: the step is chaotic w.r.t. near-tie flips
        # (fp32 vs fp64 evaluation of the reference itself moves gp by 3.6 %), see DESIGN.md
        G.debug_idx = (None, torch.from_numpy(g["s%d.idx2_d" % step].astype(np.int32)).cuda())
        loss_d, gp = tr.d_phase(x, tile(g["s%d.z_d" % step]), real, alpha)
        if step == 0:
            _check_grads(D, g, "s0.gradD.")
        G.debug_idx = (None, torch.from_numpy(g["s%d.idx2_g" % step].astype(np.int32)).cuda())
        loss_g = tr.g_phase(x, tile(g["s%d.z_g" % step]), real)
        if step == 0:
            _check_grads(G, g, "s0.gradG.")
        for key, val in (("loss_d", loss_d), ("gp", gp), ("loss_g", loss_g)):
            ref = float(g["s%d.%s" % (step, key)])
            # step 0 is a pure forward of the initial weights.  Step 1 follows one Adam update whose very first
            # step moves every weight by +-lr * sign(grad) regardless of |grad|: parameters with noise-level
            # gradients get rounding-dependent signs, so the second step is only loosely comparable.
            tol = 2e-3 if step == 0 else 0.1
            assert abs(float(val) - ref) <= tol * max(1.0, abs(ref)), (step, key, float(val), ref)
    for k, b in D.named_buffers():
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(g["end.bufD." + k])
    for k, b in G.named_buffers():
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(g["end.bufG." + k])


def test_modules_fail_loudly_on_cpu_tensors():
    pkg = _pkg()
    D = pkg.Discriminator(R.default_opts())
    with pytest.raises(RuntimeError, match="CUDA"):
        D(torch.zeros(2, 3, 64))


def test_train_step_graph_replay_equals_eager(sphere256):
    """WGANGPTrainer.capture/replay (one CUDA graph per step) against the eager trainer: same kernels, same
    inputs => the same losses and parameters after every step, and the Adam step count advances on the device."""
    pkg = _pkg()
    o = R.default_opts(np=256)
    B, N = 2, 256
    rng = np.random.default_rng(9)
    x = torch.from_numpy(np.tile(sphere256[None], (B, 1, 1))).cuda()
    batches = [dict(z_d=torch.from_numpy(R.latent_noise(rng, B, N, o.nz)[:, :1].copy()).cuda(),
                    z_g=torch.from_numpy(R.latent_noise(rng, B, N, o.nz)[:, :1].copy()).cuda(),
                    real=torch.from_numpy(R.synthetic_chairs(rng, B, N)).cuda(),
                    alpha=torch.rand(B, 1, 1, generator=torch.Generator().manual_seed(i)).cuda()) for i in range(4)]
    runs = []
    for graph in (False, True):
        G = _load(pkg.Generator(o), R.generator_spec(o), 61).train()
        D = _load(pkg.Discriminator(o), R.discriminator_spec(o), 62).train()
        tr = pkg.WGANGPTrainer(G, D)
        call = lambda d: tr.step(x, d["z_d"].expand(B, N, o.nz), d["z_g"].expand(B, N, o.nz), d["real"].transpose(2, 1),
                                 alpha=d["alpha"])
        out = []
        if graph:
            d = batches[0]
            tr.capture(x, d["z_d"].expand(B, N, o.nz), d["z_g"].expand(B, N, o.nz), d["real"].transpose(2, 1), d["alpha"],
                       warmup=0)
            assert tr.graph_launches > 100
            for d in batches:
                res = tr.replay(z_d=d["z_d"], z_g=d["z_g"], real=d["real"].transpose(2, 1), alpha=d["alpha"])
                out.append([float(t) for t in res])
        else:
            for d in batches:
                out.append([float(t) for t in call(d)])
        runs.append((out, tr.opt_d.t, tr.opt_g.t, tr.opt_d.buf.flat_p.clone(), tr.opt_g.buf.flat_p.clone()))
    (eo, edt, egt, edp, egp), (go, gdt, ggt, gdp, ggp) = runs
    assert edt == gdt == 4 and egt == ggt == 4
    # fp32 atomics in the weight-gradient kernels make runs differ at rounding level; Adam's first steps turn any
    # difference into +-lr per element, so parameters agree to a few lr and losses to the usual tolerance
    # (the first step starts from identical weights: tight; later ones inherit the chaotic divergence documented in
    # DESIGN.md "Parity" -- the reference itself is only reproducible to ~10 % there)
    for step, (a, b) in enumerate(zip(eo, go)):
        tol = 1e-3 if step == 0 else (0.05 if step == 1 else None)
        for va, vb in zip(a, b):
            assert np.isfinite(va) and np.isfinite(vb)
            if tol is not None:
                assert abs(va - vb) <= tol * max(1.0, abs(va)), (step, eo, go)
    assert float((edp - gdp).abs().max()) <= 1e-3 and float((egp - ggp).abs().max()) <= 1e-3
