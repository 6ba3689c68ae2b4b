"""Pins oracle/spgan_ref.py (torch-CPU functional restatement) against tests/golden/*.npz,
i.e. against outputs of the unmodified reference modules (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import assert_rel, golden
from oracle import spgan_ref as R

TOL = 2e-5       # same torch ops, different graph shape (gather vs index_select loop): ~1e-6 expected


def _leaf(sd):
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
        out[k] = v
    return out


def _check_grads(sd, g, prefix="grad.", tol=TOL):
    n = 0
    for k, v in sd.items():
        if v.requires_grad:
            if prefix + k not in g:      # the reference produced no gradient for this parameter
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
                continue
            assert_rel(v.grad, g[prefix + k], tol, prefix + k)
            n += 1
    assert n > 0


def _check_bufs(sd, g, prefix="buf.", tol=TOL):
    for k, v in sd.items():
        if "running_" in k:
            assert_rel(v, g[prefix + k], tol, prefix + k)
        elif k.endswith("num_batches_tracked"):
            assert int(v) == int(g[prefix + k]), k


@pytest.mark.parametrize("name", ["edgeblock", "edgeconv"])
def test_config1_blocks(name):
    g = golden("config1_" + name)
    x0 = torch.from_numpy(golden("knn_config1")["x"])
    spec = R.edge_block_spec("", 64, 64, 8) if name == "edgeblock" else R.edge_conv_spec("", 64, 64)
    fn = R.edge_block if name == "edgeblock" else R.edge_conv
    sd = _leaf(R.synth_state(spec, 11))
    x = x0.clone().requires_grad_(True)
    out = fn(sd, x, 8, training=True)
    (out * torch.from_numpy(g["r_out"])).sum().backward()
    assert_rel(out, g["out_train"], TOL, "out_train")
    assert_rel(x.grad, g["grad_x"], TOL, "grad_x")
    _check_grads(sd, g)
    _check_bufs(sd, g)
    with torch.no_grad():
        assert_rel(fn(sd, x0, 8, training=False), g["out_eval"], TOL, "out_eval")


def test_adain():
    g = golden("adain")
    sd = _leaf(R.synth_state(R._conv("style", (128, 128, 1)), 21))
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    s = torch.from_numpy(g["style"]).requires_grad_(True)
    out = R.adaptive_point_norm(sd, x, s)
    (out * torch.from_numpy(g["r"])).sum().backward()
    assert_rel(out, g["out"], TOL, "out")
    assert_rel(x.grad, g["grad_x"], TOL, "grad_x")
    assert_rel(s.grad, g["grad_style"], TOL, "grad_style")
    _check_grads(sd, g)


def test_discriminator():
    g = golden("discriminator")
    opts = R.default_opts()
    sd = _leaf(R.synth_state(R.discriminator_spec(opts), 31))
    x = torch.from_numpy(g["pts"]).transpose(2, 1).requires_grad_(True)
    out = R.discriminator_forward(sd, x, True)
    (out * torch.from_numpy(g["r"])).sum().backward()
    assert_rel(out, g["out_train"], TOL, "out_train")
    assert_rel(x.grad, g["grad_x"], TOL, "grad_x")
    _check_grads(sd, g)
    _check_bufs(sd, g)
    with torch.no_grad():
        assert_rel(R.discriminator_forward(sd, x.detach(), False), g["out_eval"], TOL, "out_eval")
    opts_s = R.default_opts(small_d=True)
    sds = R.synth_state(R.discriminator_spec(opts_s), 32)
    assert_rel(R.discriminator_forward(sds, x.detach(), True), golden("discriminator_small")["out_train"], TOL)


def test_gradient_penalty():
    g = golden("gradient_penalty")
    sd = _leaf(R.synth_state(R.discriminator_spec(R.default_opts()), 31))
    real = torch.from_numpy(g["real"]).requires_grad_(True)
    gp = R.gradient_penalty(lambda t: R.discriminator_forward(sd, t, True), real,
                            torch.from_numpy(g["fake"]), torch.from_numpy(g["alpha"]), 10.0, 1.0)
    gp.backward()
    assert abs(float(gp.detach()) - float(g["gp"])) <= TOL * abs(float(g["gp"]))
    assert_rel(real.grad, g["grad_real"], 1e-4, "grad_real")
    _check_grads(sd, g, tol=1e-4)
    _check_bufs(sd, g)


@pytest.mark.parametrize("tag,kw", [("default", {}), ("off_znorm", {"off": True, "z_norm": True}),
                                    ("use_head", {"use_head": True}), ("eql_attn", {"eql": True, "attn": True})])
def test_generator(tag, kw, sphere256):
    g = golden("generator_" + tag)
    o = R.default_opts(np=256, **kw)
    sd = _leaf(R.synth_state(R.generator_spec(o), 51))
    x = torch.from_numpy(np.tile(sphere256[None], (g["out_train"].shape[0], 1, 1)))
    z = torch.from_numpy(np.tile(g["z"], (1, 256, 1)))
    out, x1 = R.generator_forward(sd, x, z, o, training=True, return_x1=True)
    assert_rel(out, g["out_train"], 1e-4, "out_train")
    assert_rel(x1, g["x1"], 1e-4, "x1")
    if tag == "eql_attn":                  # non-default flags: equalised-lr weights + the N x N attention block
        (out * torch.from_numpy(g["r"])).sum().backward()
        _check_grads(sd, g, tol=5e-4)
    if tag != "default":
        return
    (out * torch.from_numpy(g["r"])).sum().backward()
    _check_grads(sd, g, tol=2e-4)
    _check_bufs(sd, g, tol=1e-4)
    with torch.no_grad():
        assert_rel(R.generator_forward(sd, x, z, o, training=False), g["out_eval"], 1e-4, "out_eval")
        z2 = torch.from_numpy(np.tile(g["z2"], (1, 256, 1)))
        sel = torch.from_numpy(g["selection"])
        assert_rel(R.generator_interpolate(sd, x, z.clone(), z2, sel, 0.3, o), g["interp_z"], 1e-4)
        assert_rel(R.generator_interpolate(sd, x, z.clone(), z2, sel, 0.3, o, use_latent=True),
                   g["interp_latent"], 1e-4)


def test_train_step(sphere256):
    g = golden("train_step")
    o = R.default_opts(np=256)
    st = R.TrainState(R.synth_state(R.generator_spec(o), 61), R.synth_state(R.discriminator_spec(o), 62), o)
    x = torch.from_numpy(np.tile(sphere256[None], (4, 1, 1)))
    for step in range(2):
        tile = lambda a: torch.from_numpy(np.tile(a, (1, 256, 1)))
        real = torch.from_numpy(g["s%d.data" % step]).transpose(2, 1)
        out = R.wgan_gp_train_step(st, x, tile(g["s%d.z_d" % step]), tile(g["s%d.z_g" % step]), real,
                                   torch.from_numpy(g["s%d.alpha" % step]))
        for key in ("loss_d", "gp", "loss_g"):
            ref = float(g["s%d.%s" % (step, key)])
            assert abs(out[key] - ref) <= 2e-4 * max(1.0, abs(ref)), (step, key, out[key], ref)
    for k, v in st.g.items():
        if "running_" in k:
            assert_rel(v, g["end.bufG." + k], 1e-3, k)
