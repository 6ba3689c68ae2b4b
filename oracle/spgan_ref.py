"""oracle/spgan_ref.py -- TEST INFRASTRUCTURE ONLY (CPU, torch fp32, functional style).

A from-scratch restatement of the SP-GAN hot path as pure functions over a flat
``state`` dict (same keys and shapes as the reference ``state_dict``).  It is the
checker for the CUDA product and the "port" CPU baseline of bench.py; the product
never imports it.  Each function cites the reference lines it restates
(paths relative to liruihui/SP-GAN):

  pairwise distance / kNN / edge features   Generation/modules.py:683-725
  edgeConv                                  Generation/modules.py:612-626, 779-796
  EdgeBlock                                 Generation/Generator.py:47-88
  AdaptivePointNorm                         Generation/Generator.py:24-45
  Generator.forward / .interpolate          Generation/Generator.py:160-261
  Discriminator.forward                     Generation/Discriminator.py:97-114
  GradientPenalty                           Common/gradient_penalty.py:19-37
  wgan losses                               Common/loss_utils.py:728-730, 859-863
  train step                                Generation/model.py:239-279, 94-97

Pinned by tests/test_oracle_modules.py against tests/golden/*.npz, which were
produced by the unmodified reference modules (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

NEG_SLOPE = 0.01      # Generator.py:21, Discriminator.py:19
NEG_SLOPE_2 = 0.2     # Generator.py:22
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
IN_EPS = 1e-5


def default_opts(**kw):
    """The Generation/config.py fields the hot path reads (config.py:52-54,64,87,92,99,115,121,128)."""
    o = SimpleNamespace(np=2048, nk=20, nz=128, softmax=True, off=False, attn=False,
                        use_head=False, eql=False, z_norm=False, small_d=False)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# ----------------------------------------------------------------------------------------
# state-dict specs (keys / shapes of SURVEY 8b) and deterministic synthetic weights
# ----------------------------------------------------------------------------------------
def _bn(prefix, c):
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
            (prefix + ".running_mean", (c,), "bn_rm"), (prefix + ".running_var", (c,), "bn_rv"),
            (prefix + ".num_batches_tracked", (), "count")]


def _conv(prefix, shape):
    return [(prefix + ".weight", tuple(shape), "w"), (prefix + ".bias", (shape[0],), "b")]


def edge_block_spec(prefix, fin, fout, k):
    p = prefix + "." if prefix else ""
    s = []
    s += _conv(p + "conv_w.0", (fout // 2, fin, 1, 1)) + _bn(p + "conv_w.1", fout // 2)
    s += _conv(p + "conv_w.3", (fout, fout // 2, 1, 1)) + _bn(p + "conv_w.4", fout)
    s += _conv(p + "conv_x.0", (fout, 2 * fin, 1, 1)) + _bn(p + "conv_x.1", fout)
    s += _conv(p + "conv_out", (fout, fout, 1, k))
    return s


def edge_conv_spec(prefix, fin, fout):
    p = prefix + "." if prefix else ""
    return _conv(p + "conv.conv", (fout, 2 * fin, 1, 1)) + _bn(p + "conv.bn", fout)


def _eq_conv(prefix, shape, eql, inner):
    """nn.Conv1d / nn.Linear keys, or the EqualConv1d / EqualLinear ones (modules.py:202-243, 256-272):
    <prefix>.<inner>.weight_orig ~ N(0,1) (kind "w1") and <prefix>.<inner>.bias."""
    if not eql:
        return _conv(prefix, shape)
    return [(prefix + "." + inner + ".weight_orig", tuple(shape), "w1"), (prefix + "." + inner + ".bias", (shape[0],), "b")]


def generator_spec(opts):
    """Key order follows module registration order in Generator.__init__ (Generator.py:92-156)."""
    dim, k = 128, opts.nk // 2
    eql = bool(opts.eql)
    s = []
    s += _eq_conv("head.0", (dim, 3 + opts.nz, 1), eql, "conv") + _eq_conv("head.2", (dim, dim, 1), eql, "conv")
    if opts.attn:                                              # Attention(640), modules.py:534-546
        ch = dim + 512
        s += [("attn.gamma", (), "gain"), ("attn.theta.weight", (ch // 8, ch, 1), "w"),
              ("attn.phi.weight", (ch // 8, ch, 1), "w"), ("attn.g.weight", (ch // 2, ch, 1), "w"),
              ("attn.o.weight", (ch, ch // 2, 1), "w")]
    s += _eq_conv("global_conv.0", (dim, dim), eql, "linear")
    s += _bn("global_conv.1", dim)
    s += _eq_conv("global_conv.3", (512, dim), eql, "linear")
    s += _bn("global_conv.4", 512)
    s += _conv("tail.0", (256, 512 + dim, 1)) + _conv("tail.2", (64, 256, 1)) + _conv("tail.4", (3, 64, 1))
    if opts.use_head:
        s += _eq_conv("pc_head.0", (dim // 2, 3, 1), eql, "conv") + _eq_conv("pc_head.2", (dim, dim // 2, 1), eql, "conv")
        s += edge_block_spec("EdgeConv1", dim, dim, k)
        s += _conv("adain1.style", (2 * dim, dim, 1))
        s += edge_block_spec("EdgeConv2", dim, dim, k)
        s += _conv("adain2.style", (2 * dim, dim, 1))
    else:
        s += edge_block_spec("EdgeConv1", 3, 64, k)
        s += _conv("adain1.style", (2 * 64, dim, 1))
        s += edge_block_spec("EdgeConv2", 64, dim, k)
        s += _conv("adain2.style", (2 * dim, dim, 1))
    return s


def discriminator_spec(opts):
    """Discriminator.__init__ (Discriminator.py:48-95)."""
    dim = 512 if opts.small_d else 1024
    s = []
    s += _conv("mlps.0", (64, 3, 1)) + _bn("mlps.1", 64)
    s += _conv("mlps.3", (128, 64, 1)) + _bn("mlps.4", 128)
    s += _conv("mlps.6", (256, 128, 1)) + _bn("mlps.7", 256)
    s += _conv("fc2.0", (dim, 256, 1)) + _bn("fc2.1", dim)
    s += [("mlp.0.weight", (512, dim), "w"), ("mlp.0.bias", (512,), "b"),
          ("mlp.2.weight", (256, 512), "w"), ("mlp.2.bias", (256,), "b"),
          ("mlp.4.weight", (64, 256), "w"), ("mlp.4.bias", (64,), "b"),
          ("mlp.6.weight", (1, 64), "w"), ("mlp.6.bias", (1,), "b")]
    return s


def synth_state(spec, seed):
    """Deterministic, platform-independent weights (numpy PCG64) for parity fixtures.

    Not the reference initialisation -- parity tests always move weights by state_dict
    (SURVEY 8c), so any non-degenerate values do.  Scales keep activations O(1).
    """
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for key, shape, kind in spec:
        if kind == "w":
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            v = rng.standard_normal(shape) * (1.0 / math.sqrt(fan_in))
        elif kind == "w1":                      # equalised-lr weight_orig: unit normal, scaled at use
            v = rng.standard_normal(shape)
        elif kind == "gain":                    # Attention.gamma (0 at init; non-zero here to exercise the block)
            v = np.asarray(0.5)
        elif kind == "b":
            v = 0.1 * rng.standard_normal(shape)
        elif kind == "bn_w":
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif kind == "bn_b":
            v = 0.1 * rng.standard_normal(shape)
        elif kind == "bn_rm":
            v = 0.1 * rng.standard_normal(shape)
        elif kind == "bn_rv":
            v = 1.0 + 0.1 * np.abs(rng.standard_normal(shape))
        elif kind == "count":
            sd[key] = torch.zeros((), dtype=torch.long)
            continue
        else:
            raise KeyError(kind)
        sd[key] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape))
    return sd


def sub_state(sd, prefix):
    p = prefix + "."
    return OrderedDict((k[len(p):], v) for k, v in sd.items() if k.startswith(p))


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
def _batch_norm(x, sd, p, training):
    """nn.BatchNorm{1,2}d semantics incl. the running-stat side effects (momentum 0.1,
    unbiased running var, num_batches_tracked += 1) that checkpoints observe."""
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    if training:
        sd[p + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training, BN_MOMENTUM, BN_EPS)


def pairwise_sqdist(x):
    """[B,C,N] -> [B,N,N]; expanded form in the reference's op order (modules.py:695-699)."""
    xt = x.permute(0, 2, 1)
    inner = -2 * torch.bmm(xt, x)
    sq = torch.sum(xt ** 2, dim=2, keepdim=True)
    return inner + sq + sq.permute(0, 2, 1)


def knn_indices(x, k):
    """Ranks 1..k of the full ascending sort, flattened to [B, N*k] int64 (modules.py:702-704)."""
    B, _, N = x.shape
    order = torch.sort(pairwise_sqdist(x), dim=2)[1]
    return order[:, :, 1:k + 1].contiguous().view(B, N * k)


def edge_features(x, k, idx=None, return_idx=False):
    """[B,C,N] -> [B,2C,N,k]: first C channels the centre point, last C neighbour - centre
    (modules.py:706-720).  The per-cloud index_select loop is restated as one gather."""
    B, C, N = x.shape
    if idx is None:
        idx = knn_indices(x, k)
    nbr = torch.gather(x, 2, idx.view(B, 1, N * k).expand(B, C, N * k)).view(B, C, N, k)
    ctr = x.unsqueeze(3).expand(B, C, N, k)
    ee = torch.cat([ctr, nbr - ctr], dim=1)
    return (ee, idx) if return_idx else ee


def edge_conv(sd, x, k, training=True, idx=None):
    """edgeConv: edge features -> Conv2d 1x1 + BN2d + ReLU -> max over k (modules.py:779-796)."""
    ee = edge_features(x, k, idx=idx)
    y = F.conv2d(ee, sd["conv.conv.weight"], sd["conv.conv.bias"])
    y = torch.relu(_batch_norm(y, sd, "conv.bn", training))
    return y.max(dim=3)[0]


def edge_block(sd, x, k, training=True, idx=None):
    """EdgeBlock (Generator.py:75-88): attention weights from the difference half, softmax
    over neighbours, modulate conv_x(all 2C), contract neighbours with the [1,k] conv."""
    C = x.shape[1]
    ee = edge_features(x, k, idx=idx)
    w = F.conv2d(ee[:, C:], sd["conv_w.0.weight"], sd["conv_w.0.bias"])
    w = F.leaky_relu(_batch_norm(w, sd, "conv_w.1", training), NEG_SLOPE)
    w = F.conv2d(w, sd["conv_w.3.weight"], sd["conv_w.3.bias"])
    w = F.leaky_relu(_batch_norm(w, sd, "conv_w.4", training), NEG_SLOPE)
    w = torch.softmax(w, dim=-1)
    y = F.conv2d(ee, sd["conv_x.0.weight"], sd["conv_x.0.bias"])
    y = F.leaky_relu(_batch_norm(y, sd, "conv_x.1", training), NEG_SLOPE)
    out = F.conv2d(y * w, sd["conv_out.weight"], sd["conv_out.bias"])
    return out.squeeze(3)


def adaptive_point_norm(sd, x, style):
    """InstanceNorm1d(x) * gamma + beta with (gamma, beta) = split(Conv1d(style)) (Generator.py:38-45)."""
    s = F.conv1d(style, sd["style.weight"], sd["style.bias"])
    gamma, beta = s.chunk(2, dim=1)
    return gamma * F.instance_norm(x, eps=IN_EPS) + beta


def _wb(sd, prefix, opts, inner):
    """(weight, bias) of a Conv1d / Linear, or of its equalised-lr variant: weight_orig * sqrt(2 / fan_in)
    with fan_in = in_channels * kernel elements (modules.py:256-262)."""
    if not opts.eql:
        return sd[prefix + ".weight"], sd[prefix + ".bias"]
    w = sd[prefix + "." + inner + ".weight_orig"]
    fan_in = w.size(1) * w[0][0].numel()
    return w * math.sqrt(2.0 / fan_in), sd[prefix + "." + inner + ".bias"]


def attention(sd, x):
    """Attention.forward (modules.py:548-558) on x [B, ch, N]."""
    theta = F.conv1d(x, sd["theta.weight"])
    phi = F.conv1d(x, sd["phi.weight"])
    g = F.conv1d(x, sd["g.weight"])
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = F.conv1d(torch.bmm(g, beta.transpose(1, 2)), sd["o.weight"])
    return sd["gamma"] * o + x


def _head(sd, x, z, opts):
    if opts.z_norm:
        z = z / (z.norm(p=2, dim=-1, keepdim=True) + 1e-8)
    s = torch.cat([x, z], dim=-1).transpose(2, 1).contiguous()
    s = F.leaky_relu(F.conv1d(s, *_wb(sd, "head.0", opts, "conv")), NEG_SLOPE)
    return F.leaky_relu(F.conv1d(s, *_wb(sd, "head.2", opts, "conv")), NEG_SLOPE)


def _generator_body(sd, x, style, opts, training, idx1=None, idx2=None):
    """Everything after the style head (shared by forward and interpolate)."""
    B, N, _ = x.shape
    k = opts.nk // 2
    pc = x.transpose(2, 1).contiguous()
    if opts.use_head:
        pc = F.leaky_relu(F.conv1d(pc, *_wb(sd, "pc_head.0", opts, "conv")), 0.01)
        pc = F.leaky_relu(F.conv1d(pc, *_wb(sd, "pc_head.2", opts, "conv")), 0.01)
    x1 = edge_block(sub_state_view(sd, "EdgeConv1"), pc, k, training, idx1)
    x1 = adaptive_point_norm(sub_state_view(sd, "adain1"), F.leaky_relu(x1, NEG_SLOPE_2), style)
    x2 = edge_block(sub_state_view(sd, "EdgeConv2"), x1, k, training, idx2)
    x2 = adaptive_point_norm(sub_state_view(sd, "adain2"), F.leaky_relu(x2, NEG_SLOPE_2), style)
    g = x2.max(dim=2)[0]
    g = F.linear(g, *_wb(sd, "global_conv.0", opts, "linear"))
    g = F.leaky_relu(_batch_norm(g, sd, "global_conv.1", training), NEG_SLOPE)
    g = F.linear(g, *_wb(sd, "global_conv.3", opts, "linear"))
    g = F.leaky_relu(_batch_norm(g, sd, "global_conv.4", training), NEG_SLOPE)
    feat = torch.cat([g.unsqueeze(2).expand(B, g.shape[1], N), x2], dim=1)
    if opts.attn:
        feat = attention(sub_state_view(sd, "attn"), feat)
    t = F.leaky_relu(F.conv1d(feat, sd["tail.0.weight"], sd["tail.0.bias"]), NEG_SLOPE)
    t = F.leaky_relu(F.conv1d(t, sd["tail.2.weight"], sd["tail.2.bias"]), NEG_SLOPE)
    out = torch.tanh(F.conv1d(t, sd["tail.4.weight"], sd["tail.4.bias"]))
    return pc + out if opts.off else out, x1


class sub_state_view(dict):
    """Prefix view that shares storage with the parent dict (BN buffers update in place)."""

    def __init__(self, sd, prefix):
        super().__init__()
        self._sd, self._p = sd, prefix + "."

    def __getitem__(self, k):
        return self._sd[self._p + k]

    def __setitem__(self, k, v):
        self._sd[self._p + k] = v


def generator_forward(sd, x, z, opts, training=True, idx1=None, idx2=None, return_x1=False):
    """x [B,N,3], z [B,N,nz] -> [B,3,N] (Generator.py:160-198)."""
    style = _head(sd, x, z, opts)
    out, x1 = _generator_body(sd, x, style, opts, training, idx1, idx2)
    return (out, x1) if return_x1 else out


def generator_interpolate(sd, x, z1, z2, selection, alpha, opts, use_latent=False, training=False, idx2=None,
                         return_x1=False):
    """Generator.interpolate (Generator.py:200-261); like the reference it writes into z1."""
    sel = selection == 1
    if not use_latent:
        z1[:, sel] = z1[:, sel] * (1 - alpha) + z2[:, sel] * alpha
        style = _head(sd, x, z1, opts)
    else:
        s1, s2 = _head(sd, x, z1, opts), _head(sd, x, z2, opts)
        s1[:, :, sel] = s1[:, :, sel] * (1 - alpha) + s2[:, :, sel] * alpha
        style = s1
    out, x1 = _generator_body(sd, x, style, opts, training, None, idx2)
    return (out, x1) if return_x1 else out


def discriminator_forward(sd, x, training=True):
    """PointNet critic, x [B,3,N] any strides -> [B,1] (Discriminator.py:97-114)."""
    h = x
    for conv, bn in (("mlps.0", "mlps.1"), ("mlps.3", "mlps.4"), ("mlps.6", "mlps.7"), ("fc2.0", "fc2.1")):
        h = F.conv1d(h, sd[conv + ".weight"], sd[conv + ".bias"])
        h = F.leaky_relu(_batch_norm(h, sd, bn, training), NEG_SLOPE)
    h = h.max(dim=2)[0]
    for i, lin in enumerate(("mlp.0", "mlp.2", "mlp.4", "mlp.6")):
        h = F.linear(h, sd[lin + ".weight"], sd[lin + ".bias"])
        if i < 3:
            h = F.leaky_relu(h, NEG_SLOPE)
    return h


def gradient_penalty(d_fn, real, fake, alpha, lambda_gp=10.0, gamma=1.0):
    """WGAN-GP term (gradient_penalty.py:19-37) with the U[0,1) mixing factors injected:
    alpha [B,1,1].  d_fn maps [B,3,N] -> [B,1] and must be twice differentiable."""
    B = real.shape[0]
    fake = fake[:B]
    alpha = alpha.detach().clone().requires_grad_(True)
    mix = real + alpha * (fake - real)
    out = d_fn(mix)
    (g,) = torch.autograd.grad(out, mix, grad_outputs=torch.ones_like(out),
                               create_graph=True, retain_graph=True)
    g = g.contiguous().view(B, -1)
    return (((g.norm(2, dim=1) - gamma) / gamma) ** 2).mean() * lambda_gp


def dis_loss_wgan(d_real, d_fake):
    """loss_utils.py:859-863."""
    return d_fake.mean() - d_real.mean()


def gen_loss_wgan(d_fake):
    """loss_utils.py:728-730."""
    return -d_fake.mean()


# ----------------------------------------------------------------------------------------
# the composed WGAN-GP training step
# ----------------------------------------------------------------------------------------
class TrainState:
    """G and D parameters/buffers plus their Adam optimizers (model.py:94-97)."""

    def __init__(self, g_state, d_state, opts, lr_g=1e-4, lr_d=1e-4, betas=(0.5, 0.99)):
        self.opts = opts
        self.g = OrderedDict((k, v.detach().clone()) for k, v in g_state.items())
        self.d = OrderedDict((k, v.detach().clone()) for k, v in d_state.items())
        self.g_params = [k for k, v in self.g.items() if v.is_floating_point() and "running_" not in k]
        self.d_params = [k for k, v in self.d.items() if v.is_floating_point() and "running_" not in k]
        self.opt_g = torch.optim.Adam([self.g[k] for k in self.g_params], lr=lr_g, betas=betas)
        self.opt_d = torch.optim.Adam([self.d[k] for k in self.d_params], lr=lr_d, betas=betas)

    def set_requires_grad(self, g_flag, d_flag):
        """Common/network_utils.py:92-94 applied to both nets (model.py:240-241, 264-265)."""
        for k in self.g_params:
            self.g[k].requires_grad_(g_flag)
        for k in self.d_params:
            self.d[k].requires_grad_(d_flag)


def wgan_gp_train_step(st, x, z_d, z_g, real, alpha, lambda_gp=10.0, gamma=1.0, idx2_d=None, idx2_g=None, trace=None):
    """One iteration of model.py:239-279 with gan='wgan' and GradientPenalty(lambda_gp) added
    to lossD.  x [B,N,3] sphere, z_* [B,N,nz], real [B,3,N], alpha [B,1,1].
    idx2_d / idx2_g optionally pin EdgeConv2's neighbour lists of the two generator forwards
    (int64 [B, N*k]) for sensitivity studies.  Returns python floats {loss_d, gp, loss_g}.
    `trace` (a dict) receives, per phase p in (d, g): x1_p (EdgeConv2's input), idx2_p (the neighbour list the
    reference recipe derives from it), fake_p (the generator output) -- what a parity test at full size needs."""
    opts = st.opts

    def g_forward(z, idx2, tag):
        out, x1 = generator_forward(st.g, x, z, opts, training=True, idx2=idx2, return_x1=True)
        if trace is not None:
            with torch.no_grad():
                trace["x1_" + tag] = x1.detach()
                trace["idx2_" + tag] = idx2 if idx2 is not None else knn_indices(x1.detach(), opts.nk // 2)
                trace["fake_" + tag] = out.detach()
        return out

    # ---- D phase ----
    st.set_requires_grad(False, True)
    st.opt_d.zero_grad(set_to_none=True)
    real = real.detach().clone().requires_grad_(True)     # model.py:245 (Variable(..., requires_grad=True))
    fake = g_forward(z_d, idx2_d, "d").detach()
    d_real = discriminator_forward(st.d, real, True)
    d_fake = discriminator_forward(st.d, fake, True)
    gp = gradient_penalty(lambda t: discriminator_forward(st.d, t, True), real, fake, alpha, lambda_gp, gamma)
    loss_d = dis_loss_wgan(d_real, d_fake) + gp
    loss_d.backward()
    st.opt_d.step()
    # ---- G phase ----
    st.set_requires_grad(True, False)
    st.opt_g.zero_grad(set_to_none=True)
    fake = g_forward(z_g, idx2_g, "g")
    _ = discriminator_forward(st.d, real, True)            # model.py:274 (result unused by wgan gen_loss)
    d_fake = discriminator_forward(st.d, fake, True)
    loss_g = gen_loss_wgan(d_fake)
    loss_g.backward()
    st.opt_g.step()
    return {"loss_d": float(loss_d.detach()), "gp": float(gp.detach()), "loss_g": float(loss_g.detach())}


# ----------------------------------------------------------------------------------------
# synthetic inputs: sphere, latent, "chair" clouds
# ----------------------------------------------------------------------------------------
def normalize_cloud(pc):
    """Centre and scale to unit max radius (model.py:46-52; Common/point_operation.py:21-40)."""
    pc = pc - pc.mean(axis=-2, keepdims=True)
    r = np.sqrt((pc ** 2).sum(axis=-1)).max(axis=-1)
    return pc / r[..., None, None] if pc.ndim == 3 else pc / r


def fibonacci_sphere(n):
    """Fallback unit-sphere template when the 2048-point fixture is unavailable."""
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = math.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)


def latent_noise(rng, B, N, nz, nv=0.2):
    """One N(0, nv) vector per cloud tiled over the N points (model.py:128-131)."""
    z = rng.normal(0.0, nv, (B, 1, nz))
    return np.tile(z, (1, N, 1)).astype(np.float32)


def synthetic_chairs(rng, B, N):
    """Stand-in for the ShapeNet chair H5 data (no network): N points sampled uniformly by
    area from a union of boxes (seat, back, four legs) with per-cloud random proportions,
    then normalised and shuffled like H5DataLoader (H5DataLoader.py:111-113).  [B,N,3] fp32."""
    out = np.empty((B, N, 3), np.float64)
    for b in range(B):
        w, d = rng.uniform(0.35, 0.5), rng.uniform(0.35, 0.5)
        seat_h, leg_t = rng.uniform(0.35, 0.5), rng.uniform(0.03, 0.06)
        back_h, slab = rng.uniform(0.4, 0.7), rng.uniform(0.04, 0.08)
        boxes = [((-w, seat_h, -d), (w, seat_h + slab, d)),
                 ((-w, seat_h + slab, -d), (w, seat_h + slab + back_h, -d + slab))]
        for sx in (-1, 1):
            for sz in (-1, 1):
                cx, cz = sx * (w - leg_t), sz * (d - leg_t)
                boxes.append(((cx - leg_t, 0.0, cz - leg_t), (cx + leg_t, seat_h, cz + leg_t)))
        faces, areas = [], []
        for lo, hi in boxes:
            lo, hi = np.array(lo), np.array(hi)
            e = hi - lo
            for ax in range(3):
                u, v = (ax + 1) % 3, (ax + 2) % 3
                for side in (lo[ax], hi[ax]):
                    faces.append((ax, side, lo, e, u, v))
                    areas.append(e[u] * e[v])
        areas = np.array(areas)
        which = rng.choice(len(faces), size=N, p=areas / areas.sum())
        uv = rng.uniform(size=(N, 2))
        pts = np.empty((N, 3))
        for i, f in enumerate(which):
            ax, side, lo, e, u, v = faces[f]
            p = np.empty(3)
            p[ax] = side
            p[u] = lo[u] + uv[i, 0] * e[u]
            p[v] = lo[v] + uv[i, 1] * e[v]
            pts[i] = p
        pts = normalize_cloud(pts)
        out[b] = pts[rng.permutation(N)]
    return out.astype(np.float32)
