"""AdaptivePointNorm and Generator with the reference API (Generation/Generator.py:24-45, 91-261)."""
import math

import torch
import torch.nn as nn

from . import ops
from .edge import EdgeBlock, NEG

NEG_2 = 0.2          # Generator.py:22
IN_EPS = 1e-5


class _EqualLR(nn.Module):
    """Equalised-learning-rate holder (modules.py:202-288): the wrapped layer keeps `weight_orig` (and `bias`)
    as parameters -- same state_dict keys as the reference's EqualConv1d / EqualLinear -- and the weight used
    is weight_orig * sqrt(2 / fan_in), rebuilt (differentiably) at every use like the reference's
    forward-pre-hook does."""

    def _wrap(self, layer):
        layer.weight.data.normal_()
        layer.bias.data.zero_()
        w = layer.weight
        del layer._parameters["weight"]
        layer.register_parameter("weight_orig", nn.Parameter(w.data))
        fan_in = w.data.size(1) * w.data[0][0].numel()            # modules.py:259
        self._scale = math.sqrt(2.0 / fan_in)
        return layer

    def _layer(self):
        raise NotImplementedError

    @property
    def weight(self):
        return ops.scale(self._layer().weight_orig, self._scale)

    @property
    def bias(self):
        return self._layer().bias


class EqualConv1d(_EqualLR):
    """modules.py:202-213 (parameters: conv.weight_orig, conv.bias)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.conv = self._wrap(nn.Conv1d(*args, **kwargs))

    def _layer(self):
        return self.conv


class EqualLinear(_EqualLR):
    """modules.py:230-243 (parameters: linear.weight_orig, linear.bias)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear = self._wrap(nn.Linear(in_dim, out_dim))

    def _layer(self):
        return self.linear


class Attention(nn.Module):
    """Optional N x N self-attention over the 640-channel feature (--attn; modules.py:534-558): parameter
    holder for theta / phi / g / o (bias-free Conv1d) and the learnable gain gamma; forward_rows composes the
    same primitives as the rest of the generator (no tuned path: non-default flag)."""

    def __init__(self, ch, name="attention"):
        super().__init__()
        self.ch = ch
        self.theta = nn.Conv1d(ch, ch // 8, 1, bias=False)
        self.phi = nn.Conv1d(ch, ch // 8, 1, bias=False)
        self.g = nn.Conv1d(ch, ch // 2, 1, bias=False)
        self.o = nn.Conv1d(ch // 2, ch, 1, bias=False)
        self.gamma = nn.Parameter(torch.tensor(0.), requires_grad=True)

    def forward_rows(self, x_rows, B, N):
        theta = ops.linear(x_rows, self.theta.weight)                        # [B*N, ch/8]
        phi = ops.linear(x_rows, self.phi.weight)
        g = ops.linear(x_rows, self.g.weight)                                # [B*N, ch/2]
        beta = ops.RowSoftmax.apply(ops.SegGemm.apply(theta, phi, B, True))  # softmax_j(theta_i . phi_j), [B*N, N]
        o = ops.linear(ops.SegGemm.apply(beta, g, B, False), self.o.weight)  # sum_j beta_ij g_j -> [B*N, ch]
        n = o.numel()
        gain = ops.BcastSeg.apply(self.gamma.view(1, 1), n, n)               # the scalar gain, differentiable
        return ops.add(ops.Mul.apply(o.view(n, 1), gain).view(o.shape), x_rows)

    def forward(self, x, y=None):
        B, C, N = x.shape
        return ops.RowsToBcn.apply(self.forward_rows(ops.BcnToRows.apply(x), B, N), B, C, N)


class AdaptivePointNorm(nn.Module):
    """InstanceNorm1d(input) * gamma + beta with (gamma, beta) = chunk(Conv1d(style)) (Generator.py:24-45)."""

    def __init__(self, in_channel, style_dim, use_eql=False):
        super().__init__()
        if use_eql:
            # the reference's own AdaptivePointNorm(use_eql=True) raises AttributeError (Generator.py:32 touches
            # EqualConv1d.weight before any forward); Generator never passes it (Generator.py:146-153)
            raise NotImplementedError("AdaptivePointNorm(use_eql=True) is unusable in the reference as well")
        self.in_channel = in_channel
        self.norm = nn.InstanceNorm1d(in_channel)
        self.style = nn.Conv1d(style_dim, in_channel * 2, 1)
        self.style.weight.data.normal_()
        self.style.bias.data.zero_()
        self.style.bias.data[:in_channel] = 1
        self.style.bias.data[in_channel:] = 0

    def forward_rows(self, x_rows, style_rows, N, style_slope=None):
        """style_slope: the style rows are PRE-activation values of the generator head and LeakyReLU(style_slope) is
        still to be applied -- it is, inside the style conv's GEMM (ops.act_linear), so the activated style tensor is
        never written."""
        if style_slope is None:
            s = ops.linear(style_rows, self.style.weight, self.style.bias)          # [P, 2C]
        else:
            s = ops.act_linear(style_rows, style_slope, self.style.weight, self.style.bias)
        return ops.AdaIN.apply(x_rows, s, N, self.norm.eps)

    def forward(self, input, style):
        B, C, N = input.shape
        out = self.forward_rows(ops.BcnToRows.apply(input), ops.BcnToRows.apply(style), N)
        return ops.RowsToBcn.apply(out, B, C, N)


class Generator(nn.Module):
    """Sphere-guided generator: forward(x [B,N,3], z [B,N,nz]) -> [B,3,N] (Generator.py:160-198).

    Reads the same opts fields as the reference (np, nk, nz, softmax, off, attn, use_head, eql,
    z_norm).  Activations stay point-major between the module boundaries; the global feature's
    512 input channels of tail[0] are folded into a per-cloud bias (results-preserving)."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.np = opts.np
        self.nk = opts.nk // 2
        self.nz = opts.nz
        self.off = opts.off
        self.use_attn = opts.attn
        self.use_head = opts.use_head
        eql = bool(getattr(opts, "eql", False))
        Conv = EqualConv1d if eql else nn.Conv1d                  # Generator.py:103-104
        Linear = EqualLinear if eql else nn.Linear
        dim = 128
        self.head = nn.Sequential(
            Conv(3 + self.nz, dim, 1), nn.LeakyReLU(NEG, inplace=True),
            Conv(dim, dim, 1), nn.LeakyReLU(NEG, inplace=True))
        if self.use_attn:
            self.attn = Attention(dim + 512)                      # Generator.py:116-117
        self.global_conv = nn.Sequential(
            Linear(dim, dim), nn.BatchNorm1d(dim), nn.LeakyReLU(NEG, inplace=True),
            Linear(dim, 512), nn.BatchNorm1d(512), nn.LeakyReLU(NEG, inplace=True))
        self.tail = nn.Sequential(
            nn.Conv1d(512 + dim, 256, 1), nn.LeakyReLU(NEG, inplace=True),
            nn.Conv1d(256, 64, 1), nn.LeakyReLU(NEG, inplace=True),
            nn.Conv1d(64, 3, 1), nn.Tanh())
        if self.use_head:
            self.pc_head = nn.Sequential(
                Conv(3, dim // 2, 1), nn.LeakyReLU(inplace=True),
                Conv(dim // 2, dim, 1), nn.LeakyReLU(inplace=True))
            self.EdgeConv1 = EdgeBlock(dim, dim, self.nk)
            self.adain1 = AdaptivePointNorm(dim, dim)
            self.EdgeConv2 = EdgeBlock(dim, dim, self.nk)
            self.adain2 = AdaptivePointNorm(dim, dim)
        else:
            self.EdgeConv1 = EdgeBlock(3, 64, self.nk)
            self.adain1 = AdaptivePointNorm(64, dim)
            self.EdgeConv2 = EdgeBlock(64, dim, self.nk)
            self.adain2 = AdaptivePointNorm(dim, dim)
        self.lrelu1 = nn.LeakyReLU(NEG_2)
        self.lrelu2 = nn.LeakyReLU(NEG_2)
        self._graph_cache = None          # neighbour list of the static sphere (model.py:231)
        self.cache_sphere_graph = True
        self.debug_idx = None             # (idx1, idx2) int32 overrides for parity tests

    # ------------------------------------------------------------------ pieces
    def _style(self, x_rows, z, B, N, activated=True):
        """head(cat([x, z])) (Generator.py:163-168) -> [B*N, 128].  activated=False returns the values BEFORE the head's
        last LeakyReLU: the consumers (the two AdaIN style convs) apply it inside their GEMMs."""
        nz = z.shape[-1]
        if self.opts.z_norm:
            z = ops.row_l2_normalize(z if z.is_contiguous() else z.contiguous(), 1e-8)
        if z.dim() == 3 and z.shape[1] == N and z.stride(1) == 0:
            zz, bcast = z[:, 0, :], True                 # expanded per-cloud latent: broadcast in-kernel
        else:
            zz, bcast = z.reshape(B * N, nz), False
        s = ops.ConcatCols.apply(x_rows, zz, N, bcast)
        s = ops.linear(s, self.head[0].weight, self.head[0].bias)
        s = ops.act_linear(s, NEG, self.head[2].weight, self.head[2].bias)      # LeakyReLU(head[0]) inside head[2]'s GEMM
        return ops.LRelu.apply(s, NEG) if activated else s

    def _sphere_graph(self, x, pc_rows, B, N):
        """Neighbour list of EdgeConv1's input.  Without pc_head that input is the sphere itself,
        which train.py generates once and reuses for every step (model.py:231), so the list is
        cached on (storage, version, shape) of `x` (with `x` itself kept alive) and recomputed whenever `x` changes."""
        if self.debug_idx is not None and self.debug_idx[0] is not None:
            return self.debug_idx[0]
        knn = lambda: ops.knn_indices_rows(pc_rows, B, N, self.nk)
        if not self.cache_sphere_graph or self.use_head:
            return knn()
        # The entry keeps a strong reference to `x`: while it lives, its storage cannot be freed and handed to a
        # different point set at the same address, so (address, version counter, geometry) identifies the CONTENT.
        # (Writes that bypass the version counter -- x.data.copy_() -- are the caller's to announce: set
        # cache_sphere_graph = False or call invalidate_sphere_graph().)
        key = (x.data_ptr(), x._version, tuple(x.shape), tuple(x.stride()), self.nk, str(x.device))
        if self._graph_cache is None or self._graph_cache[0] != key:
            self._graph_cache = (key, knn(), x)
        return self._graph_cache[1]

    def invalidate_sphere_graph(self):
        self._graph_cache = None

    def _body(self, x, x_rows, style, B, N, style_slope=None):
        pc_rows = x_rows
        if self.use_head:
            slope = self.pc_head[1].negative_slope
            pc_rows = ops.LRelu.apply(ops.linear(pc_rows, self.pc_head[0].weight, self.pc_head[0].bias), slope)
            pc_rows = ops.LRelu.apply(ops.linear(pc_rows, self.pc_head[2].weight, self.pc_head[2].bias), slope)
        idx1 = self._sphere_graph(x, pc_rows, B, N)

        x1 = self.EdgeConv1.forward_rows(pc_rows, idx1, B, N)
        x1 = self.adain1.forward_rows(ops.LRelu.apply(x1, NEG_2), style, N, style_slope)

        if self.debug_idx is not None and self.debug_idx[1] is not None:
            idx2 = self.debug_idx[1]
        else:
            idx2 = ops.knn_indices_rows(x1, B, N, self.nk)       # point-major: no [B,C,N] transpose of the features
        x2 = self.EdgeConv2.forward_rows(x1, idx2, B, N)
        x2 = self.adain2.forward_rows(ops.LRelu.apply(x2, NEG_2), style, N, style_slope)
        self._last_x1 = x1.detach()

        g = ops.SegMax.apply(x2, N)                                              # [B, 128]
        gc = self.global_conv
        g = ops.batch_norm_act(ops.linear(g, gc[0].weight, gc[0].bias, zero_bias_grad=ops.feeds_train_bn(gc[1])), gc[1], NEG)
        g = ops.batch_norm_act(ops.linear(g, gc[3].weight, gc[3].bias, zero_bias_grad=ops.feeds_train_bn(gc[4])), gc[4], NEG)   # [B, 512]

        W0 = self.tail[0].weight.view(self.tail[0].weight.shape[0], -1)
        ng = g.shape[1]
        if self.use_attn:
            # the attention block mixes all 640 channels of cat(global, x2): materialise it (Generator.py:189-192)
            feat = ops.ConcatCols.apply(ops.BcastSeg.apply(g, B * N, N), x2, N, False)
            feat = self.attn.forward_rows(feat, B, N)
            t = ops.linear(feat, W0, self.tail[0].bias)
        else:
            # tail[0] over cat(global, x2): the global half is constant per cloud -> per-cloud bias
            gb = ops.linear(g, self.tail[0].weight, self.tail[0].bias, cols=(0, ng))           # [B, 256]
            t = ops.AddSegVec.apply(ops.linear(x2, self.tail[0].weight, cols=(ng, W0.shape[1])), gb, N)
        t = ops.LRelu.apply(ops.act_linear(t, NEG, self.tail[2].weight, self.tail[2].bias), NEG)   # LeakyReLU(tail[0]) in tail[2]'s GEMM
        o = ops.Tanh.apply(ops.linear(t, self.tail[4].weight, self.tail[4].bias))
        if self.off:
            o = ops.add(pc_rows, o)
        return ops.RowsToBcn.apply(o, B, 3, N)

    @staticmethod
    def _rows(x):
        B, N, C = x.shape
        x = x if x.is_contiguous() else ops.contiguous(x)
        return x.view(B * N, C)

    # ------------------------------------------------------------------ reference API
    def forward(self, x, z):
        B, N, _ = x.size()
        x_rows = self._rows(x)
        style = self._style(x_rows, z, B, N, activated=False)
        return self._body(x, x_rows, style, B, N, style_slope=NEG)

    def interpolate(self, x, z1, z2, selection, alpha, use_latent=False):
        """Latent / style blending on the points where selection == 1 (Generator.py:200-261).
        Like the reference, the non-latent branch writes the blend into z1 in place.  The blend
        itself is index bookkeeping on the caller's tensors and is done with torch indexing."""
        B, N, _ = x.size()
        x_rows = self._rows(x)
        sel = selection == 1
        if not use_latent:
            z = z1
            z[:, sel] = z1[:, sel] * (1 - alpha) + z2[:, sel] * alpha
            style = self._style(x_rows, z, B, N)
        else:
            s1 = self._style(x_rows, z1, B, N).view(B, N, -1)
            s2 = self._style(x_rows, z2, B, N).view(B, N, -1)
            s1[:, sel] = s1[:, sel] * (1 - alpha) + s2[:, sel] * alpha
            style = s1.view(B * N, -1)
        return self._body(x, x_rows, style, B, N)
