"""kNN graph + EdgeConv blocks: get_edge_features, edgeConv, EdgeBlock.

Same constructor / forward signatures and parameter names as the reference
(Generation/modules.py:683-725, 612-626, 779-796 and Generation/Generator.py:47-88) so that
state_dicts move freely; the nn.Conv2d / nn.BatchNorm2d children are parameter holders only
-- forward never calls them, it runs the libspgan_b200 kernels through `ops`.
"""
import torch
import torch.nn as nn

from . import ops

NEG = 0.01           # Generator.py:21


def get_edge_features(x, k, num=-1, idx=None, return_idx=False):
    """x [B, dims, N] -> ee [B, 2*dims, N, k] (first half centre, second half neighbour - centre),
    optionally the int64 neighbour list [B, N*k].  Mirrors modules.py:683-725, including the
    `idx=` argument (a precomputed flattened neighbour list) and the unused `num`."""
    B, dims, N = x.shape
    if idx is None:
        if x.requires_grad and torch.is_grad_enabled():
            idx32 = ops.knn_indices(x, k)
            ee = ops.Group.apply(x, idx32, k)
        else:
            idx32, ee = ops.knn_indices(x, k, want_ee=True)       # fused kNN + group kernel
        if return_idx:
            return ee, ops.idx_to_int64(idx32).view(B, N * k)
        return ee
    idx32 = ops.idx_to_int32(idx.reshape(B, N, k))
    ee = ops.Group.apply(x, idx32, k)
    if return_idx:
        return ee, idx
    return ee


class _KnnMixin:
    """Neighbour-list plumbing shared by the blocks: accepts the reference's channel-first
    input, or point-major rows plus a precomputed list from the generator's fast path."""

    def _graph(self, x_bcn, idx):
        B, C, N = x_bcn.shape
        if idx is None:
            return ops.knn_indices(x_bcn, self.k)
        if idx.dtype == torch.int64:
            return ops.idx_to_int32(idx.reshape(B, N, self.k))
        return idx.reshape(B, N, self.k)


class conv2dbr(nn.Module):
    """Parameter holder with the reference's names (modules.py:612-626): conv, bn."""

    def __init__(self, Fin, Fout, kernel_size, stride=[1, 1]):
        super().__init__()
        self.conv = nn.Conv2d(Fin, Fout, kernel_size, stride)
        self.bn = nn.BatchNorm2d(Fout)
        self.ac = nn.ReLU(True)


class edgeConv(nn.Module, _KnnMixin):
    """[B, Fin, N] -> [B, Fout, N]: edge features -> 1x1 conv + BN + ReLU -> max over k
    (modules.py:779-796)."""

    def __init__(self, Fin, Fout, k):
        super().__init__()
        self.k, self.Fin, self.Fout = k, Fin, Fout
        self.conv = conv2dbr(2 * Fin, Fout, 1)

    def forward_rows(self, x_rows, idx32, B, N):
        C, F = self.Fin, self.Fout
        W = self.conv.conv.weight
        a = ops.linear(x_rows, W, cols=(0, C))                       # centre half, per point
        d = ops.linear(x_rows, W, engine=0, cols=(C, 2 * C))         # difference half, per point (differenced: exact fp32)
        y = ops.EdgeCombine.apply(a, d, self.conv.conv.bias, idx32, N, self.k, ops.feeds_train_bn(self.conv.bn))
        y = ops.batch_norm_act(y, self.conv.bn, 0.0)
        return ops.KMax.apply(y, self.k)

    def forward(self, x, idx=None):
        B, Fin, N = x.shape
        idx32 = self._graph(x, idx)
        out = self.forward_rows(ops.BcnToRows.apply(x), idx32, B, N)
        return ops.RowsToBcn.apply(out, B, self.Fout, N)


class EdgeBlock(nn.Module, _KnnMixin):
    """[B, Fin, N] -> [B, Fout, N] (Generator.py:47-88): softmax-over-neighbours weights from the
    difference half modulate conv_x of the full edge feature; a [1,k] conv contracts neighbours.
    `attn` is accepted and ignored, as in the reference."""

    def __init__(self, Fin, Fout, k, attn=True):
        super().__init__()
        self.k, self.Fin, self.Fout = k, Fin, Fout
        self.conv_w = nn.Sequential(
            nn.Conv2d(Fin, Fout // 2, 1), nn.BatchNorm2d(Fout // 2), nn.LeakyReLU(NEG, inplace=True),
            nn.Conv2d(Fout // 2, Fout, 1), nn.BatchNorm2d(Fout), nn.LeakyReLU(NEG, inplace=True))
        self.conv_x = nn.Sequential(
            nn.Conv2d(2 * Fin, Fout, [1, 1], [1, 1]), nn.BatchNorm2d(Fout), nn.LeakyReLU(NEG, inplace=True))
        self.conv_out = nn.Conv2d(Fout, Fout, [1, k], [1, 1])

    def forward_rows(self, x_rows, idx32, B, N):
        """x_rows [B*N, Fin] point-major, idx32 [B, N, k] -> [B*N, Fout]."""
        C, F, k = self.Fin, self.Fout, self.k
        P = x_rows.shape[0]
        cw0, bw0, _, cw1, bw1, _ = self.conv_w
        cx, bx, _ = self.conv_x
        # conv_w on the difference half: W (x_j - x_i) + b == (W x)_j - (W x)_i + b
        p1 = ops.linear(x_rows, cw0.weight, engine=0)          # differenced below: exact fp32 products
        fuse_w = ops.edge_attention_fusable(bw1, bx, k) and ops.fused_linear_ok(P * k, cw1.weight, bw0, bw1)
        stats_w = None
        if fuse_w:
            # the gather kernel leaves conv_w[1]'s batch statistics; BN + LeakyReLU of conv_w[0..2] run inside conv_w[3]'s
            # operand converter; the statistics of its output (the attention logits before their BatchNorm) come out
            # of the same GEMM's epilogue
            if ops.edge_stats_fusable(P, F // 2, bw0):
                # gather + conv_w[1..3] as one node: see ops.EdgeGatherBnActLinearTrain
                w, st, var_w = ops.edge_gather_bn_act_linear(p1, cw0.bias, idx32, N, k, bw0, cw1.weight, cw1.bias, NEG, bw1,
                                                             ops.feeds_train_bn(bw1))
            else:
                w = ops.EdgeCombine.apply(None, p1, cw0.bias, idx32, N, k, ops.feeds_train_bn(bw0))
                st0 = ops.bn_train_stats(w, bw0)
                w, st, var_w = ops.bn_act_linear(w, st0, bw0, NEG, cw1.weight, cw1.bias, next_bn=bw1,
                                                 zero_bias_grad=ops.feeds_train_bn(bw1))
            stats_w = (st[0], st[1], var_w)
        else:
            w = ops.EdgeCombine.apply(None, p1, cw0.bias, idx32, N, k, ops.feeds_train_bn(bw0))   # [P*k, F/2], pre-BN
            w = ops.batch_norm_act(w, bw0, NEG)
            w = ops.linear(w, cw1.weight, cw1.bias, zero_bias_grad=ops.feeds_train_bn(bw1))   # [P*k, F], pre-BN
        # conv_x on [centre, difference]
        a = ops.linear(x_rows, cx.weight, cols=(0, C))
        d = ops.linear(x_rows, cx.weight, engine=0, cols=(C, 2 * C))
        if ops.edge_attention_stats_fusable(P, F, k, bw1, bx):
            # gather + its BatchNorm statistics, then BN + LeakyReLU on both branches, softmax over k and y * w in one pass
            y = ops.edge_attention(w, stats_w, bw1, a, d, cx.bias, idx32, N, k, bx, NEG)
        else:
            y = ops.EdgeCombine.apply(a, d, cx.bias, idx32, N, k, ops.feeds_train_bn(bx))   # [P*k, F], pre-BN
            y = ops.bn_act_softmax_mul_k(w, bw1, y, bx, NEG, k, stats_w)
        # conv_out: kernel [1, k] == one dense contraction over (neighbour, channel)
        Wo = ops.PermuteOCK.apply(self.conv_out.weight)                         # [F, k*F]
        return ops.Gemm.apply(y.view(P, k * F), Wo, self.conv_out.bias, False, True)

    def forward(self, x, idx=None):
        B, C, N = x.shape
        idx32 = self._graph(x, idx)
        out = self.forward_rows(ops.BcnToRows.apply(x), idx32, B, N)
        return ops.RowsToBcn.apply(out, B, self.Fout, N)
