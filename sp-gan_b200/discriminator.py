"""PointNet critic with the reference API (Generation/Discriminator.py:48-114)."""
import torch.nn as nn

from . import ops

NEG = 0.01           # Discriminator.py:19


class Discriminator(nn.Module):
    """forward(x [B,3,N], any strides) -> [B,1].  Children are parameter holders with the
    reference's names (mlps.{0,1,3,4,6,7}, fc2.{0,1}, mlp.{0,2,4,6}); reads opts.small_d."""

    def __init__(self, opts, num_point=2048):
        super().__init__()
        self.num_point = num_point
        self.small_d = opts.small_d
        self.mlps = nn.Sequential(
            nn.Conv1d(3, 64, 1), nn.BatchNorm1d(64), nn.LeakyReLU(NEG, inplace=True),
            nn.Conv1d(64, 128, 1), nn.BatchNorm1d(128), nn.LeakyReLU(NEG, inplace=True),
            nn.Conv1d(128, 256, 1), nn.BatchNorm1d(256), nn.LeakyReLU(NEG, inplace=True))
        self.mode = "max"
        dim = 1024
        if self.small_d:
            dim = dim // 2
        self.fc2 = nn.Sequential(nn.Conv1d(256, dim, 1), nn.BatchNorm1d(dim), nn.LeakyReLU(NEG, inplace=True))
        self.mlp = nn.Sequential(
            nn.Linear(dim, 512), nn.LeakyReLU(NEG, inplace=True),
            nn.Linear(512, 256), nn.LeakyReLU(NEG, inplace=True),
            nn.Linear(256, 64), nn.LeakyReLU(NEG, inplace=True),
            nn.Linear(64, 1))

    def forward(self, x):
        B, _, N = x.shape
        h = ops.BcnToRows.apply(x)                                               # [B*N, 3]
        layers = ((self.mlps[0], self.mlps[1]), (self.mlps[3], self.mlps[4]), (self.mlps[6], self.mlps[7]),
                  (self.fc2[0], self.fc2[1]))
        conv0, bn0 = layers[0]
        y = ops.linear(h, conv0.weight, conv0.bias, zero_bias_grad=ops.feeds_train_bn(bn0))      # K = 3: CUDA cores
        if all(ops.fused_linear_ok(y.shape[0], layers[i + 1][0].weight, layers[i][1], layers[i + 1][1] if i < 2 else None)
               for i in range(3)):
            # conv -> BN -> LeakyReLU -> conv chains as ONE pass per layer: BN + LeakyReLU live in the next GEMM's
            # operand converter, the batch statistics come out of the producing GEMM's epilogue
            stats = ops.bn_train_stats(y, bn0)
            for i in range(3):
                conv, bn_in, nxt = layers[i + 1][0], layers[i][1], layers[i + 1][1]
                if i < 2:
                    y, stats, _ = ops.bn_act_linear(y, stats, bn_in, NEG, conv.weight, conv.bias, next_bn=nxt,
                                                    zero_bias_grad=ops.feeds_train_bn(nxt))
                else:
                    y = ops.bn_act_linear(y, stats, bn_in, NEG, conv.weight, conv.bias,
                                          zero_bias_grad=ops.feeds_train_bn(nxt))
            h = y
        else:
            h = ops.batch_norm_act(y, bn0, NEG)
            for conv, bn in layers[1:3]:
                h = ops.batch_norm_act(ops.linear(h, conv.weight, conv.bias, zero_bias_grad=ops.feeds_train_bn(bn)), bn, NEG)
            h = ops.linear(h, self.fc2[0].weight, self.fc2[0].bias, zero_bias_grad=ops.feeds_train_bn(self.fc2[1]))
        # fc2 -> BN -> LeakyReLU -> max over points: fused, the [B*N, dim] normalised tensor is never written
        h = ops.batch_norm_act_segmax(h, self.fc2[1], NEG, N)                    # [B, dim]
        for i in (0, 2, 4):
            h = ops.LRelu.apply(ops.linear(h, self.mlp[i].weight, self.mlp[i].bias), NEG)
        return ops.linear(h, self.mlp[6].weight, self.mlp[6].bias)
