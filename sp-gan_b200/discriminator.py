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
        for conv, bn in ((self.mlps[0], self.mlps[1]), (self.mlps[3], self.mlps[4]), (self.mlps[6], self.mlps[7])):
            h = ops.batch_norm_act(ops.linear(h, conv.weight, conv.bias, zero_bias_grad=ops.feeds_train_bn(bn)), bn, NEG)
        # fc2 -> BN -> LeakyReLU -> max over points: fused, the [B*N, dim] normalised tensor is never written
        h = ops.linear(h, self.fc2[0].weight, self.fc2[0].bias, zero_bias_grad=ops.feeds_train_bn(self.fc2[1]))
        h = ops.batch_norm_act_segmax(h, self.fc2[1], NEG, N)                    # [B, dim]
        for i in (0, 2, 4):
            h = ops.LRelu.apply(ops.linear(h, self.mlp[i].weight, self.mlp[i].bias), NEG)
        return ops.linear(h, self.mlp[6].weight, self.mlp[6].bias)
