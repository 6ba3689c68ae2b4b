"""The composed WGAN-GP training step: the loop body of
Generation/model.py:239-279 with gan='wgan' (Common/loss_utils.py:728-730, 859-863) and
Common/gradient_penalty.py::GradientPenalty(lambdaGP) added to lossD, Adam(1e-4, (0.5, 0.99))
as in model.py:94-97.

Data parallelism: one process per GPU, the batch is sharded across ranks, BN
statistics stay per replica (the reference's nn.DataParallel semantics), and each phase ends
with ONE NCCL all-reduce of a flat fp32 gradient buffer followed by a fused Adam kernel over
the flat parameter buffer.
"""
import torch

from . import ops
from .gradient_penalty import GradientPenalty
from .parallel import FlatBuffers


def requires_grad(model, flag=True):
    """Common/network_utils.py:92-94."""
    for p in model.parameters():
        p.requires_grad = flag


def dis_loss_wgan(d_real, d_fake):
    """loss_utils.py:859-863: mean(d_fake) - mean(d_real)."""
    return ops.add(ops.MeanScale.apply(d_fake, 1.0), ops.MeanScale.apply(d_real, -1.0))


def gen_loss_wgan(d_fake):
    """loss_utils.py:728-730: -mean(d_fake)."""
    return ops.MeanScale.apply(d_fake, -1.0)


class FlatAdam:
    """torch.optim.Adam semantics (model.py:94-97) as ONE fused kernel over the flat parameter buffer,
    preceded by ONE all-reduce of the flat gradient buffer when running data parallel."""

    def __init__(self, module, lr=1e-4, betas=(0.5, 0.99), eps=1e-8):
        self.buf = FlatBuffers(module)
        n, dev = self.buf.numel, self.buf.flat_p.device
        self.m = ops.full((n,), 0.0, dev)
        self.v = ops.full((n,), 0.0, dev)
        self.lr, self.betas, self.eps, self.t, self.n = lr, betas, eps, 0, n

    def zero_grad(self):
        ops.fill_(self.buf.flat_g, 0.0)
        self.buf.rebind_grads()

    def step(self):
        scale = self.buf.allreduce_grads()
        self.t += 1
        ops.L().adam_step(self.buf.flat_p.data_ptr(), self.buf.flat_g.data_ptr(), self.m.data_ptr(),
                          self.v.data_ptr(), self.n, self.lr, self.betas[0], self.betas[1], self.eps, self.t,
                          scale, ops._stream())


class WGANGPTrainer:
    """One object per rank.  step(x, z_d, z_g, real) runs the D phase then the G phase."""

    def __init__(self, G, D, lambda_gp=10.0, gamma=1.0, lr_g=1e-4, lr_d=1e-4, betas=(0.5, 0.99)):
        self.G, self.D = G, D
        self.gp = GradientPenalty(lambda_gp, gamma=gamma)
        self.opt_g = FlatAdam(G, lr_g, betas)
        self.opt_d = FlatAdam(D, lr_d, betas)
        self.opt_g.buf.broadcast_params()
        self.opt_d.buf.broadcast_params()

    def d_phase(self, x, z, real, alpha=None):
        G, D = self.G, self.D
        requires_grad(G, False)
        requires_grad(D, True)
        self.opt_d.zero_grad()
        fake = G(x, z).detach()                       # model.py:248,250
        d_real = D(real)                              # model.py:253
        d_fake = D(fake)                              # model.py:254
        gp = self.gp(D, real, fake, alpha=alpha)
        loss_d = ops.add(dis_loss_wgan(d_real, d_fake), gp)
        loss_d.backward()                             # model.py:259
        self.opt_d.step()                   # model.py:260
        return loss_d.detach(), gp.detach()

    def g_phase(self, x, z, real):
        G, D = self.G, self.D
        requires_grad(G, True)
        requires_grad(D, False)
        self.opt_g.zero_grad()
        fake = G(x, z)                                # model.py:271
        with torch.no_grad():
            D(real)                                   # model.py:274: result unused, BN buffers still advance
        loss_g = gen_loss_wgan(D(fake))               # model.py:275-276
        loss_g.backward()                             # model.py:278
        self.opt_g.step()                   # model.py:279
        return loss_g.detach()

    def step(self, x, z_d, z_g, real, alpha=None):
        """x [B,N,3] sphere, z_* [B,N,nz], real [B,3,N] (any strides).  Returns device scalars
        (loss_d, gp, loss_g); nothing here synchronises with the host."""
        loss_d, gp = self.d_phase(x, z_d, real, alpha)
        loss_g = self.g_phase(x, z_g, real)
        return loss_d, gp, loss_g
