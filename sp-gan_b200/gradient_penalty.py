"""WGAN-GP gradient penalty with the reference API (Common/gradient_penalty.py:4-37)."""
import torch

from . import ops


class GradientPenalty:
    """lambdaGP * mean(((||d netD(mix) / d mix||_2 - gamma) / gamma)^2), mix = real + alpha (fake - real).

    `alpha` ~ U[0,1) of shape [B,1,1] is drawn from the CPU torch generator exactly like the
    reference (gradient_penalty.py:24); pass `alpha=` to inject it (parity tests).  The returned
    scalar is differentiable w.r.t. netD's parameters (double backward through the critic)."""

    def __init__(self, lambdaGP, gamma=1, vertex_num=2500, device=None):
        self.lambdaGP = lambdaGP
        self.gamma = gamma
        self.vertex_num = vertex_num
        self.device = device

    def __call__(self, netD, real_data, fake_data, alpha=None):
        batch_size = real_data.size(0)
        fake_data = fake_data[:batch_size]
        if alpha is None:
            alpha = torch.rand(batch_size, 1, 1)
        alpha = alpha.detach().to(device=real_data.device, dtype=torch.float32).reshape(-1)
        mix = ops.gp_interpolate(real_data.detach(), fake_data.detach(), alpha).requires_grad_(True)
        with ops.twice_differentiable():
            out = netD(mix)
        ones = ops.full(tuple(out.shape), 1.0, out.device)
        with ops.input_grad_only():
            (grads,) = torch.autograd.grad(outputs=out, inputs=mix, grad_outputs=ones, create_graph=True,
                                           retain_graph=True, only_inputs=True)
        return ops.GradPenalty.apply(grads.reshape(batch_size, -1), float(self.gamma), float(self.lambdaGP))
