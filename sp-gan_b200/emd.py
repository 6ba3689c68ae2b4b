"""Approximate Earth Mover's Distance for the evaluation path: the reference's `emdModule`
(metrics/emd/emd_module.py:33-71, auction kernels metrics/emd/emd_cuda.cu:95-316) and `emd_approx`
(Common/GAN_metrics.py:396-409) on one sm_100a kernel per batch of cloud pairs (csrc/emd.cu).

Same call signature and outputs: `emdModule()(xyz1, xyz2, eps, iters) -> (dist [B, n], assignment [B, n] int32)`,
gradient for xyz1 only (the assignment is held fixed, xyz2 receives zeros like the reference).  Relaxed: n need not be
a multiple of 1024 and B is not limited to 512; required: equal sizes, n <= ~5200, iters >= 1.
Results are bit-identical to oracle/emd_recipe.c; parity with the reference binary itself is unpinned (no CPU path,
test or golden vector exists for it, see the oracle's header).
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3:
            raise ValueError("emd: expected [B, n, 3] clouds")
        if xyz1.shape[1] != xyz2.shape[1] or xyz1.shape[0] != xyz2.shape[0]:     # emd_module.py:39-40
            raise AssertionError("emd: the two batches must have the same shape")
        a, b = ops._c(xyz1.detach().float(), "xyz1"), ops._c(xyz2.detach().float(), "xyz2")
        B, n, _ = a.shape
        dist = torch.empty((B, n), device=a.device, dtype=torch.float32)
        assignment = torch.empty((B, n), device=a.device, dtype=torch.int32)
        ops.L().emd_auction(a.data_ptr(), b.data_ptr(), B, n, float(eps), int(iters), dist.data_ptr(),
                            assignment.data_ptr(), ops._stream())
        ctx.save_for_backward(a, b, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    @once_differentiable
    def backward(ctx, graddist, gradidx):
        a, b, assignment = ctx.saved_tensors
        g = ops._c(graddist)
        B, n, _ = a.shape
        ga = torch.empty_like(a)
        ops.L().emd_grad(a.data_ptr(), b.data_ptr(), g.data_ptr(), assignment.data_ptr(), B, n, ga.data_ptr(),
                         ops._stream())
        return ga, ops.full(tuple(b.shape), 0.0, b.device), None, None


class emdModule(nn.Module):
    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)


def emd_approx(sample, ref, eps=0.005, iters=300):
    """Common/GAN_metrics.py:396-409: mean over the batch and the points of the matched squared distances."""
    dist, _ = emdFunction.apply(sample, ref, eps, iters)
    return ops.MeanScale.apply(dist, 1.0)
