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


def _dense_copy_(dst, src):
    """dst.copy_(src) as ONE memcpy when both are the same permutation of a dense tensor (e.g. the transposed
    [B,3,N] view of a [B,N,3] batch, model.py:249): keeps pinned-host -> device copies asynchronous."""
    order = sorted(range(dst.dim()), key=lambda i: -dst.stride(i))
    d, s_ = dst.permute(order), src.permute(order)
    if d.is_contiguous() and s_.is_contiguous():
        d.copy_(s_, non_blocking=True)
    else:
        dst.copy_(src, non_blocking=True)


class FlatAdam:
    """torch.optim.Adam semantics (model.py:94-97) as ONE fused kernel over the flat parameter buffer,
    preceded by ONE all-reduce of the flat gradient buffer when running data parallel.  The step count lives
    on the device (spgan_adam_step_dev), so a captured CUDA graph of the step can be replayed."""

    def __init__(self, module, lr=1e-4, betas=(0.5, 0.99), eps=1e-8):
        self.buf = FlatBuffers(module)
        n, dev = self.buf.numel, self.buf.flat_p.device
        self.m = ops.full((n,), 0.0, dev)
        self.v = ops.full((n,), 0.0, dev)
        self.state = torch.zeros(4, device=dev, dtype=torch.int32)      # {step, bias corrections}: see the header
        self.lr, self.betas, self.eps, self.n = lr, betas, eps, n

    @property
    def t(self):
        return int(self.state[0])          # host read-back (synchronises): diagnostics only

    def zero_grad(self):
        ops.fill_(self.buf.flat_g, 0.0)
        self.buf.rebind_grads()

    def step(self):
        scale = self.buf.allreduce_grads()
        ops.weights_changed()              # the update below goes through raw pointers: drop the split-weight cache
        ops.L().adam_step_dev(self.buf.flat_p.data_ptr(), self.buf.flat_g.data_ptr(), self.m.data_ptr(),
                              self.v.data_ptr(), self.n, self.lr, self.betas[0], self.betas[1], self.eps,
                              self.state.data_ptr(), scale, ops._stream())


class WGANGPTrainer:
    """One object per rank.  step(x, z_d, z_g, real) runs the D phase then the G phase."""

    def __init__(self, G, D, lambda_gp=10.0, gamma=1.0, lr_g=1e-4, lr_d=1e-4, betas=(0.5, 0.99)):
        self.G, self.D = G, D
        self.gp = GradientPenalty(lambda_gp, gamma=gamma)
        self.opt_g = FlatAdam(G, lr_g, betas)
        self.opt_d = FlatAdam(D, lr_d, betas)
        self.opt_g.buf.broadcast_params()
        self.opt_d.buf.broadcast_params()
        ops.weights_changed()              # parameters were re-pointed / broadcast: no split-weight cache entry survives
        self._graph, self._graph_out, self._static, self.graph_launches = None, None, None, 0

    def d_phase(self, x, z, real, alpha=None):
        with ops.weight_cache_scope():
            return self._d_phase(x, z, real, alpha)

    def g_phase(self, x, z, real):
        with ops.weight_cache_scope():
            return self._g_phase(x, z, real)

    def _d_phase(self, x, z, real, alpha=None):
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

    def _g_phase(self, x, z, real):
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
        with ops.weight_cache_scope():         # one scope over both phases: G's weights are split once per step
            loss_d, gp = self.d_phase(x, z_d, real, alpha)
            loss_g = self.g_phase(x, z_g, real)
        return loss_d, gp, loss_g

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def capture(self, x, z_d, z_g, real, alpha, warmup=2):
        """Record one full step (both phases, the gradient all-reduces and the Adam updates: ~700 kernel
        launches) into a CUDA graph over static input buffers.  `warmup` eager steps run first on the capture
        stream (they DO train, like any other step).  Afterwards `replay(...)` costs one graph launch of host
        time per step.  All arguments must be CUDA tensors; their shapes / strides become part of the graph."""
        if self._graph is not None:
            raise RuntimeError("a step graph has already been captured")
        self._static = {"x": x, "z_d": z_d.clone(memory_format=torch.preserve_format) if z_d.stride(1) else z_d[:, :1].clone().expand_as(z_d),
                        "z_g": z_g.clone(memory_format=torch.preserve_format) if z_g.stride(1) else z_g[:, :1].clone().expand_as(z_g),
                        "real": real.clone(memory_format=torch.preserve_format), "alpha": alpha.detach().clone()}
        st = self._static
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(st["x"], st["z_d"], st["z_g"], st["real"], st["alpha"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.weights_changed()              # nothing split outside the graph may be referenced from inside it
        l0 = ops.L().launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.step(st["x"], st["z_d"], st["z_g"], st["real"], st["alpha"])
        self.graph_launches = ops.L().launches - l0          # kernels recorded in the graph
        self._graph, self._graph_out = g, out
        return self

    def release_graph(self):
        """Drop the captured step graph and its static buffers (call before destroying the process group: the
        graph holds the two NCCL all-reduces)."""
        self._graph, self._graph_out, self._static = None, None, None

    def replay(self, z_d=None, z_g=None, real=None, alpha=None):
        """Copy the new inputs into the static buffers (any of them may be omitted = unchanged) and replay the
        captured step.  Returns the static (loss_d, gp, loss_g) device scalars, overwritten by the next replay."""
        st = self._static
        for key, val in (("z_d", z_d), ("z_g", z_g), ("real", real), ("alpha", alpha)):
            if val is not None:
                dst = st[key]
                if dst.dim() == 3 and dst.stride(1) == 0:                 # broadcast latent: one vector per cloud
                    dst[:, :1].copy_(val[:, :1], non_blocking=True)
                else:
                    _dense_copy_(dst, val)
        self._graph.replay()
        return self._graph_out
