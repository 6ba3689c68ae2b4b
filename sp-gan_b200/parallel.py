"""Host-side data-parallel plumbing: flat parameter / gradient buffers and the one
all-reduce per phase.  Device-agnostic on purpose -- it only moves views and calls
torch.distributed -- so the N>1 logic is covered by world_size-2 gloo tests on CPU; every kernel
that touches the buffers lives in train_step.FlatAdam."""
import torch
import torch.distributed as dist


# Set False to run a trainer as a stand-alone replica inside an initialised process group (no broadcast, no
# all-reduce): the multi-rank parity test compares the data-parallel step against such replicas.
SYNC_ENABLED = True


def world_size():
    return dist.get_world_size() if SYNC_ENABLED and dist.is_available() and dist.is_initialized() else 1


def shard_bounds(total, rank, world):
    """Contiguous, equal shards of the global batch (64 clouds per GPU in BASELINE configs[3]);
    the global batch must divide evenly so that the mean of per-rank mean losses equals the
    global mean."""
    if total % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (total, world))
    per = total // world
    return rank * per, (rank + 1) * per


ALIGN_ELEMS = 64          # 256 bytes


class FlatBuffers:
    """Re-points every parameter of `module` (and its .grad) into two contiguous fp32 buffers.
    Layout = parameter registration order, so all ranks agree on offsets."""

    def __init__(self, module):
        self.params = list(module.parameters())
        if not self.params:
            raise ValueError("module has no parameters")
        dev = self.params[0].device
        # every parameter starts on a 256-byte boundary so that kernels keep their vectorised
        # (16-byte) access paths on the re-pointed weights; the padding stays zero forever
        # (zero gradient -> Adam leaves it at zero).
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
        self.numel = off
        self.flat_p = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.data.reshape(-1))      # one-time re-layout at construction
                p.data = self.flat_p[off:off + k].view(p.shape)
                p.grad = self.flat_g[off:off + k].view(p.shape)

    def rebind_grads(self):
        """Autograd keeps accumulating into .grad in place; re-attach any that were dropped."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)

    def allreduce_grads(self):
        """ONE collective per phase over the flat gradient buffer (sum; the 1/world factor is folded
        into the optimizer kernel).  Returns the factor to apply."""
        w = world_size()
        if w > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
        return 1.0 / w

    def broadcast_params(self, src=0):
        """Make replicas identical at start-up (DataParallel replicates rank 0's weights, model.py:79-84)."""
        if world_size() > 1:
            dist.broadcast(self.flat_p, src=src)
