"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat buffers alias the module's
parameters and gradients, one all-reduce yields the sum on every rank, shards are disjoint."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spgan_b200.parallel import FlatBuffers, shard_bounds, world_size
    torch.manual_seed(7 + rank)                       # replicas start different on purpose
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    fb = FlatBuffers(net)
    ok = world_size() == world and fb.numel >= sum(p.numel() for p in net.parameters()) and all(o % 64 == 0 for o in fb.offsets)
    fb.broadcast_params(0)
    ref = [torch.empty_like(fb.flat_p) for _ in range(world)]
    dist.all_gather(ref, fb.flat_p)
    ok = ok and all(torch.equal(r, ref[0]) for r in ref)
    # parameters are views of the flat buffer
    net[0].weight.data.add_(1.0)
    ok = ok and torch.equal(fb.flat_p[:35].view(7, 5), net[0].weight.data)
    # autograd accumulates into the flat gradient buffer in place
    lo, hi = shard_bounds(8, rank, world)
    x = torch.arange(8 * 5, dtype=torch.float32).view(8, 5)[lo:hi] / 10
    net(x).sum().backward()
    ok = ok and fb.flat_g.abs().sum() > 0 and net[2].bias.grad.data_ptr() == fb.flat_g.data_ptr() + 4 * fb.offsets[-1]
    local = fb.flat_g.clone()
    both = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(both, local)
    scale = fb.allreduce_grads()
    ok = ok and abs(scale - 1.0 / world) < 1e-12 and torch.allclose(fb.flat_g, sum(both))
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_flat_buffers_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == (0, 4) and res[1][2] == (4, 8)


def test_shard_bounds_rejects_ragged():
    from spgan_b200.parallel import shard_bounds
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
