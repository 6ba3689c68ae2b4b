"""Host-side logic of the operator layer that needs no GPU: the scoped split-weight cache and the guard of the gradient
penalty's accumulation hand-over (sp-gan_b200/ops.py)."""
import torch
import torch.nn as nn


def _ops():
    import spgan_b200
    return spgan_b200.ops


def test_split_weight_cache_is_scoped_and_follows_versions():
    ops = _ops()
    W = nn.Parameter(torch.randn(8, 16))
    dev = torch.device("cpu")
    ops.weights_changed()
    # outside a scope: a fresh workspace every time, nothing remembered
    ws, flag = ops._cached_ws("gemm", W, True, None, 1024, dev)
    assert flag == 0 and len(ops._WCACHE) == 0
    with ops.weight_cache_scope():
        ws1, f1 = ops._cached_ws("gemm", W, True, None, 1024, dev)
        ws2, f2 = ops._cached_ws("gemm", W, True, None, 1024, dev)
        assert (f1, f2) == (0, 2) and ws2 is ws1                     # second product: the split is already there
        _, f3 = ops._cached_ws("gemm", W, True, ("other route",), 1024, dev)
        assert f3 == 0                                               # the layout (route) is part of the key
        view = W.view(8, 16, 1)[:, :, 0]                             # a view of the parameter maps to the parameter
        assert ops._param_of(view) is W
        with torch.no_grad():
            W.mul_(2.0)                                              # in-place update: version bump
        _, f4 = ops._cached_ws("gemm", W, True, None, 1024, dev)
        assert f4 == 0
        ops.weights_changed()                                        # raw-pointer update announced by the optimiser
        _, f5 = ops._cached_ws("gemm", W, True, None, 1024, dev)
        assert f5 == 0
        with ops.weight_cache_scope():                               # nested scopes share the cache
            _, f6 = ops._cached_ws("gemm", W, True, None, 1024, dev)
            assert f6 == 2
        assert len(ops._WCACHE) >= 1
        _, f7 = ops._cached_ws("gemm", W.detach().clone(), True, None, 1024, dev)
        assert f7 == 0                                               # not a parameter: never cached
    assert len(ops._WCACHE) == 0                                     # leaving the outermost scope drops everything
    # a parameter freed and another one allocated: entries are tied to the object, not to its address
    with ops.weight_cache_scope():
        P1 = nn.Parameter(torch.randn(4, 4))
        ops._cached_ws("gemm", P1, True, None, 64, dev)
        key_count = len(ops._WCACHE)
        del P1
        P2 = nn.Parameter(torch.randn(4, 4))
        _, f = ops._cached_ws("gemm", P2, True, None, 64, dev)
        assert f == 0 and len(ops._WCACHE) >= key_count


def test_hand_over_guard_asks_the_engine():
    """ops._node_will_run: True for a node the running backward pass will execute, False for one it will not reach (the
    gradient term is then returned the normal way), False outside a backward pass."""
    ops = _ops()
    seen = {}

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, node_yes, node_no):
            ctx.nodes = (node_yes, node_no)
            return a + 1

        @staticmethod
        def backward(ctx, g):
            seen["yes"] = ops._node_will_run(ctx.nodes[0])
            seen["no"] = ops._node_will_run(ctx.nodes[1])
            return g, None, None

    x = torch.ones(3, requires_grad=True)
    y = x * 2                      # reached by the backward below
    other = torch.ones(3, requires_grad=True) * 3      # a different graph
    z = Probe.apply(y, y.grad_fn, other.grad_fn)
    z.sum().backward()
    assert seen == {"yes": True, "no": False}
    assert ops._node_will_run(y.grad_fn) is False       # no backward pass is running
    assert ops._node_will_run(None) is False
