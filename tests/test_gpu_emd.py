"""GPU parity of the auction-EMD kernel against oracle/emd_recipe.c: dist and assignment bit for bit
(same arithmetic, same tie rules), the gradient, convergence to a near-optimal bijection, argument errors.
"""
import numpy as np
import pytest
import torch

from oracle import emd_ref

pytestmark = pytest.mark.gpu


def _emd():
    from spgan_b200 import emd
    return emd


def clouds(seed, B, n):
    rng = np.random.default_rng(seed)
    return rng.random((B, n, 3), dtype=np.float32), rng.random((B, n, 3), dtype=np.float32)


@pytest.mark.parametrize("B,n,eps,iters", [(3, 256, 0.005, 50), (2, 1024, 0.005, 50), (2, 2048, 0.005, 20),
                                           (4, 100, 0.002, 2000), (1, 33, 0.01, 5), (2, 1500, 0.005, 3)])
def test_matches_oracle_bit_for_bit(B, n, eps, iters):
    a, b = clouds(10 + n, B, n)
    dist, ass = _emd().emdModule()(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), eps, iters)
    rd, ra = emd_ref.emd(a, b, eps, iters)
    assert ass.dtype == torch.int32 and tuple(dist.shape) == (B, n)
    assert np.array_equal(ass.cpu().numpy(), ra)
    assert np.array_equal(dist.cpu().numpy(), rd)


def test_converged_result_is_a_bijection_and_identical_clouds_cost_nothing():
    a, b = clouds(3, 2, 512)
    # eps = 0.005 needs ~5000 iterations on the second pair (the oracle's trace still shows 2 bidders at 3000)
    dist, ass = _emd().emdModule()(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 0.005, 10000)
    rd, ra, trace = emd_ref.emd(a, b, 0.005, 10000, return_trace=True)
    assert (trace[:, -1] == 0).all()
    assert np.array_equal(ass.cpu().numpy(), ra) and np.array_equal(dist.cpu().numpy(), rd)
    for i in range(2):
        assert sorted(ass[i].cpu().tolist()) == list(range(512))
    d0, a0 = _emd().emdModule()(torch.from_numpy(a).cuda(), torch.from_numpy(a).cuda(), 0.005, 50)
    assert not d0.any() and torch.equal(a0.cpu(), torch.arange(512, dtype=torch.int32).repeat(2, 1))


def test_gradient_and_emd_approx():
    a, b = clouds(4, 2, 256)
    x = torch.from_numpy(a).cuda().requires_grad_()
    y = torch.from_numpy(b).cuda().requires_grad_()
    dist, ass = _emd().emdModule()(x, y, 0.005, 100)
    w = torch.rand(2, 256, device="cuda")
    (dist * w).sum().backward()
    ref = emd_ref.emd_grad(a, b, w.cpu().numpy(), ass.cpu().numpy())
    assert np.allclose(x.grad.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    assert not y.grad.any()
    val = _emd().emd_approx(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 0.005, 100)
    rd, _ = emd_ref.emd(a, b, 0.005, 100)
    assert abs(float(val) - float(rd.astype(np.float64).mean())) <= 1e-5 * float(rd.mean())


def test_argument_errors():
    e = _emd()
    a = torch.rand(1, 64, 3, device="cuda")
    with pytest.raises(AssertionError):
        e.emdModule()(a, torch.rand(1, 32, 3, device="cuda"), 0.005, 10)
    with pytest.raises(Exception):
        e.emdModule()(a, a, 0.005, 0)                      # iters >= 1
    with pytest.raises(Exception):
        e.emdModule()(torch.rand(1, 8000, 3, device="cuda"), torch.rand(1, 8000, 3, device="cuda"), 0.005, 5)
    with pytest.raises(RuntimeError):
        e.emdModule()(torch.rand(1, 64, 3), torch.rand(1, 64, 3), 0.005, 10)       # CPU tensors: no fallback


def test_kernel_matches_the_reference_binary_golden():
    """spgan_emd_auction against dist / assignment of the reference's own kernels (metrics/emd/emd_cuda.cu compiled
    unmodified, run on a B200: tests/golden/emd_reference.npz), with the deviation rule of tests/test_oracle_emd.py."""
    from test_oracle_emd import check_against_reference_golden

    def run(a, b, eps, iters):
        d, s = _emd().emdModule()(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), eps, iters)
        return d.cpu().numpy(), s.cpu().numpy()
    check_against_reference_golden(run)
