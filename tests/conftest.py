import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SUBSAMPLE_ABOVE = 32768      # keep in sync with tests/golden/make_golden.py::pack
SUBSAMPLE_STRIDE = 16


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def pack_like_golden(a):
    """Mirror of make_golden.pack for comparing against stored (possibly subsampled) tensors."""
    import torch
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    if a.size > SUBSAMPLE_ABOVE:
        a = np.ascontiguousarray(a.reshape(-1)[::SUBSAMPLE_STRIDE])
    return a


def rel_err(a, b):
    """(max-abs error / max-abs reference, L2 error / L2 reference)."""
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.abs(a - b)
    return d.max() / (np.abs(b).max() + 1e-30), np.sqrt((d ** 2).sum()) / (np.sqrt((b ** 2).sum()) + 1e-30)


def assert_rel(a, b, tol, what="", max_tol=None):
    """The parity bar of BASELINE.json: features/losses/grads within `tol` relative (fp32).
    Both the max-abs error relative to the largest reference magnitude and the relative L2 error are
    checked; `max_tol` loosens only the former (gradients through LeakyReLU: one pre-activation within
    an ulp of zero flips its mask and moves single elements by O(1), see DESIGN.md "Parity")."""
    a = pack_like_golden(a) if np.asarray(b).size != np.asarray(a.detach().cpu() if hasattr(a, "detach") else a).size else a
    import torch
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    emax, el2 = rel_err(a, b)
    mt = tol if max_tol is None else max_tol
    assert emax <= mt and el2 <= tol, "%s: rel max err %.3e (tol %.1e), rel L2 err %.3e (tol %.1e)" % (
        what, emax, mt, el2, tol)


@pytest.fixture(scope="session")
def sphere2048():
    return np.load(os.path.join(GOLDEN, "sphere_2048.npy"))


@pytest.fixture(scope="session")
def sphere256():
    return np.load(os.path.join(GOLDEN, "sphere_256.npy"))
