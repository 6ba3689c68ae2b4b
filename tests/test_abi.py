"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/spgan_b200.h declares; argument validation returns codes instead of crashing."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _lib_mod():
    import spgan_b200._lib as m
    return m


def test_library_exports_every_declared_symbol():
    m = _lib_mod()
    protos = m.parse_header()
    assert len(protos) >= 50
    cdll = ctypes.CDLL(m.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, missing
    hdr = open(os.path.join(ROOT, "include", "spgan_b200.h")).read()
    declared = set(re.findall(r"\b(spgan_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(protos), declared ^ set(protos)


def test_abi_version_and_error_strings():
    L = _lib_mod().lib()
    assert L.abi_version() == 1
    assert b"bad argument" in L.cdll.spgan_error_string(-1)
    assert b"envelope" in L.cdll.spgan_error_string(-2)


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call."""
    m = _lib_mod()
    L = m.lib()
    with pytest.raises(m.SpganError):
        L.knn_group(None, None, 1, 3, 16, 4, None, None, None)
    with pytest.raises(m.SpganError):
        L.gemm(0, 0, 4, 4, 4, None, 4, None, 4, None, 4, None, 0, 0, None, 0, None)
    assert L.cdll.spgan_knn_group(1, 1, 1, 3, 64, 40, 1, None, None) == -2     # k + 1 > 32
    assert L.cdll.spgan_knn_group(1, 1, 1, 3, 4, 4, 1, None, None) == -1      # k + 1 > N
    assert L.colreduce_workspace(0, 4, 1, 1) == 0


def test_no_oracle_or_fallback_in_product():
    """The product path must not import oracle/ nor contain a CPU fallback."""
    pkg = os.path.join(ROOT, "sp-gan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
