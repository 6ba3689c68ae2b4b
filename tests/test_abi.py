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


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/spgan_b200.h compiles as C99 (no C++ or torch types in the signatures) and a
    C translation unit links against the library (symbols resolved by name, nothing mangled)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "use.c"
    src.write_text('#include "spgan_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    size_t ws = spgan_gemm_workspace(3, 128, 64);\n'
                   '    printf("%d %s %zu\\n", spgan_abi_version(), spgan_error_string(-2), ws);\n'
                   '    return spgan_abi_version() == 1 ? 0 : 1;\n}\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    m = _lib_mod()
    exe = tmp_path / "use"
    libdir = os.path.dirname(m.LIB_PATH)
    rc = subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:libspgan_b200.so",
                         "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("1 "), (out.stdout, out.stderr)
