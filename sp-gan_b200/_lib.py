"""ctypes binding of libspgan_b200.so.

Signatures are parsed from include/spgan_b200.h so the header stays the single source of
truth for the C ABI.  There is no fallback: if the CUDA library is missing or a symbol is
absent, importing this module raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspgan_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "spgan_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "float": ctypes.c_float,
    "size_t": ctypes.c_size_t, "spgan_stream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every `spgan_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"((?:const\s+)?\b[A-Za-z_0-9]+\s*\**)\s*\b(spgan_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _CTYPES[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPES[ty])
        protos[name] = (restype, argtypes)
    return protos


class SpganError(RuntimeError):
    pass


# kernel launches behind one C-ABI call (1 unless listed)
_LAUNCHES = {"spgan_knn_rows": 3, "spgan_colstats_bn": 2, "spgan_norm_bwd_reduce_acc": 2, "spgan_gemm_fused": 2, "spgan_colsum": 2, "spgan_coldot": 2, "spgan_colstats": 2, "spgan_norm_bwd_reduce": 2,
             "spgan_bn_dbl_bwd_reduce": 2, "spgan_bn_dbl_bwd_apply": 2, "spgan_gp_penalty": 2,
             "spgan_bn_pool_fwd": 2, "spgan_bn_pool_bwd": 2, "spgan_adam_step_dev": 2}


class _Library:
    """Counters: `launches` = kernels launched so far through this binding; `profile` (None or list):
    when a list, every call is bracketed by CUDA events on the current torch stream and appended as
    (name, int args, start_event, end_event) -- used by bench.py for per-kernel device times."""

    def __init__(self):
        self.launches = 0
        self.profile = None
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libspgan_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `python sp-gan_b200/build.py`; there is no CPU or PyTorch fallback." % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
            if restype is ctypes.c_int and name != "spgan_abi_version":
                setattr(self, name[len("spgan_"):], self._checked(name, fn))
            else:
                setattr(self, name[len("spgan_"):], fn)
        if self.abi_version() != 1:
            raise ImportError("libspgan_b200.so ABI version mismatch")

    def _checked(self, name, fn):
        err = self.cdll.spgan_error_string
        nl = _LAUNCHES.get(name, 1)

        def call(*args):
            prof = self.profile
            if prof is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = fn(*args)
            if rc != 0:
                raise SpganError("%s failed: %s (code %d)" % (name, err(rc).decode(), rc))
            self.launches += nl
            if prof is not None:
                e1.record()
                prof.append((name, tuple(a for a in args if isinstance(a, int) and abs(a) < (1 << 40)), e0, e1))
        call.__name__ = name
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Library()
    return _lib
