"""Differentiable operators over libspgan_b200 (torch.autograd.Function wrappers).

PyTorch supplies device memory (the caching allocator), the current CUDA stream and the
autograd graph; every arithmetic operation below is a kernel of libspgan_b200.so called
through the C ABI of include/spgan_b200.h.  The set used by the critic is closed under
differentiation (the backward of each op is built from ops of the same set), which is what
the gradient penalty's create_graph=True pass needs (Common/gradient_penalty.py:31-33).

Layout: activations are point-major rows [R, C] (R = B*N points or B*N*k edges).
"""
from __future__ import annotations

import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._lib import lib

_L = None


def L():
    global _L
    if _L is None:
        _L = lib()
    return _L


# Set by GradientPenalty while it differentiates w.r.t. the interpolates only
# (gradient_penalty.py:31-33, only_inputs=True): parameter gradients are not requested there.
_INPUT_GRAD_ONLY = False

# engine for spgan_gemm: 0 = fp32 CUDA cores; 1 = tcgen05 tensor cores with the fp32-faithful TF32x3 split;
# 2 = tcgen05 with the faster bf16x3 split (~2^-16 per product, not parity-safe); 3 (default) = tcgen05 with the
# fp16x3 split with scaled residuals (22 significant bits like TF32x3, twice its MMA rate): K <= 256 products run
# on the TMEM-resident-A kernel (csrc/gemm_ts.cu), the rest on the streaming kernel (csrc/gemm_tc.cu).
# Override with SPGAN_GEMM_ENGINE.
import os as _os
GEMM_ENGINE = int(_os.environ.get("SPGAN_GEMM_ENGINE", "3"))


# Set by GradientPenalty around netD(interpolates): the critic then records the unfused,
# twice-differentiable operator chain instead of the fused first-order one.
_TWICE_DIFFERENTIABLE = False


class twice_differentiable:
    def __enter__(self):
        global _TWICE_DIFFERENTIABLE
        self.prev = _TWICE_DIFFERENTIABLE
        _TWICE_DIFFERENTIABLE = True

    def __exit__(self, *a):
        global _TWICE_DIFFERENTIABLE
        _TWICE_DIFFERENTIABLE = self.prev


def want_twice_differentiable():
    return _TWICE_DIFFERENTIABLE


class input_grad_only:
    def __enter__(self):
        global _INPUT_GRAD_ONLY
        self.prev = _INPUT_GRAD_ONLY
        _INPUT_GRAD_ONLY = True

    def __exit__(self, *a):
        global _INPUT_GRAD_ONLY
        _INPUT_GRAD_ONLY = self.prev


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError("spgan_b200: %s must live on a CUDA device (no CPU fallback exists)" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("spgan_b200: %s must be float32, got %s" % (name, t.dtype))
    return t


def _c(t, name="tensor"):
    """Contiguous fp32 CUDA tensor (copies through our own kernel path only when needed)."""
    _chk(t, name)
    return t if t.is_contiguous() else contiguous(t)


def contiguous(t):
    """Materialise a strided <=3-D view with the layout kernel (no torch compute)."""
    if t.is_contiguous():
        return t
    if t.dim() == 2:
        t3 = t.unsqueeze(0)
    elif t.dim() == 3:
        t3 = t
    else:
        raise RuntimeError("spgan_b200.contiguous: unsupported rank %d" % t.dim())
    # treat as [B, C, N] -> rows [B*N, C] of the transposed view, i.e. copy with strides
    B, C, N = t3.shape
    out = torch.empty((B, C, N), device=t.device, dtype=t.dtype)
    # rows_to view: write out[b, c, n] = t3[b, c, n]; use bcn_to_rows on the (b, n, c) permutation
    tp = t3.permute(0, 2, 1)                          # [B, N, C] view, we want out as [B, C*?]
    L().bcn_to_rows(tp.data_ptr(), tp.stride(0), tp.stride(1), tp.stride(2), B, N, C, out.data_ptr(), _stream())
    return out.view(t.shape)


def _ws(R, C, seg_rows, nvals, device):
    n = L().colreduce_workspace(R, C, seg_rows, nvals)
    return torch.empty((max(n, 4) + 3) // 4, device=device, dtype=torch.float32)


def _rows2d(t):
    if t.dim() != 2:
        raise RuntimeError("expected a [rows, channels] matrix, got shape %s" % (tuple(t.shape),))
    return t


def _ld(t):
    """Leading dimension of a 2-D row-major operand view (unit column stride required)."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        return None
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return ld if ld >= t.shape[1] else None


def _gemm_operand(t):
    _chk(t)
    if _ld(t) is None:
        t = contiguous(t)
    return t, _ld(t)


# =========================================================================================
# dense contraction
# =========================================================================================
LAST_TC_WORKSPACE = None      # most recent tcgen05 workspace (its first int is the kernel's status word)
_TN_STATUS = {}               # per-device zeroed status block of the weight-gradient kernel


# Split-weight cache: the tensor engines split every fp32 weight into fp16 hi / lo tiles before the product.  Within
# one optimiser phase the same weight is multiplied many times (the critic runs five times per step), so the split
# workspace of a PARAMETER is kept and re-used (transB | 2 in the C ABI) until the weights change: every in-place
# update bumps the tensor version, FlatAdam (raw-pointer update) calls weights_changed().  Only inside a
# weight_cache_scope (below).
_WCACHE = {}
WEIGHT_EPOCH = 0
WEIGHT_CACHE = _os.environ.get("SPGAN_WEIGHT_CACHE", "1") != "0"


def weights_changed():
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1
    _WCACHE.clear()


_WC_DEPTH = 0


class weight_cache_scope:
    """The split-weight cache is only live inside such a scope, opened by a caller that OWNS the parameter updates
    (train_step.WGANGPTrainer: every update inside goes through FlatAdam, which calls weights_changed()).  Outside a
    scope every product splits its weight afresh: a bare module stays correct under any way of writing its weights,
    including `param.data.copy_()` / raw-pointer writes that bump no version counter.  Entering the outermost scope
    drops whatever was cached before (the weights may have been touched in between); leaving it drops everything."""

    def __enter__(self):
        global _WC_DEPTH
        if _WC_DEPTH == 0:
            weights_changed()
        _WC_DEPTH += 1
        return self

    def __exit__(self, *exc):
        global _WC_DEPTH
        _WC_DEPTH -= 1
        if _WC_DEPTH == 0:
            _WCACHE.clear()
        return False


def _param_of(t):
    """The nn.Parameter `t` is (a view of), or None."""
    if isinstance(t, torch.nn.Parameter):
        return t
    base = t._base
    return base if base is not None and isinstance(base, torch.nn.Parameter) else None


def _cached_ws(kind, Bm, tb, extra, ws_bytes, device):
    """-> (workspace, flag): flag 2 = the workspace already holds the split of this weight.  Entries are tied to the
    Parameter OBJECT (weak reference), not to its address: a freed parameter's address, shape and version can all
    recur in another module."""
    param = _param_of(Bm) if _WC_DEPTH > 0 else None
    if param is None:
        return torch.empty(ws_bytes // 4 + 64, device=device, dtype=torch.float32), 0
    key = (kind, id(param), Bm.data_ptr(), Bm._version, tuple(Bm.shape), tuple(Bm.stride()), bool(tb), extra, WEIGHT_EPOCH)
    ent = _WCACHE.get(key)
    if ent is not None and ent[1]() is param:
        return ent[0], 2
    ws = torch.empty(ws_bytes // 4 + 64, device=device, dtype=torch.float32)
    if len(_WCACHE) > 512:
        _WCACHE.clear()
    _WCACHE[key] = (ws, weakref.ref(param))
    return ws, 0


def gemm_raw(A, B, bias=None, ta=False, tb=False, out=None, accumulate=False, engine=None, wcache=False):
    global LAST_TC_WORKSPACE
    A, lda = _gemm_operand(A)
    B, ldb = _gemm_operand(B)
    M = A.shape[1] if ta else A.shape[0]
    K = A.shape[0] if ta else A.shape[1]
    Kb = B.shape[1] if tb else B.shape[0]
    N = B.shape[0] if tb else B.shape[1]
    if K != Kb:
        raise RuntimeError("gemm: inner dimensions differ (%d vs %d)" % (K, Kb))
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    if bias is not None:
        bias = _c(bias)
    engine = GEMM_ENGINE if engine is None else engine
    ws, ws_bytes, wflag = None, 0, 0
    tc = engine in (1, 2, 3) and ((not ta and M >= 128 and N >= 16 and K >= 16) or
                               (ta and not tb and bias is None and K >= 4096 and M >= 16 and N >= 16))
    if tc and ta:
        # weight-gradient (TN) kernel: converts both operands on the fly, needs only the 256-byte status block
        # (sizing this from gemm_workspace(N, K) with K = the ROW count asked for gigabytes per backward)
        if engine == 3:
            # status block + the deterministic split-K partial tiles of csrc/gemm_wg.cu (a few MB)
            ws_bytes = L().gemm_wgrad_workspace(M, N, K)
            ws = torch.empty(ws_bytes // 4, device=A.device, dtype=torch.float32)
        else:
            ws_bytes = 256
            ws = _TN_STATUS.get(A.device)
            if ws is None:
                ws = _TN_STATUS[A.device] = full((64,), 0.0, A.device)   # stays zero: only a (trapping) timeout writes it
        LAST_TC_WORKSPACE = ws
    elif tc or (M <= 128 and 256 <= K < 2048):       # pre-split weight operand / small-batch split-K partial tiles
        ws_bytes = L().gemm_workspace(engine, N, K)
        if tc and wcache and WEIGHT_CACHE:
            # the route (gemm_ts vs gemm_tc) is part of the key: it fixes the layout of the split
            route = (engine == 3 and L().gemm_fused_workspace(M, N, K, A.data_ptr(), lda) != 0,
                     engine == 3 and not ta and L().gemm_bigk_route(M, N, K, A.data_ptr(), lda) != 0)
            ws, wflag = _cached_ws("gemm", B, tb, (engine, route), ws_bytes, A.device)
        else:
            ws = torch.empty(ws_bytes // 4 + 64, device=A.device, dtype=torch.float32)
        if tc:
            LAST_TC_WORKSPACE = ws
    L().gemm(int(ta), int(tb) | wflag, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, out.data_ptr(), _ld(out),
             bias.data_ptr() if bias is not None else None, int(accumulate), engine,
             ws.data_ptr() if ws is not None else None, ws_bytes, _stream())
    return out


def gemm_fused_raw(A, B, bias=None, tb=True, out=None, accumulate=False, a_scale=None, a_shift=None, a_slope=1.0,
                   want_stats=False, wcache=False):
    """spgan_gemm_fused: C = lrelu(A * a_scale + a_shift) @ op(B) + bias, optionally with the per-column partial sums
    of C and C^2 (-> (C, col_sum, col_sqsum)).  Returns None when the shape is outside the fused kernel's envelope."""
    global LAST_TC_WORKSPACE
    A, lda = _gemm_operand(A)
    B, ldb = _gemm_operand(B)
    M, K = A.shape
    N = B.shape[0] if tb else B.shape[1]
    ws_bytes = L().gemm_fused_workspace(M, N, K, A.data_ptr(), lda)
    if ws_bytes == 0:
        return None
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    wflag = 0
    if wcache and WEIGHT_CACHE:
        ws, wflag = _cached_ws("fused", B, tb, None, ws_bytes, A.device)
    else:
        ws = torch.empty(ws_bytes // 4 + 64, device=A.device, dtype=torch.float32)
    LAST_TC_WORKSPACE = ws
    cs = cq = None
    if want_stats:
        rows = L().gemm_fused_stats_rows(M)
        cs = torch.empty((rows, N), device=A.device, dtype=torch.float32)
        cq = torch.empty((rows, N), device=A.device, dtype=torch.float32)
    L().gemm_fused(int(tb) | wflag, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, out.data_ptr(), _ld(out),
                   bias.data_ptr() if bias is not None else None, int(accumulate),
                   a_scale.data_ptr() if a_scale is not None else None,
                   a_shift.data_ptr() if a_shift is not None else None, float(a_slope),
                   cs.data_ptr() if cs is not None else None, cq.data_ptr() if cq is not None else None,
                   ws.data_ptr(), ws_bytes, _stream())
    return (out, cs, cq) if want_stats else out


def _wmat(B, cols):
    """Weight operand as a matrix: [Cout, Cin, 1(, 1)] parameters are viewed as [Cout, K]; `cols` = (c0, c1) selects a
    column block (the two halves of an EdgeConv weight, the global / local halves of tail[0])."""
    m = B if B.dim() == 2 else B.reshape(B.shape[0], -1)
    return m if cols is None else m[:, cols[0]:cols[1]]


# Weight gradients of leaf parameters whose .grad already exists (the flat gradient buffer of train_step.FlatAdam)
# are ACCUMULATED IN PLACE by the weight-gradient GEMM (C += A^T B) and the Function returns None for them: autograd's
# AccumulateGrad would otherwise launch one at::add per parameter and backward pass (143 per step, 2.4 % of it) and,
# for column blocks of a weight, a zero-fill + copy for the slice.  Off under create_graph (double backward).
DIRECT_WEIGHT_GRAD = _os.environ.get("SPGAN_DIRECT_WEIGHT_GRAD", "1") != "0"


class Gemm(Function):
    """C = op(A) @ op(B) + bias.  Closed under differentiation (backward = three Gemm/ColSum).
    `engine` (None = module default) pins the arithmetic of the FORWARD product only: the per-point
    projections that EdgeCombine differences (pn[j] - pn[p]) use engine 0 (exact fp32 FMA chains).
    B may be a conv / linear weight of any rank (viewed as [Cout, K]) and `cols` a column block of it."""

    @staticmethod
    def forward(ctx, A, B, bias, ta, tb, engine=None, zero_bias_grad=False, cols=None):
        ctx.ta, ctx.tb, ctx.cols = ta, tb, cols
        ctx.save_for_backward(A, B)
        ctx.has_bias = bias is not None
        ctx.zero_bias_grad = zero_bias_grad
        return gemm_raw(A, _wmat(B, cols), bias, ta, tb, engine=engine, wcache=True)

    @staticmethod
    def backward(ctx, g):
        A, B = ctx.saved_tensors
        ta, tb, cols = ctx.ta, ctx.tb, ctx.cols
        dA = dB = db = None
        params_too = not _INPUT_GRAD_ONLY
        Bm = _wmat(B, cols)
        if ctx.needs_input_grad[0]:
            # C = op(A) op(B):  d op(A) = g op(B)^T
            dA = Gemm.apply(g, Bm, None, False, not tb) if not ta else Gemm.apply(Bm, g, None, tb, True)
        if ctx.needs_input_grad[1] and params_too:
            if (DIRECT_WEIGHT_GRAD and not torch.is_grad_enabled() and B.is_leaf and B.grad is not None
                    and B.grad.is_contiguous() and B.grad.dtype == torch.float32 and B.grad.is_cuda):
                gm = _wmat(B.grad, cols)                     # view of the parameter's gradient (block)
                if not tb:
                    gemm_raw(A, g, None, not ta, False, out=gm, accumulate=True)
                else:
                    gemm_raw(g, A, None, True, ta, out=gm, accumulate=True)
            else:
                dB = Gemm.apply(A, g, None, not ta, False) if not tb else Gemm.apply(g, A, None, True, ta)
                if cols is not None or B.dim() != 2:         # back to the parameter's own shape
                    full_ = dB
                    if cols is not None:
                        full_ = full((Bm.shape[0], _wmat(B, None).shape[1]), 0.0, dB.device)
                        full_[:, cols[0]:cols[1]] = dB
                    dB = full_.reshape(B.shape)
        if ctx.has_bias and ctx.needs_input_grad[2] and params_too and not ctx.zero_bias_grad:
            db = ColSum.apply(g, g.shape[0]).view(-1)
        return dA, dB, db, None, None, None, None, None


# A bias added right before a train-mode BatchNorm has an exactly zero gradient (the batch mean removes it; the
# gradient reaching the conv sums to zero over the batch): the reference computes rounding noise there (~1e-7 of the
# layer's gradient scale).  With this switch on, those gradients are returned as exact zeros instead of spending a
# full read of the [rows, C] gradient tensor on the noise (SPGAN_EXACT_ZERO_BIAS_GRAD=0 restores the reduction).
# "Exact zero" = no gradient is returned for that bias (its .grad keeps the zeros of zero_grad).
EXACT_ZERO_BIAS_GRAD = _os.environ.get("SPGAN_EXACT_ZERO_BIAS_GRAD", "1") != "0"


def feeds_train_bn(bn):
    """True when a bias added just before `bn` cannot receive gradient: train-mode batch statistics."""
    return EXACT_ZERO_BIAS_GRAD and (bn.training or not bn.track_running_stats)


def linear(x, weight, bias=None, engine=None, zero_bias_grad=False, cols=None):
    """x [R, Cin] @ weight[Cout, Cin(,1(,1))]^T + bias -- Conv1d(k=1) / Conv2d(1x1) / Linear; `cols` = (c0, c1)
    restricts the product to the input-channel block weight[:, c0:c1]."""
    return Gemm.apply(x, weight, bias, False, True, engine, zero_bias_grad, cols)


# =========================================================================================
# column reductions / broadcasts (mutually adjoint)
# =========================================================================================
class ColSum(Function):
    """[R, C] -> [R/seg_rows, C]"""

    @staticmethod
    def forward(ctx, x, seg_rows):
        x = _c(_rows2d(x))
        R, C = x.shape
        ctx.R, ctx.seg_rows = R, seg_rows
        out = torch.empty((R // seg_rows, C), device=x.device, dtype=torch.float32)
        L().colsum(x.data_ptr(), R, C, seg_rows, out.data_ptr(), _ws(R, C, seg_rows, 1, x.device).data_ptr(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return BcastSeg.apply(g, ctx.R, ctx.seg_rows), None


class BcastSeg(Function):
    """v [nseg, C] -> [R, C] with out[r] = v[r // seg_rows]"""

    @staticmethod
    def forward(ctx, v, R, seg_rows):
        v = _c(v)
        C = v.shape[-1]
        ctx.seg_rows = seg_rows
        out = torch.empty((R, C), device=v.device, dtype=torch.float32)
        L().bcast_segvec(v.data_ptr(), R, C, seg_rows, out.data_ptr(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return ColSum.apply(g, ctx.seg_rows), None, None


class AddSegVec(Function):
    """x [R, C] + v[r // seg_rows] (bias add; per-cloud bias of the folded global feature)."""

    @staticmethod
    def forward(ctx, x, v, seg_rows):
        x, v = _c(_rows2d(x)), _c(v)
        R, C = x.shape
        ctx.seg_rows, ctx.vshape = seg_rows, v.shape
        out = torch.empty_like(x)
        L().add_segvec(x.data_ptr(), v.data_ptr(), R, C, seg_rows, out.data_ptr(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        dv = None
        if ctx.needs_input_grad[1] and not _INPUT_GRAD_ONLY:
            dv = ColSum.apply(g, ctx.seg_rows).view(ctx.vshape)
        return g, dv, None


# =========================================================================================
# elementwise
# =========================================================================================
class Axpby(Function):
    """a*x + b*y (y may be None)."""

    @staticmethod
    def forward(ctx, x, y, a, b):
        x = _c(x)
        ctx.a, ctx.b = a, b
        out = torch.empty_like(x)
        if y is not None:
            y = _c(y)
            if y.shape != x.shape:
                raise RuntimeError("axpby: shape mismatch")
        L().axpby(a, x.data_ptr(), b, y.data_ptr() if y is not None else None, out.data_ptr(), x.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        gx = Axpby.apply(g, None, ctx.a, 0.0) if ctx.needs_input_grad[0] else None
        gy = Axpby.apply(g, None, ctx.b, 0.0) if ctx.needs_input_grad[1] else None
        return gx, gy, None, None


def add(x, y):
    return Axpby.apply(x, y, 1.0, 1.0)


def sub(x, y):
    return Axpby.apply(x, y, 1.0, -1.0)


def scale(x, a):
    return Axpby.apply(x, None, float(a), 0.0)


class Mul(Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _c(x), _c(y)
        ctx.save_for_backward(x, y)
        out = torch.empty_like(x)
        L().mul(x.data_ptr(), y.data_ptr(), out.data_ptr(), x.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        return (Mul.apply(g, y) if ctx.needs_input_grad[0] else None,
                Mul.apply(g, x) if ctx.needs_input_grad[1] else None)


class LRelu(Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = _c(x)
        ctx.slope = slope
        ctx.save_for_backward(x)
        out = torch.empty_like(x)
        L().lrelu(x.data_ptr(), slope, out.data_ptr(), x.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return LReluBwd.apply(g, x, ctx.slope), None


class LReluBwd(Function):
    """g * (x > 0 ? 1 : slope): linear in g, piecewise constant in x."""

    @staticmethod
    def forward(ctx, g, x, slope):
        g = _c(g)
        ctx.slope = slope
        ctx.save_for_backward(x)
        out = torch.empty_like(g)
        L().lrelu_bwd(g.data_ptr(), x.data_ptr(), slope, out.data_ptr(), g.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, gg):
        (x,) = ctx.saved_tensors
        return LReluBwd.apply(gg, x, ctx.slope), None, None


class Tanh(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        out = torch.empty_like(x)
        L().tanh(x.data_ptr(), out.data_ptr(), x.numel(), _stream())
        ctx.save_for_backward(out)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _c(g)
        dx = torch.empty_like(g)
        L().tanh_bwd(g.data_ptr(), y.data_ptr(), dx.data_ptr(), g.numel(), _stream())
        return dx


# =========================================================================================
# layout
# =========================================================================================
class BcnToRows(Function):
    """[B, C, N] (any strides) -> rows [B*N, C]"""

    @staticmethod
    def forward(ctx, x):
        _chk(x)
        B, C, N = x.shape
        ctx.shape = (B, C, N)
        out = torch.empty((B * N, C), device=x.device, dtype=torch.float32)
        L().bcn_to_rows(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), B, C, N, out.data_ptr(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return RowsToBcn.apply(g, *ctx.shape)


class RowsToBcn(Function):
    """rows [B*N, C] -> contiguous [B, C, N]"""

    @staticmethod
    def forward(ctx, rows, B, C, N):
        rows = _c(rows)
        out = torch.empty((B, C, N), device=rows.device, dtype=torch.float32)
        L().rows_to_bcn(rows.data_ptr(), B, C, N, out.data_ptr(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return BcnToRows.apply(g), None, None, None


class ConcatCols(Function):
    """[a | b] along channels; b may be one row per segment broadcast over seg_rows rows
    (the tiled latent of model.py:128-131) when b_rows_per_seg == 1."""

    @staticmethod
    def forward(ctx, a, b, seg_rows, b_broadcast):
        a, b = _c(_rows2d(a)), _c(_rows2d(b))
        R, Ca = a.shape
        Cb = b.shape[1]
        ctx.dims = (R, Ca, Cb, seg_rows, b_broadcast)
        out = torch.empty((R, Ca + Cb), device=a.device, dtype=torch.float32)
        if b_broadcast:
            L().concat_cols(a.data_ptr(), Ca, Ca, b.data_ptr(), 0, Cb, seg_rows, Cb, R, out.data_ptr(), _stream())
        else:
            L().concat_cols(a.data_ptr(), Ca, Ca, b.data_ptr(), Cb, seg_rows * Cb, seg_rows, Cb, R, out.data_ptr(),
                            _stream())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        R, Ca, Cb, seg_rows, bb = ctx.dims
        g = _c(g)
        ga = torch.empty((R, Ca), device=g.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        gb = torch.empty((R, Cb), device=g.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        if ga is not None or gb is not None:
            L().split_cols_add(g.data_ptr(), R, Ca, Cb, ga.data_ptr() if ga is not None else None,
                               gb.data_ptr() if gb is not None else None, _stream())
        if gb is not None and bb:
            gb = ColSum.apply(gb, seg_rows)
        return ga, gb, None, None


class PermuteOCK(Function):
    """Conv2d(F, F, [1,k]) weight [O, C, 1, k] -> matrix [O, k*C] (and back in backward)."""

    @staticmethod
    def forward(ctx, w):
        w = _c(w)
        O, Cc, _, k = w.shape
        ctx.dims = (O, Cc, k)
        out = torch.empty((O, k * Cc), device=w.device, dtype=torch.float32)
        L().permute_ock_to_okc(w.data_ptr(), O, Cc, k, out.data_ptr(), _stream())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        O, Cc, k = ctx.dims
        g = _c(g)
        out = torch.empty((O, Cc, 1, k), device=g.device, dtype=torch.float32)
        L().permute_okc_to_ock(g.data_ptr(), O, Cc, k, out.data_ptr(), _stream())
        return out


# =========================================================================================
# normalisation
# =========================================================================================
def bn_running(bn):
    """(running_mean, running_var, num_batches_tracked, momentum) when a train-mode forward must advance them."""
    if bn.training and bn.track_running_stats:
        if bn.momentum is None:
            raise NotImplementedError("spgan_b200: BatchNorm(momentum=None) (cumulative average) is not supported; the "
                                      "reference never uses it (nn.BatchNorm defaults, momentum=0.1)")
        return (bn.running_mean, bn.running_var, bn.num_batches_tracked, float(bn.momentum))
    return None


def col_stats(x, seg_rows, eps, run=None):
    """(mean, rstd, biased var) per segment and column; with `run` (bn_running) the BatchNorm running statistics are
    advanced by the same launch pair (one segment only)."""
    x = _c(_rows2d(x))
    R, C = x.shape
    nseg = R // seg_rows
    mean = torch.empty((nseg, C), device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    var = torch.empty_like(mean)
    ws = _ws(R, C, seg_rows, 2, x.device)
    if run is not None and seg_rows == R:
        rm, rv, nbt, mom = run
        L().colstats_bn(x.data_ptr(), R, C, eps, mean.data_ptr(), rstd.data_ptr(), var.data_ptr(), mom, rm.data_ptr(),
                        rv.data_ptr(), nbt.data_ptr(), ws.data_ptr(), _stream())
    else:
        L().colstats(x.data_ptr(), R, C, seg_rows, eps, mean.data_ptr(), rstd.data_ptr(), var.data_ptr(), ws.data_ptr(),
                     _stream())
        if run is not None:
            rm, rv, nbt, mom = run
            L().bn_update_running(mean.data_ptr(), var.data_ptr(), mean.numel(), R, mom, rm.data_ptr(), rv.data_ptr(),
                                  nbt.data_ptr(), _stream())
    return mean, rstd, var


def _direct_ok(*params):
    """Parameter gradients may be accumulated in place by the producing kernel (see DIRECT_WEIGHT_GRAD)."""
    if not DIRECT_WEIGHT_GRAD or torch.is_grad_enabled():
        return False
    for p in params:
        if p is None or not p.is_leaf or not p.requires_grad or p.grad is None or not p.grad.is_contiguous() \
                or not p.grad.is_cuda or p.grad.dtype != torch.float32:
            return False
    return True


def _norm_bwd(g, x, slope, seg_rows, mean, rstd, gamma, beta, acc=None, addend=None):
    """(dx, sum g', sum g' xhat) of y = LeakyReLU_slope(norm(x) * gamma + beta); the activation mask is
    recomputed from x inside the kernels (slope 1 = no activation).  acc = (gamma_param, beta_param): their .grad
    receive dgamma / dbeta in place from the reduction's finalize pass (one segment only)."""
    R, C = x.shape
    nseg = R // seg_rows
    sg = torch.empty((nseg, C), device=x.device, dtype=torch.float32)
    sgx = torch.empty_like(sg)
    gp = gamma.data_ptr() if gamma is not None else None
    bp = beta.data_ptr() if beta is not None else None
    ws = _ws(R, C, seg_rows, 2, x.device)
    if acc is not None and seg_rows == R:
        L().norm_bwd_reduce_acc(g.data_ptr(), x.data_ptr(), slope, R, C, mean.data_ptr(), rstd.data_ptr(), gp, bp,
                                sg.data_ptr(), sgx.data_ptr(), acc[1].grad.data_ptr(), acc[0].grad.data_ptr(),
                                ws.data_ptr(), _stream())
    else:
        L().norm_bwd_reduce(g.data_ptr(), x.data_ptr(), slope, R, C, seg_rows, mean.data_ptr(), rstd.data_ptr(), gp, bp,
                            sg.data_ptr(), sgx.data_ptr(), ws.data_ptr(), _stream())
    dx = torch.empty_like(x)
    if addend is not None and seg_rows == R and C % 4 == 0 and addend.is_contiguous():
        # dx = BatchNorm backward + a second gradient term of x handed over by the double-backward node
        L().norm_bwd_apply_add(g.data_ptr(), x.data_ptr(), slope, R, C, mean.data_ptr(), rstd.data_ptr(), gp, bp,
                               sg.data_ptr(), sgx.data_ptr(), addend.data_ptr(), dx.data_ptr(), _stream())
        return dx, sg, sgx
    L().norm_bwd_apply(g.data_ptr(), x.data_ptr(), slope, R, C, seg_rows, mean.data_ptr(), rstd.data_ptr(), gp, bp,
                       sg.data_ptr(), sgx.data_ptr(), dx.data_ptr(), _stream())
    if addend is not None:
        dx = add(dx, addend)
    return dx, sg, sgx


def _bn_bwd_direct(ctx, gy, x, gamma, beta, mean, rstd, slope, addend=None):
    """Shared first-order backward of the train-mode BatchNorm Functions when no higher-order graph is being built:
    dgamma / dbeta go straight into the parameters' .grad where possible.  addend: a second gradient term of x (from
    the double-backward node of the same layer), added inside the apply kernel."""
    acc = (gamma, beta) if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2] and not _INPUT_GRAD_ONLY
                            and _direct_ok(gamma, beta)) else None
    dx, sg, sgx = _norm_bwd(_c(gy), x, slope, x.shape[0], mean, rstd, gamma, beta, acc, addend)
    if acc is not None or _INPUT_GRAD_ONLY:
        return dx, None, None
    return dx, sgx.view(-1), sg.view(-1)


class BatchNormTrain(Function):
    """Train-mode batch norm over the rows (BatchNorm1d/2d on [B,C,N(,k)]); twice differentiable.
    Returns (y, batch_mean, biased_batch_var); the last two feed the running-stat update."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, run=None):
        x = _c(_rows2d(x))
        R, C = x.shape
        mean, rstd, var = col_stats(x, R, eps, run)
        y = torch.empty_like(x)
        L().norm_apply(x.data_ptr(), R, C, R, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                       1.0, y.data_ptr(), _stream())
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.mark_non_differentiable(mean, var)
        ctx.set_materialize_grads(False)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, _gm, _gv):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        if gy is None:
            return None, None, None, None, None
        if not torch.is_grad_enabled():            # plain first-order pass: no graph of the backward is needed
            return _bn_bwd_direct(ctx, gy, x, gamma, beta, mean, rstd, 1.0) + (None, None)
        dx, dgamma, dbeta = BatchNormTrainBwd.apply(gy, x, gamma, mean, rstd)
        if _INPUT_GRAD_ONLY:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, None, None


class BatchNormTrainBwd(Function):
    """(g, x, gamma) -> (dx, dgamma, dbeta) of train-mode BN; its backward is the closed-form
    double backward (w.r.t. g, x, gamma) for a cotangent on dx."""

    @staticmethod
    def forward(ctx, g, x, gamma, mean, rstd):
        g = _c(g)
        ctx.set_materialize_grads(False)
        dx, sg, sgx = _norm_bwd(g, x, 1.0, x.shape[0], mean, rstd, gamma, None)
        ctx.save_for_backward(g, x, gamma, mean, rstd)
        return dx, sgx.view(-1), sg.view(-1)

    @staticmethod
    @once_differentiable
    def backward(ctx, u, a, c):
        g, x, gamma, mean, rstd = ctx.saved_tensors
        if a is not None or c is not None:
            raise NotImplementedError("double backward through BatchNorm parameter gradients is not needed by "
                                      "the WGAN-GP step (only_inputs=True) and is not implemented")
        if u is None:
            return None, None, None, None, None
        u = _c(u)
        R, C = x.shape
        sums = torch.empty((5, C), device=x.device, dtype=torch.float32)
        L().bn_dbl_bwd_reduce(g.data_ptr(), u.data_ptr(), x.data_ptr(), R, C, mean.data_ptr(), sums.data_ptr(),
                              _ws(R, C, R, 5, x.device).data_ptr(), _stream())
        gg = torch.empty_like(x)
        gx = torch.empty_like(x)
        ggamma = torch.empty((C,), device=x.device, dtype=torch.float32)
        L().bn_dbl_bwd_apply(g.data_ptr(), u.data_ptr(), x.data_ptr(), R, C, mean.data_ptr(), rstd.data_ptr(),
                             gamma.data_ptr(), sums.data_ptr(), gg.data_ptr(), gx.data_ptr(), ggamma.data_ptr(),
                             _stream())
        return gg, gx, ggamma, None, None


class BatchNormActTrain2(Function):
    """Train-mode BN + LeakyReLU in one pass, TWICE differentiable (the critic under the gradient penalty,
    gradient_penalty.py:28-33): same forward as BatchNormActTrain; the backward is itself a Function whose backward
    is the closed-form double backward with the activation mask folded in."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope, run=None):
        x = _c(_rows2d(x))
        R, C = x.shape
        mean, rstd, var = col_stats(x, R, eps, run)
        y = torch.empty_like(x)
        L().norm_apply(x.data_ptr(), R, C, R, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                       slope, y.data_ptr(), _stream())
        ctx.slope = slope
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.mark_non_differentiable(mean, var)
        ctx.set_materialize_grads(False)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, _gm, _gv):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        link = getattr(ctx, "link", None)
        if not torch.is_grad_enabled():            # the final (first-order) pass over the penalty's graph
            # The double-backward node of this layer ran earlier in this pass (it was created later) and left ITS
            # gradient term for x here instead of returning it: the two terms are summed inside the apply kernel, not
            # by a separate accumulation pass of autograd over the [P, C] tensor.
            extra = link.pop("gx", None) if link else None
            if gy is None:
                return extra, None, None, None, None, None
            return _bn_bwd_direct(ctx, gy, x, gamma, beta, mean, rstd, ctx.slope, extra) + (None, None, None)
        if gy is None:
            return None, None, None, None, None, None
        link = ctx.link = {"fwd": weakref.ref(ctx)}      # (weak: ctx -> link -> ctx would keep the saved activations alive)
        dx, dgamma, dbeta = BatchNormActTrainBwd2.apply(gy, x, gamma, beta, mean, rstd, ctx.slope, link)
        if _INPUT_GRAD_ONLY:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, None, None, None


class BatchNormActTrainBwd2(Function):
    """(g, x, gamma, beta) -> (dx, dgamma, dbeta) of fused BN + LeakyReLU; backward = double backward for a
    cotangent on dx (the mask is piecewise constant in x: it multiplies g on the way in and gg on the way out)."""

    @staticmethod
    def forward(ctx, g, x, gamma, beta, mean, rstd, slope, link=None):
        g = _c(g)
        ctx.set_materialize_grads(False)
        dx, sg, sgx = _norm_bwd(g, x, slope, x.shape[0], mean, rstd, gamma, beta)
        ctx.slope, ctx.link = slope, link
        ctx.hand_over = link is not None and ctx.needs_input_grad[1]
        ctx.save_for_backward(g, x, gamma, beta, mean, rstd)
        return dx, sgx.view(-1), sg.view(-1)

    @staticmethod
    @once_differentiable
    def backward(ctx, u, a, c):
        g, x, gamma, beta, mean, rstd = ctx.saved_tensors
        if a is not None or c is not None:
            raise NotImplementedError("double backward through BatchNorm parameter gradients is not needed by "
                                      "the WGAN-GP step (only_inputs=True) and is not implemented")
        if u is None:
            return None, None, None, None, None, None, None, None
        u = _c(u)
        R, C = x.shape
        sums = torch.empty((5, C), device=x.device, dtype=torch.float32)
        L().bn_act_dbl_bwd_reduce(g.data_ptr(), u.data_ptr(), x.data_ptr(), ctx.slope, R, C, mean.data_ptr(),
                                  rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), sums.data_ptr(),
                                  _ws(R, C, R, 5, x.device).data_ptr(), _stream())
        gg = torch.empty_like(x)
        gx = torch.empty_like(x)
        ggamma = torch.empty((C,), device=x.device, dtype=torch.float32)
        L().bn_act_dbl_bwd_apply(g.data_ptr(), u.data_ptr(), x.data_ptr(), ctx.slope, R, C, mean.data_ptr(),
                                 rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), sums.data_ptr(), gg.data_ptr(),
                                 gx.data_ptr(), ggamma.data_ptr(), _stream())
        fwd = ctx.link["fwd"]() if ctx.hand_over else None
        if fwd is not None and FUSE_GP_ACCUMULATE and _node_will_run(fwd):
            global GP_HANDOVERS
            GP_HANDOVERS += 1
            ctx.link["gx"] = gx                    # picked up by BatchNormActTrain2.backward of the same layer (see there)
            gx = None
        return gg, gx, ggamma, None, None, None, None, None


FUSE_GP_ACCUMULATE = _os.environ.get("SPGAN_FUSE_GP_ACCUMULATE", "1") != "0"
GP_HANDOVERS = 0          # diagnostics: gradient terms handed from a double-backward node to its layer's backward


def _node_will_run(node):
    """Will the autograd engine execute `node` in the backward pass that is running now?  (The hand-over of a gradient
    term to a node is only sound if that node's backward really follows; otherwise the term is returned the normal way.)"""
    if node is None:
        return False
    try:
        return bool(torch._C._will_engine_execute_node(node))
    except Exception:                                        # noqa: BLE001 -- private API: absent or refusing = no hand-over
        return False


class BatchNormActTrain(Function):
    """Fused train-mode BN + LeakyReLU (slope 0 = ReLU); first-order only (generator path)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope, run=None):
        x = _c(_rows2d(x))
        R, C = x.shape
        mean, rstd, var = col_stats(x, R, eps, run)
        y = torch.empty_like(x)
        L().norm_apply(x.data_ptr(), R, C, R, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                       slope, y.data_ptr(), _stream())
        ctx.slope = slope
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.mark_non_differentiable(mean, var)
        ctx.set_materialize_grads(False)
        return y, mean, var

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, _gm, _gv):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        if gy is None:
            return None, None, None, None, None, None
        return _bn_bwd_direct(ctx, gy, x, gamma, beta, mean, rstd, ctx.slope) + (None, None, None)


class BatchNormActSegMaxTrain(Function):
    """Fused train-mode BN + LeakyReLU + max over the points of each cloud (the critic's fc2 -> max pool,
    Discriminator.py:77-81,104): one read of x in the forward, the normalised tensor is never materialised;
    the backward is sparse in the incoming gradient.  First-order only.
    Returns (pooled [nseg, C], batch_mean, biased_batch_var)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope, seg_rows):
        x = _c(_rows2d(x))
        R, C = x.shape
        nseg = R // seg_rows
        dev = x.device
        mean = torch.empty((1, C), device=dev, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        var = torch.empty_like(mean)
        pooled = torch.empty((nseg, C), device=dev, dtype=torch.float32)
        arg = torch.empty((nseg, C), device=dev, dtype=torch.int32)
        ws = torch.empty((L().bn_pool_workspace(R, C, seg_rows) + 15) // 16 * 4, device=dev, dtype=torch.float32)
        L().bn_pool_fwd(x.data_ptr(), R, C, seg_rows, gamma.data_ptr(), beta.data_ptr(), eps, slope, mean.data_ptr(),
                        rstd.data_ptr(), var.data_ptr(), pooled.data_ptr(), arg.data_ptr(), ws.data_ptr(), _stream())
        ctx.slope, ctx.seg_rows = slope, seg_rows
        ctx.save_for_backward(x, gamma, beta, mean, rstd, arg)
        ctx.mark_non_differentiable(mean, var)
        ctx.set_materialize_grads(False)
        return pooled, mean, var

    @staticmethod
    @once_differentiable
    def backward(ctx, gp, _gm, _gv):
        x, gamma, beta, mean, rstd, arg = ctx.saved_tensors
        if gp is None:
            return None, None, None, None, None, None
        gp = _c(gp)
        R, C = x.shape
        dev = x.device
        gprime = torch.empty_like(gp)
        sg = torch.empty((C,), device=dev, dtype=torch.float32)
        sgx = torch.empty_like(sg)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        L().bn_pool_bwd(gp.data_ptr(), x.data_ptr(), arg.data_ptr(), R, C, ctx.seg_rows, mean.data_ptr(), rstd.data_ptr(),
                        gamma.data_ptr(), beta.data_ptr(), ctx.slope, gprime.data_ptr(), sg.data_ptr(), sgx.data_ptr(),
                        dx.data_ptr() if dx is not None else None, _stream())
        return dx, sgx, sg, None, None, None


FUSE_WGRAD_PROLOGUE = _os.environ.get("SPGAN_FUSE_WGRAD_PROLOGUE", "1") != "0"


def wgrad_bn_act(gz, x_pre, scale, shift, slope, mean, rstd, gamma, beta, W, direct):
    """dW of z = lrelu(bn(x_pre)) @ W^T: gz^T @ lrelu(x_pre * scale + shift).  The activated input is re-formed inside
    the weight-gradient kernel's operand converter (spgan_gemm_wgrad_fused); outside that kernel's envelope it is
    recomputed by a norm_apply pass first.  direct: accumulate into W.grad (returns None), else returns dW."""
    R, K = x_pre.shape
    Cout = gz.shape[1]
    out = _wmat(W.grad, None) if direct else torch.empty((Cout, K), device=gz.device, dtype=torch.float32)
    if (FUSE_WGRAD_PROLOGUE and GEMM_ENGINE == 3 and R >= 4096 and 16 <= Cout <= 65536 and K >= 16 and 0.0 < slope <= 1.0
            and gz.stride(0) % 4 == 0 and x_pre.stride(0) % 4 == 0 and gz.data_ptr() % 16 == 0 and x_pre.data_ptr() % 16 == 0
            and gz.stride(1) == 1 and x_pre.stride(1) == 1):
        global LAST_TC_WORKSPACE
        ws_bytes = L().gemm_wgrad_workspace(Cout, K, R)
        ws = torch.empty(ws_bytes // 4, device=gz.device, dtype=torch.float32)
        LAST_TC_WORKSPACE = ws
        L().gemm_wgrad_fused(Cout, K, R, gz.data_ptr(), gz.stride(0), x_pre.data_ptr(), x_pre.stride(0), scale.data_ptr(),
                             shift.data_ptr(), float(slope), out.data_ptr(), out.stride(0), int(direct), ws.data_ptr(),
                             ws_bytes, _stream())
    else:
        a = torch.empty_like(x_pre)                       # recompute the activated input for the weight gradient
        if mean is None:                                  # activation-only prologue (ActLinear)
            L().lrelu(x_pre.data_ptr(), slope, a.data_ptr(), x_pre.numel(), _stream())
        else:
            L().norm_apply(x_pre.data_ptr(), R, K, R, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                           slope, a.data_ptr(), _stream())
        gemm_raw(gz, a, None, True, False, out=out, accumulate=direct)
    return None if direct else out.reshape(W.shape)


class BnActLinearTrain(Function):
    """z = LeakyReLU(BatchNorm_train(x_pre)) @ W^T + b in ONE pass over x_pre (spgan_gemm_fused): the normalised,
    activated tensor never reaches HBM -- it is formed in the GEMM's A-operand converter -- and, with `next_bn`, the
    batch statistics of z (the next BatchNorm's input) come out of the GEMM's epilogue, so neither the statistics pass
    nor the apply pass of conv -> BN -> LeakyReLU -> conv chains (Discriminator.py:55-81, Generator.py:56-62) exists.
    `stats` = (mean, rstd, scale, shift) of x_pre (from a previous BnActLinearTrain or bn_train_stats).
    Returns (z, mean_z, rstd_z, scale_z, shift_z) (the last four None-like empties without next_bn).
    Backward (first order): the weight gradient needs the activated input, which is recomputed from x_pre."""

    @staticmethod
    def forward(ctx, x_pre, mean, rstd, scale, shift, gamma, beta, W, bias, slope, next_bn, zero_bias_grad):
        x_pre = _c(_rows2d(x_pre))
        R = x_pre.shape[0]
        Wm = _wmat(W, None)
        want = next_bn is not None
        res = gemm_fused_raw(x_pre, Wm, bias, tb=True, a_scale=scale, a_shift=shift, a_slope=slope, want_stats=want,
                             wcache=True)
        if res is None:
            raise RuntimeError("BnActLinearTrain: shape outside spgan_gemm_fused's envelope (caller must check fused_linear_ok)")
        ctx.slope, ctx.zero_bias_grad, ctx.has_bias = slope, zero_bias_grad, bias is not None
        ctx.save_for_backward(x_pre, mean, rstd, gamma, beta, W, scale, shift)
        ctx.set_materialize_grads(False)
        if not want:
            return res
        z, cs, cq = res
        C2 = z.shape[1]
        dev = z.device
        m2 = torch.empty((1, C2), device=dev, dtype=torch.float32)
        r2, v2, sc2, sh2 = torch.empty_like(m2), torch.empty_like(m2), torch.empty_like(m2), torch.empty_like(m2)
        run = bn_running(next_bn)
        L().bn_finalize(cs.data_ptr(), cq.data_ptr(), cs.shape[0], C2, R, float(next_bn.eps), next_bn.weight.data_ptr(),
                        next_bn.bias.data_ptr(), m2.data_ptr(), r2.data_ptr(), v2.data_ptr(), sc2.data_ptr(), sh2.data_ptr(),
                        run[3] if run else 0.0, run[0].data_ptr() if run else None, run[1].data_ptr() if run else None,
                        run[2].data_ptr() if run else None, _stream())
        ctx.mark_non_differentiable(m2, r2, v2, sc2, sh2)
        return z, m2, r2, v2, sc2, sh2

    @staticmethod
    @once_differentiable
    def backward(ctx, gz, *_unused):
        x_pre, mean, rstd, gamma, beta, W, scale, shift = ctx.saved_tensors
        if gz is None:
            return (None,) * 12
        gz = _c(gz)
        R, K = x_pre.shape
        Wm = _wmat(W, None)
        dW = db = dgamma = dbeta = dx = None
        params_too = not _INPUT_GRAD_ONLY
        if ctx.needs_input_grad[7] and params_too:
            dW = wgrad_bn_act(gz, x_pre, scale, shift, ctx.slope, mean, rstd, gamma, beta, W, _direct_ok(W))
        if ctx.has_bias and ctx.needs_input_grad[8] and params_too and not ctx.zero_bias_grad:
            db = ColSum.apply(gz, R).view(-1)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[5] or ctx.needs_input_grad[6]:
            ga = gemm_raw(gz, Wm, None, False, False, wcache=True)         # d(activated input)
            acc = (gamma, beta) if (ctx.needs_input_grad[5] and ctx.needs_input_grad[6] and params_too
                                    and _direct_ok(gamma, beta)) else None
            dx, sg, sgx = _norm_bwd(ga, x_pre, ctx.slope, R, mean, rstd, gamma, beta, acc)
            if acc is None and params_too:
                dgamma, dbeta = sgx.view(-1), sg.view(-1)
        return dx, None, None, None, None, dgamma, dbeta, dW, db, None, None, None


_UNIT_TABLES = {}


def _unit_tables(K, device):
    """(ones[K], zeros[K]): the identity scale / shift of an activation-only operand prologue."""
    key = (K, str(device))
    t = _UNIT_TABLES.get(key)
    if t is None:
        t = _UNIT_TABLES[key] = (full((K,), 1.0, device), full((K,), 0.0, device))
    return t


FUSE_ACT_LINEAR = _os.environ.get("SPGAN_FUSE_ACT_LINEAR", "1") != "0"


class ActLinear(Function):
    """z = LeakyReLU_slope(x_pre) @ W^T + b with the activation inside the GEMM's operand converter (spgan_gemm_fused,
    identity scale / shift): the activated tensor of a conv -> LeakyReLU -> conv pair (Generator.py:107-110,128-133)
    is never written.  Backward: d x_pre = LeakyReLU'(x_pre) * (g @ W); the weight gradient re-forms the activated
    operand inside its own converter.  `cols` = (c0, c1) restricts W to an input-channel block.  First order only."""

    @staticmethod
    def forward(ctx, x_pre, W, bias, slope, cols):
        x_pre = _c(_rows2d(x_pre))
        K = x_pre.shape[1]
        ones, zeros = _unit_tables(K, x_pre.device)
        res = gemm_fused_raw(x_pre, _wmat(W, cols), bias, tb=True, a_scale=ones, a_shift=zeros, a_slope=slope, wcache=True)
        if res is None:
            raise RuntimeError("ActLinear: shape outside spgan_gemm_fused's envelope (caller must check act_linear_ok)")
        ctx.slope, ctx.cols, ctx.has_bias = slope, cols, bias is not None
        ctx.save_for_backward(x_pre, W)
        ctx.set_materialize_grads(False)
        return res

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        if g is None:
            return None, None, None, None, None
        x_pre, W = ctx.saved_tensors
        g = _c(g)
        R, K = x_pre.shape
        Wm = _wmat(W, ctx.cols)
        dx = dW = db = None
        params_too = not _INPUT_GRAD_ONLY
        if ctx.needs_input_grad[1] and params_too:
            ones, zeros = _unit_tables(K, x_pre.device)
            direct = _direct_ok(W)
            if ctx.cols is None:
                dW = wgrad_bn_act(g, x_pre, ones, zeros, ctx.slope, None, None, None, None, W, direct)
            else:
                a = LRelu.apply(x_pre, ctx.slope)
                if direct:
                    gemm_raw(g, a, None, True, False, out=_wmat(W.grad, ctx.cols), accumulate=True)
                else:
                    blk = gemm_raw(g, a, None, True, False)
                    dW = full((Wm.shape[0], _wmat(W, None).shape[1]), 0.0, g.device)
                    dW[:, ctx.cols[0]:ctx.cols[1]] = blk
                    dW = dW.reshape(W.shape)
        if ctx.has_bias and ctx.needs_input_grad[2] and params_too:
            db = ColSum.apply(g, R).view(-1)
        if ctx.needs_input_grad[0]:
            ga = gemm_raw(g, Wm, None, False, False, wcache=True)
            dx = torch.empty_like(ga)
            L().lrelu_bwd(ga.data_ptr(), x_pre.data_ptr(), ctx.slope, dx.data_ptr(), ga.numel(), _stream())
        return dx, dW, db, None, None


def act_linear_ok(R, weight, cols=None):
    """Can LeakyReLU(x) @ weight^T run as one spgan_gemm_fused launch (first-order graph, kernel envelope)?"""
    if not FUSE_ACT_LINEAR or _TWICE_DIFFERENTIABLE or GEMM_ENGINE != 3 or not weight.is_cuda:
        return False
    K = _wmat(weight, cols).shape[1]
    return K % 4 == 0 and L().gemm_fused_workspace(R, weight.shape[0], K, 256, K) != 0


def act_linear(x_pre, slope, weight, bias=None, cols=None):
    """LeakyReLU_slope(x_pre) @ weight^T + bias: fused (ActLinear) where the kernel allows, else the two-op chain."""
    x2 = _rows2d(x_pre)
    if act_linear_ok(x2.shape[0], weight, cols) and x2.is_contiguous():
        return ActLinear.apply(x2, weight, bias, slope, cols)
    return linear(LRelu.apply(x_pre, slope), weight, bias, cols=cols)


def fused_linear_ok(R, weight, bn_in, next_bn=None):
    """Can conv(LeakyReLU(BatchNorm(x))) for a contiguous x [R, K] run as one spgan_gemm_fused launch?  Train-mode
    statistics with running buffers, first-order graph, the kernel's shape envelope."""
    if not FUSE_BN_GEMM or _TWICE_DIFFERENTIABLE or GEMM_ENGINE != 3 or not weight.is_cuda:
        return False
    if not (bn_in.training and bn_in.track_running_stats and bn_in.momentum is not None):
        return False
    if next_bn is not None and not (next_bn.training and next_bn.track_running_stats and next_bn.momentum is not None
                                    and weight.shape[0] <= 256):
        return False
    K = weight[0].numel()
    return L().gemm_fused_workspace(R, weight.shape[0], K, 256, K) != 0          # (256: any 16-byte aligned address)


def bn_train_stats(x_pre, bn):
    """(mean, rstd, scale, shift) of a pre-normalisation tensor whose producer is not a fused GEMM (a K = 3 conv, the
    edge gather): the statistics pass (which also advances the running buffers) plus the 1-launch table kernel."""
    x_pre = _c(_rows2d(x_pre))
    R, C = x_pre.shape
    mean, rstd, _ = col_stats(x_pre.detach(), R, bn.eps, bn_running(bn))
    scale, shift = torch.empty_like(mean), torch.empty_like(mean)
    L().bn_tables(mean.data_ptr(), rstd.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(), C, scale.data_ptr(),
                  shift.data_ptr(), _stream())
    return mean, rstd, scale, shift


def bn_act_linear(x_pre, stats, bn_in, slope, weight, bias, next_bn=None, zero_bias_grad=False):
    """conv(LeakyReLU_slope(bn_in(x_pre))) fused (see BnActLinearTrain); stats = (mean, rstd, scale, shift) of x_pre.
    -> z, or (z, (mean, rstd, scale, shift), var) of z when next_bn is given."""
    mean, rstd, scale, shift = stats
    out = BnActLinearTrain.apply(x_pre, mean, rstd, scale, shift, bn_in.weight, bn_in.bias, weight, bias, slope, next_bn,
                                 zero_bias_grad)
    if next_bn is None:
        return out
    z, m2, r2, v2, sc2, sh2 = out
    return z, (m2, r2, sc2, sh2), v2


FUSE_BN_GEMM = _os.environ.get("SPGAN_FUSE_BN_GEMM", "1") != "0"
FUSE_BN_POOL = _os.environ.get("SPGAN_FUSE_BN_POOL", "1") != "0"


def batch_norm_act_segmax(y, bn, slope, seg_rows):
    """max over each cloud of LeakyReLU(BatchNorm(y)): the fused kernel on the first-order train-mode path,
    the generic chain (batch_norm_act + SegMax) in eval mode and under the gradient penalty's double backward."""
    R, C = y.shape
    if (FUSE_BN_POOL and bn.training and bn.track_running_stats and not _TWICE_DIFFERENTIABLE and C % 4 == 0
            and 0.0 < slope <= 1.0):
        pooled, mean, var = BatchNormActSegMaxTrain.apply(y, bn.weight, bn.bias, bn.eps, slope, seg_rows)
        bn_update_running(mean, var, R, bn)
        return pooled
    return SegMax.apply(batch_norm_act(y, bn, slope), seg_rows)


class NormAffineEval(Function):
    """Eval-mode BN (+activation): y = act((x - rm) * rstd * gamma + beta) with frozen statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, rm, rstd, slope):
        x = _c(_rows2d(x))
        R, C = x.shape
        y = torch.empty_like(x)
        L().norm_apply(x.data_ptr(), R, C, R, rm.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                       slope, y.data_ptr(), _stream())
        ctx.slope = slope
        ctx.save_for_backward(x, y, gamma, beta, rm, rstd)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, y, gamma, beta, rm, rstd = ctx.saved_tensors
        gy = _c(gy)
        R, C = x.shape
        # frozen statistics: dx = g' * gamma * rstd, dgamma = sum g' xhat, dbeta = sum g'
        sg = torch.empty((1, C), device=x.device, dtype=torch.float32)
        sgx = torch.empty_like(sg)
        L().norm_bwd_reduce(gy.data_ptr(), x.data_ptr(), ctx.slope, R, C, R, rm.data_ptr(), rstd.data_ptr(),
                            gamma.data_ptr(), beta.data_ptr(), sg.data_ptr(), sgx.data_ptr(),
                            _ws(R, C, R, 2, x.device).data_ptr(), _stream())
        gmask = gy
        if ctx.slope != 1.0:
            gmask = torch.empty_like(gy)
            L().lrelu_bwd(gy.data_ptr(), y.data_ptr(), ctx.slope, gmask.data_ptr(), gy.numel(), _stream())
        coef = torch.empty((C,), device=x.device, dtype=torch.float32)
        L().mul(gamma.data_ptr(), rstd.data_ptr(), coef.data_ptr(), C, _stream())
        dx = torch.empty_like(x)
        L().mul_segvec(gmask.data_ptr(), coef.data_ptr(), R, C, R, dx.data_ptr(), _stream())
        return dx, sgx.view(-1), sg.view(-1), None, None, None


class AdaIN(Function):
    """out = s[:, :C] * InstanceNorm(x) + s[:, C:], statistics per cloud (segment) and channel
    (Generator.py:38-45); first-order only."""

    @staticmethod
    def forward(ctx, x, s, seg_rows, eps):
        x, s = _c(_rows2d(x)), _c(_rows2d(s))
        R, C = x.shape
        mean, rstd, _ = col_stats(x, seg_rows, eps)
        out = torch.empty_like(x)
        L().adain_apply(x.data_ptr(), s.data_ptr(), R, C, seg_rows, mean.data_ptr(), rstd.data_ptr(), out.data_ptr(),
                        _stream())
        ctx.seg_rows = seg_rows
        ctx.save_for_backward(x, s, mean, rstd)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, s, mean, rstd = ctx.saved_tensors
        g = _c(g)
        R, C = x.shape
        ds = torch.empty_like(s) if ctx.needs_input_grad[1] else None
        gxh = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        L().adain_bwd(g.data_ptr(), x.data_ptr(), s.data_ptr(), R, C, ctx.seg_rows, mean.data_ptr(), rstd.data_ptr(),
                      ds.data_ptr() if ds is not None else None, gxh.data_ptr() if gxh is not None else None, _stream())
        dx = None
        if gxh is not None:
            dx, _, _ = _norm_bwd(gxh, x, 1.0, ctx.seg_rows, mean, rstd, None, None)
        return dx, ds, None, None


def rsqrt_eps(v, eps):
    v = _c(v)
    out = torch.empty_like(v)
    L().rsqrt_eps(v.data_ptr(), float(eps), v.numel(), out.data_ptr(), _stream())
    return out


def row_l2_normalize(x, eps):
    """x / (||x||_2 + eps) over the last axis; forward only (the latent carries no gradient)."""
    if x.requires_grad:
        raise NotImplementedError("z_norm with a latent that requires grad is not supported")
    x = _c(x)
    C = x.shape[-1]
    out = torch.empty_like(x)
    L().row_l2_normalize(x.data_ptr(), x.numel() // C, C, float(eps), out.data_ptr(), _stream())
    return out


FUSE_BN_ACT_2ND = _os.environ.get("SPGAN_FUSE_BN_ACT_2ND", "1") != "0"


def batch_norm_act(y, bn, slope):
    """nn.BatchNorm{1,2}d (+ LeakyReLU(slope); slope 1 = none) on rows [R, C], honouring bn.training and
    updating the running statistics like the reference modules do in train mode."""
    if bn.training or not bn.track_running_stats:
        run = bn_running(bn)                # the statistics pass also advances the running buffers
        if _TWICE_DIFFERENTIABLE and FUSE_BN_ACT_2ND and slope != 1.0 and y.shape[1] % 4 == 0:
            z, mean, var = BatchNormActTrain2.apply(y, bn.weight, bn.bias, bn.eps, slope, run)
        elif _TWICE_DIFFERENTIABLE:
            z, mean, var = BatchNormTrain.apply(y, bn.weight, bn.bias, bn.eps, run)
            if slope != 1.0:
                z = LRelu.apply(z, slope)
        else:
            z, mean, var = BatchNormActTrain.apply(y, bn.weight, bn.bias, bn.eps, slope, run)
        return z
    if _TWICE_DIFFERENTIABLE:
        # GradientPenalty(D.eval(), ...): frozen statistics make the layer an affine map followed by the activation;
        # composed from the closed (twice differentiable) operator set: y = lrelu(x * (gamma rstd) + (beta - rm gamma rstd))
        y2 = _c(_rows2d(y))
        R, C = y2.shape
        sc = Mul.apply(bn.weight.view(1, C), rsqrt_eps(bn.running_var, bn.eps).view(1, C))
        sh = sub(bn.bias.view(1, C), Mul.apply(bn.running_mean.view(1, C), sc))
        z = AddSegVec.apply(Mul.apply(y2, BcastSeg.apply(sc, R, R)), sh, R)
        return LRelu.apply(z, slope) if slope != 1.0 else z
    return NormAffineEval.apply(y, bn.weight, bn.bias, bn.running_mean, rsqrt_eps(bn.running_var, bn.eps), slope)


def bn_update_running(mean, var, R, bn):
    """Side effect of a train-mode forward on nn.BatchNorm buffers (momentum, unbiased var, count)."""
    m = bn_running(bn)[3]
    L().bn_update_running(mean.data_ptr(), var.data_ptr(), mean.numel(), R, float(m), bn.running_mean.data_ptr(),
                          bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), _stream())


# =========================================================================================
# pooling over points / neighbours
# =========================================================================================
class SegMax(Function):
    """max over each segment of seg_rows rows -> ([nseg, C]); gradient to the first arg max."""

    @staticmethod
    def forward(ctx, x, seg_rows):
        x = _c(_rows2d(x))
        R, C = x.shape
        out = torch.empty((R // seg_rows, C), device=x.device, dtype=torch.float32)
        arg = torch.empty((R // seg_rows, C), device=x.device, dtype=torch.int32)
        L().segmax(x.data_ptr(), R, C, seg_rows, out.data_ptr(), arg.data_ptr(), _stream())
        ctx.dims = (R, seg_rows)
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        return SegMaxScatter.apply(g, arg, *ctx.dims), None


class SegMaxScatter(Function):
    @staticmethod
    def forward(ctx, g, arg, R, seg_rows):
        g = _c(g)
        C = g.shape[1]
        dx = torch.empty((R, C), device=g.device, dtype=torch.float32)
        L().segmax_scatter(g.data_ptr(), arg.data_ptr(), R, C, seg_rows, dx.data_ptr(), _stream())
        ctx.dims = (R, seg_rows)
        ctx.save_for_backward(arg)
        return dx

    @staticmethod
    def backward(ctx, gg):
        (arg,) = ctx.saved_tensors
        return SegMaxGather.apply(gg, arg, *ctx.dims), None, None, None


class SegMaxGather(Function):
    @staticmethod
    def forward(ctx, x, arg, R, seg_rows):
        x = _c(x)
        C = x.shape[1]
        out = torch.empty((R // seg_rows, C), device=x.device, dtype=torch.float32)
        L().segmax_gather(x.data_ptr(), arg.data_ptr(), R, C, seg_rows, out.data_ptr(), _stream())
        ctx.dims = (R, seg_rows)
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        return SegMaxScatter.apply(g, arg, *ctx.dims), None, None, None


class SoftmaxK(Function):
    """softmax over the k neighbours; x is [P*k, C] edge-major (F.softmax(w, -1), Generator.py:79)."""

    @staticmethod
    def forward(ctx, x, k):
        x = _c(_rows2d(x))
        E, C = x.shape
        y = torch.empty_like(x)
        L().softmax_k(x.data_ptr(), E // k, k, C, y.data_ptr(), _stream())
        ctx.k = k
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _c(g)
        E, C = y.shape
        dx = torch.empty_like(y)
        L().softmax_k_bwd(g.data_ptr(), y.data_ptr(), E // ctx.k, ctx.k, C, dx.data_ptr(), _stream())
        return dx, None


class SoftmaxMulK(Function):
    """y * softmax_k(x) in one pass (EdgeBlock's attention modulation, Generator.py:79,82); x, y are
    [P*k, C] edge-major.  First-order only."""

    @staticmethod
    def forward(ctx, x, y, k):
        x, y = _c(_rows2d(x)), _c(_rows2d(y))
        E, C = x.shape
        w = torch.empty_like(x)
        prod = torch.empty_like(x)
        L().softmax_mul_k(x.data_ptr(), y.data_ptr(), E // k, k, C, w.data_ptr(), prod.data_ptr(), _stream())
        ctx.k = k
        ctx.save_for_backward(y, w)
        return prod

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        y, w = ctx.saved_tensors
        g = _c(g)
        E, C = y.shape
        dx = torch.empty_like(y) if ctx.needs_input_grad[0] else None
        dy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        L().softmax_mul_k_bwd(g.data_ptr(), y.data_ptr(), w.data_ptr(), E // ctx.k, ctx.k, C,
                              dx.data_ptr() if dx is not None else None, dy.data_ptr() if dy is not None else None,
                              _stream())
        return dx, dy, None


class RowSoftmax(Function):
    """softmax over the last axis of [R, N] (the attention map of --attn, modules.py:552). First-order only."""

    @staticmethod
    def forward(ctx, x):
        x = _c(_rows2d(x))
        R, N = x.shape
        y = torch.empty_like(x)
        L().row_softmax(x.data_ptr(), R, N, y.data_ptr(), _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = _c(g)
        R, N = y.shape
        dx = torch.empty_like(y)
        L().row_softmax_bwd(g.data_ptr(), y.data_ptr(), R, N, dx.data_ptr(), _stream())
        return dx


class SegGemm(Function):
    """Per-cloud product (torch.bmm over row blocks): A [nseg*ra, ca], B [nseg*rb, cb] -> C [nseg*M, N] with
    C_s = A_s op(B_s), op = transpose when tb.  One spgan_gemm launch per cloud.  First-order only
    (used by the optional Attention block, modules.py:552-554)."""

    @staticmethod
    def forward(ctx, A, B, nseg, tb):
        A, B = _c(_rows2d(A)), _c(_rows2d(B))
        ra, rb = A.shape[0] // nseg, B.shape[0] // nseg
        M, N = ra, (rb if tb else B.shape[1])
        out = torch.empty((nseg * M, N), device=A.device, dtype=torch.float32)
        for s in range(nseg):
            gemm_raw(A[s * ra:(s + 1) * ra], B[s * rb:(s + 1) * rb], None, False, tb, out=out[s * M:(s + 1) * M])
        ctx.dims = (nseg, tb, ra, rb)
        ctx.save_for_backward(A, B)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        A, B = ctx.saved_tensors
        nseg, tb, ra, rb = ctx.dims
        g = _c(g)
        M = ra
        dA = torch.empty_like(A) if ctx.needs_input_grad[0] else None
        dB = torch.empty_like(B) if ctx.needs_input_grad[1] else None
        for s in range(nseg):
            As, Bs, gs = A[s * ra:(s + 1) * ra], B[s * rb:(s + 1) * rb], g[s * M:(s + 1) * M]
            if dA is not None:      # C = A op(B): dA = g op(B)^T
                gemm_raw(gs, Bs, None, False, not tb, out=dA[s * ra:(s + 1) * ra])
            if dB is not None:      # tb: C = A B^T -> dB = g^T A;  else C = A B -> dB = A^T g
                if tb:
                    gemm_raw(gs, As, None, True, False, out=dB[s * rb:(s + 1) * rb])
                else:
                    gemm_raw(As, gs, None, True, False, out=dB[s * rb:(s + 1) * rb])
        return dA, dB, None, None


class BnActSoftmaxMulKTrain(Function):
    """prod = lrelu(bn_y(xy)) * softmax_k(lrelu(bn_w(xw))) with both train-mode BatchNorm2d + LeakyReLU
    applications folded into the loads (EdgeBlock, Generator.py:78-82): the two normalised [E, C] tensors are
    never written (4 fewer full-tensor passes per EdgeBlock forward).  xw, xy are the PRE-normalisation
    tensors [P*k, C].  `stats_w`: (mean, rstd, var) of xw when the producing GEMM's epilogue already has them.
    Returns prod.  First-order only."""

    @staticmethod
    def forward(ctx, xw, gamma_w, beta_w, xy, gamma_y, beta_y, eps_w, eps_y, slope, k, run_w=None, run_y=None,
                stats_w=None):
        xw, xy = _c(_rows2d(xw)), _c(_rows2d(xy))
        E, C = xw.shape
        if stats_w is not None:           # (mean, rstd, var) from the producing GEMM's epilogue (running stats done there)
            mean_w, rstd_w, var_w = stats_w
        else:
            mean_w, rstd_w, var_w = col_stats(xw, E, eps_w, run_w)
        mean_y, rstd_y, var_y = col_stats(xy, E, eps_y, run_y)
        w = torch.empty_like(xw)
        prod = torch.empty_like(xw)
        L().bn_softmax_mul_k(xw.data_ptr(), xy.data_ptr(), E // k, k, C, mean_w.data_ptr(), rstd_w.data_ptr(),
                             gamma_w.data_ptr(), beta_w.data_ptr(), mean_y.data_ptr(), rstd_y.data_ptr(),
                             gamma_y.data_ptr(), beta_y.data_ptr(), slope, w.data_ptr(), prod.data_ptr(), _stream())
        ctx.k, ctx.slope = k, slope
        ctx.save_for_backward(xw, xy, w, gamma_w, beta_w, gamma_y, beta_y, mean_w, rstd_w, mean_y, rstd_y)
        ctx.set_materialize_grads(False)
        return prod

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        xw, xy, w, gamma_w, beta_w, gamma_y, beta_y, mean_w, rstd_w, mean_y, rstd_y = ctx.saved_tensors
        if g is None:
            return (None,) * 13
        g = _c(g)
        E, C = xw.shape
        dwa = torch.empty_like(xw)
        dya = torch.empty_like(xw)
        L().bn_softmax_mul_k_bwd(g.data_ptr(), xy.data_ptr(), w.data_ptr(), E // ctx.k, ctx.k, C, mean_y.data_ptr(),
                                 rstd_y.data_ptr(), gamma_y.data_ptr(), beta_y.data_ptr(), ctx.slope, dwa.data_ptr(),
                                 dya.data_ptr(), _stream())
        acc_w = (gamma_w, beta_w) if ctx.needs_input_grad[1] and ctx.needs_input_grad[2] and _direct_ok(gamma_w, beta_w) else None
        acc_y = (gamma_y, beta_y) if ctx.needs_input_grad[4] and ctx.needs_input_grad[5] and _direct_ok(gamma_y, beta_y) else None
        dxw, sg_w, sgx_w = _norm_bwd(dwa, xw, ctx.slope, E, mean_w, rstd_w, gamma_w, beta_w, acc_w)
        del dwa
        dxy, sg_y, sgx_y = _norm_bwd(dya, xy, ctx.slope, E, mean_y, rstd_y, gamma_y, beta_y, acc_y)
        return (dxw, None if acc_w else sgx_w.view(-1), None if acc_w else sg_w.view(-1),
                dxy, None if acc_y else sgx_y.view(-1), None if acc_y else sg_y.view(-1)) + (None,) * 7


FUSE_EDGE_ATTENTION = _os.environ.get("SPGAN_FUSE_EDGE_ATTENTION", "1") != "0"


def _bn_finalize(cs, cq, R, bn_gamma, bn_beta, eps, run, want_tables=True):
    """Per-CTA column partials (sum, sum of squares) -> (mean, rstd, var, scale, shift) [1, C]; advances the running
    statistics when `run` = bn_running(bn)."""
    rows, C = cs.shape
    dev = cs.device
    mean = torch.empty((1, C), device=dev, dtype=torch.float32)
    rstd, var = torch.empty_like(mean), torch.empty_like(mean)
    scale = torch.empty_like(mean) if want_tables else None
    shift = torch.empty_like(mean) if want_tables else None
    L().bn_finalize(cs.data_ptr(), cq.data_ptr(), rows, C, R, float(eps), bn_gamma.data_ptr(), bn_beta.data_ptr(),
                    mean.data_ptr(), rstd.data_ptr(), var.data_ptr(), scale.data_ptr() if want_tables else None,
                    shift.data_ptr() if want_tables else None, run[3] if run else 0.0,
                    run[0].data_ptr() if run else None, run[1].data_ptr() if run else None,
                    run[2].data_ptr() if run else None, _stream())
    return mean, rstd, var, scale, shift


def _edge_combine_stats(pc, pn, bias, idx, N, k):
    """spgan_edge_combine_stats -> (out [P*k, C], col_sum, col_sqsum partial rows)."""
    P, C = pn.shape
    rows = L().edge_stats_rows(P, C)
    out = torch.empty((P * k, C), device=pn.device, dtype=torch.float32)
    cs = torch.empty((rows, C), device=pn.device, dtype=torch.float32)
    cq = torch.empty_like(cs)
    L().edge_combine_stats(pc.data_ptr() if pc is not None else None, pn.data_ptr(), idx.data_ptr(),
                           bias.data_ptr() if bias is not None else None, P, N, k, C, out.data_ptr(), cs.data_ptr(),
                           cq.data_ptr(), _stream())
    return out, cs, cq


FUSE_EDGE_STATS = _os.environ.get("SPGAN_FUSE_EDGE_STATS", "1") != "0"


def edge_stats_fusable(P, C, bn):
    """Can the edge gather in front of `bn` also produce bn's batch statistics (spgan_edge_combine_stats)?"""
    return (FUSE_EDGE_STATS and not _TWICE_DIFFERENTIABLE and bn.training and bn.track_running_stats
            and bn.momentum is not None and L().edge_stats_rows(P, C) != 0)


class EdgeCombineStatsTrain(Function):
    """EdgeCombine in front of a train-mode BatchNorm: the gather kernel also accumulates the column statistics of
    what it writes, so the BatchNorm's statistics pass over the [P*k, C] tensor disappears (Generator.py:56-58,66-68).
    Returns (out, mean, rstd, var, scale, shift); the running statistics of the BatchNorm are advanced here."""

    @staticmethod
    def forward(ctx, pc, pn, bias, idx, N, k, gamma, beta, eps, run):
        pn = _c(_rows2d(pn))
        P, C = pn.shape
        if pc is not None:
            pc = _c(pc)
        if bias is not None:
            bias = _c(bias)
        out, cs, cq = _edge_combine_stats(pc, pn, bias, idx, N, k)
        mean, rstd, var, scale, shift = _bn_finalize(cs, cq, P * k, gamma, beta, eps, run)
        ctx.dims = (P, C, N, k)
        ctx.has_pc = pc is not None
        ctx.save_for_backward(idx)
        ctx.mark_non_differentiable(mean, rstd, var, scale, shift)
        ctx.set_materialize_grads(False)
        return out, mean, rstd, var, scale, shift

    @staticmethod
    @once_differentiable
    def backward(ctx, g, *_unused):
        (idx,) = ctx.saved_tensors
        P, C, N, k = ctx.dims
        if g is None:
            return (None,) * 10
        g = _c(g)
        dpc = torch.empty((P, C), device=g.device, dtype=torch.float32) if (ctx.has_pc and ctx.needs_input_grad[0]) else None
        dpn = torch.empty((P, C), device=g.device, dtype=torch.float32)
        L().edge_combine_bwd(g.data_ptr(), idx.data_ptr(), P, N, k, C, dpc.data_ptr() if dpc is not None else None,
                             dpn.data_ptr(), _stream())
        # the bias feeds a train-mode BatchNorm: its gradient is exactly zero (column sums of a BN input gradient)
        return (dpc, dpn) + (None,) * 8


class EdgeGatherBnActLinearTrain(Function):
    """conv_w of EdgeBlock as ONE node (Generator.py:56-62,78), train mode, first order:

        w0 = p1[nbr] - p1[p] + b0                (edge gather; BatchNorm statistics from the same kernel)
        z  = lrelu(bn0(w0)) @ W^T + b            (BN + LeakyReLU in the GEMM's operand converter; statistics of z, the
                                                  next BatchNorm's input, from its epilogue)

    Backward: the weight gradient re-forms the activated operand inside its own converter (no norm_apply pass), and the
    BatchNorm backward of bn0 is applied inside the scatter of the gather's backward (d w0 is never written).
    Returns (z, mean_z, rstd_z, var_z, scale_z, shift_z)."""

    @staticmethod
    def forward(ctx, p1, bias0, idx, N, k, gamma0, beta0, eps0, run0, W, bias, slope, next_bn, zero_bias_grad):
        p1 = _c(_rows2d(p1))
        P, C0 = p1.shape
        if bias0 is not None:
            bias0 = _c(bias0)
        w0, cs, cq = _edge_combine_stats(None, p1, bias0, idx, N, k)
        E = P * k
        mean0, rstd0, _, scale0, shift0 = _bn_finalize(cs, cq, E, gamma0, beta0, eps0, run0)
        res = gemm_fused_raw(w0, _wmat(W, None), bias, tb=True, a_scale=scale0, a_shift=shift0, a_slope=slope,
                             want_stats=True, wcache=True)
        if res is None:
            raise RuntimeError("EdgeGatherBnActLinearTrain: shape outside spgan_gemm_fused's envelope")
        z, cs2, cq2 = res
        m2, r2, v2, sc2, sh2 = _bn_finalize(cs2, cq2, E, next_bn.weight, next_bn.bias, next_bn.eps, bn_running(next_bn))
        ctx.dims, ctx.slope, ctx.has_bias, ctx.zero_bias_grad = (P, C0, N, k), slope, bias is not None, zero_bias_grad
        ctx.mark_non_differentiable(m2, r2, v2, sc2, sh2)
        ctx.set_materialize_grads(False)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(w0, idx, mean0, rstd0, gamma0, beta0, scale0, shift0, W)
        return z, m2, r2, v2, sc2, sh2

    @staticmethod
    @once_differentiable
    def backward(ctx, gz, *_unused):
        if gz is None:
            return (None,) * 14
        w0, idx, mean0, rstd0, gamma0, beta0, scale0, shift0, W = ctx.saved_tensors
        P, C0, N, k = ctx.dims
        E = P * k
        gz = _c(gz)
        params_too = not _INPUT_GRAD_ONLY
        dW = db = dgamma = dbeta = dp1 = None
        if ctx.needs_input_grad[9] and params_too:
            dW = wgrad_bn_act(gz, w0, scale0, shift0, ctx.slope, mean0, rstd0, gamma0, beta0, W, _direct_ok(W))
        if ctx.has_bias and ctx.needs_input_grad[10] and params_too and not ctx.zero_bias_grad:
            db = ColSum.apply(gz, E).view(-1)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[5] or ctx.needs_input_grad[6]:
            ga = gemm_raw(gz, _wmat(W, None), None, False, False, wcache=True)         # d(activated w0)
            acc = params_too and ctx.needs_input_grad[5] and ctx.needs_input_grad[6] and _direct_ok(gamma0, beta0)
            sg = torch.empty((1, C0), device=ga.device, dtype=torch.float32)
            sgx = torch.empty_like(sg)
            ws = _ws(E, C0, E, 2, ga.device)
            if acc:
                L().norm_bwd_reduce_acc(ga.data_ptr(), w0.data_ptr(), ctx.slope, E, C0, mean0.data_ptr(), rstd0.data_ptr(),
                                        gamma0.data_ptr(), beta0.data_ptr(), sg.data_ptr(), sgx.data_ptr(),
                                        beta0.grad.data_ptr(), gamma0.grad.data_ptr(), ws.data_ptr(), _stream())
            else:
                L().norm_bwd_reduce(ga.data_ptr(), w0.data_ptr(), ctx.slope, E, C0, E, mean0.data_ptr(), rstd0.data_ptr(),
                                    gamma0.data_ptr(), beta0.data_ptr(), sg.data_ptr(), sgx.data_ptr(), ws.data_ptr(), _stream())
                if params_too:
                    dgamma, dbeta = sgx.view(-1), sg.view(-1)
            if ctx.needs_input_grad[0]:
                dp1 = torch.empty((P, C0), device=ga.device, dtype=torch.float32)
                L().edge_combine_bwd_bn(ga.data_ptr(), w0.data_ptr(), idx.data_ptr(), P, N, k, C0, mean0.data_ptr(),
                                        rstd0.data_ptr(), gamma0.data_ptr(), beta0.data_ptr(), sg.data_ptr(), sgx.data_ptr(),
                                        ctx.slope, None, dp1.data_ptr(), _stream())
        # bias0 feeds a train-mode BatchNorm: exactly zero gradient
        return dp1, None, None, None, None, dgamma, dbeta, None, None, dW, db, None, None, None


def edge_gather_bn_act_linear(p1, bias0, idx, N, k, bn0, weight, bias, slope, next_bn, zero_bias_grad):
    """-> (z, (mean, rstd, scale, shift) of z, var of z): see EdgeGatherBnActLinearTrain."""
    z, m2, r2, v2, sc2, sh2 = EdgeGatherBnActLinearTrain.apply(p1, bias0, idx, N, k, bn0.weight, bn0.bias, bn0.eps,
                                                              bn_running(bn0), weight, bias, slope, next_bn, zero_bias_grad)
    return z, (m2, r2, sc2, sh2), v2


def edge_combine_bn_stats(pc, pn, bias, idx, N, k, bn):
    """-> (pre-BN edge tensor, (mean, rstd, scale, shift), var) for `bn` (train mode; see edge_stats_fusable)."""
    out, mean, rstd, var, scale, shift = EdgeCombineStatsTrain.apply(pc, pn, bias, idx, N, k, bn.weight, bn.bias, bn.eps,
                                                                     bn_running(bn))
    return out, (mean, rstd, scale, shift), var


class EdgeAttentionTrain(Function):
    """EdgeBlock from the per-point conv_x projections to the attention product (Generator.py:66-69,78-82), train mode:

        y    = a[p] + d[nbr] - d[p] + bias_x            (edge gather, its BatchNorm statistics from the same kernel)
        prod = lrelu(bn_y(y)) * softmax_k(lrelu(bn_w(xw)))

    Backward: ONE pass produces the gradients w.r.t. both activated tensors and the four column sums the two BatchNorm
    backwards need (no norm_bwd_reduce passes); the conv_x branch's BatchNorm backward is applied inside the scatter
    of the gather's backward (d y is never written).  Returns prod.  First-order only."""

    @staticmethod
    def forward(ctx, xw, mean_w, rstd_w, gamma_w, beta_w, a, d, bias_x, idx, N, k, gamma_y, beta_y, eps_y, slope, run_y):
        xw = _c(_rows2d(xw))
        a, d = _c(_rows2d(a)), _c(_rows2d(d))
        E, C = xw.shape
        P = d.shape[0]
        if bias_x is not None:
            bias_x = _c(bias_x)
        xy, cs, cq = _edge_combine_stats(a, d, bias_x, idx, N, k)
        mean_y, rstd_y, _, _, _ = _bn_finalize(cs, cq, E, gamma_y, beta_y, eps_y, run_y, want_tables=False)
        # the softmax weights are kept for the backward pass only: a no-grad forward (the critic phase's generator
        # pass, model.py:246-248) does not write them
        need_bwd = any(ctx.needs_input_grad)
        w = torch.empty_like(xw) if need_bwd else None
        prod = torch.empty_like(xw)
        L().bn_softmax_mul_k(xw.data_ptr(), xy.data_ptr(), P, k, C, mean_w.data_ptr(), rstd_w.data_ptr(),
                             gamma_w.data_ptr(), beta_w.data_ptr(), mean_y.data_ptr(), rstd_y.data_ptr(),
                             gamma_y.data_ptr(), beta_y.data_ptr(), slope, w.data_ptr() if need_bwd else None,
                             prod.data_ptr(), _stream())
        ctx.dims, ctx.slope = (P, C, N, k), slope
        if not need_bwd:
            return prod
        ctx.save_for_backward(xw, xy, w, idx, gamma_w, beta_w, gamma_y, beta_y, mean_w, rstd_w, mean_y, rstd_y)
        ctx.set_materialize_grads(False)
        return prod

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        xw, xy, w, idx, gamma_w, beta_w, gamma_y, beta_y, mean_w, rstd_w, mean_y, rstd_y = ctx.saved_tensors
        if g is None:
            return (None,) * 16
        g = _c(g)
        P, C, N, k = ctx.dims
        E = P * k
        dev = g.device
        rows = L().attn_bwd_rows(P, k, C)
        part = torch.empty((rows, 4, C), device=dev, dtype=torch.float32)
        dwa = torch.empty_like(xw)
        dya = torch.empty_like(xw)
        L().bn_softmax_mul_k_bwd_stats(g.data_ptr(), xw.data_ptr(), xy.data_ptr(), w.data_ptr(), P, k, C, mean_w.data_ptr(),
                                       rstd_w.data_ptr(), gamma_w.data_ptr(), beta_w.data_ptr(), mean_y.data_ptr(),
                                       rstd_y.data_ptr(), gamma_y.data_ptr(), beta_y.data_ptr(), ctx.slope,
                                       dwa.data_ptr(), dya.data_ptr(), part.data_ptr(), _stream())
        params_too = not _INPUT_GRAD_ONLY
        acc_w = params_too and ctx.needs_input_grad[3] and ctx.needs_input_grad[4] and _direct_ok(gamma_w, beta_w)
        acc_y = params_too and ctx.needs_input_grad[11] and ctx.needs_input_grad[12] and _direct_ok(gamma_y, beta_y)
        sums = torch.empty((4, C), device=dev, dtype=torch.float32)          # sg_w, sgx_w, sg_y, sgx_y
        L().partials_finalize(part.data_ptr(), rows, 4, C, sums.data_ptr(),
                              beta_w.grad.data_ptr() if acc_w else None, gamma_w.grad.data_ptr() if acc_w else None,
                              beta_y.grad.data_ptr() if acc_y else None, gamma_y.grad.data_ptr() if acc_y else None,
                              _stream())
        del part
        dxw = None
        if ctx.needs_input_grad[0]:
            dxw = torch.empty_like(xw)
            L().norm_bwd_apply(dwa.data_ptr(), xw.data_ptr(), ctx.slope, E, C, E, mean_w.data_ptr(), rstd_w.data_ptr(),
                               gamma_w.data_ptr(), beta_w.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(),
                               dxw.data_ptr(), _stream())
        del dwa
        da = torch.empty((P, C), device=dev, dtype=torch.float32) if ctx.needs_input_grad[5] else None
        dd = torch.empty((P, C), device=dev, dtype=torch.float32)
        L().edge_combine_bwd_bn(dya.data_ptr(), xy.data_ptr(), idx.data_ptr(), P, N, k, C, mean_y.data_ptr(),
                                rstd_y.data_ptr(), gamma_y.data_ptr(), beta_y.data_ptr(), sums[2].data_ptr(),
                                sums[3].data_ptr(), ctx.slope, da.data_ptr() if da is not None else None, dd.data_ptr(),
                                _stream())
        dgw = dbw = dgy = dby = None
        if params_too and not acc_w:
            dgw, dbw = sums[1].clone(), sums[0].clone()
        if params_too and not acc_y:
            dgy, dby = sums[3].clone(), sums[2].clone()
        # bias_x feeds a train-mode BatchNorm: its gradient is exactly zero
        return (dxw, None, None, dgw, dbw, da, dd, None, None, None, None, dgy, dby, None, None, None)


def edge_attention_stats_fusable(P, C, k, bn_w, bn_y):
    return (edge_attention_fusable(bn_w, bn_y, k) and edge_stats_fusable(P, C, bn_y) and bn_w.momentum is not None
            and L().attn_bwd_rows(P, k, C) != 0)


def edge_attention(xw, stats_w, bn_w, a, d, bias_x, idx, N, k, bn_y, slope):
    """prod of EdgeAttentionTrain; stats_w = (mean, rstd, var) of xw from its producer, or None (computed here)."""
    if stats_w is None:
        xw2 = _c(_rows2d(xw))
        mean_w, rstd_w, _ = col_stats(xw2.detach(), xw2.shape[0], bn_w.eps, bn_running(bn_w))
    else:
        mean_w, rstd_w = stats_w[0], stats_w[1]
    return EdgeAttentionTrain.apply(xw, mean_w, rstd_w, bn_w.weight, bn_w.bias, a, d, bias_x, idx, N, k, bn_y.weight,
                                    bn_y.bias, bn_y.eps, slope, bn_running(bn_y))




def edge_attention_fusable(bn_w, bn_y, k):
    return (FUSE_EDGE_ATTENTION and k <= 16 and bn_w.training and bn_y.training and bn_w.track_running_stats
            and bn_y.track_running_stats and not _TWICE_DIFFERENTIABLE)


def bn_act_softmax_mul_k(xw, bn_w, xy, bn_y, slope, k, stats_w=None):
    """EdgeBlock's y * softmax_k(w) over the two BatchNorm2d + LeakyReLU branches (Generator.py:78-82), fused on
    the train-mode path; eval mode runs the generic chain."""
    if edge_attention_fusable(bn_w, bn_y, k):
        return BnActSoftmaxMulKTrain.apply(xw, bn_w.weight, bn_w.bias, xy, bn_y.weight, bn_y.bias, bn_w.eps, bn_y.eps,
                                           slope, k, None if stats_w is not None else bn_running(bn_w),
                                           bn_running(bn_y), stats_w)
    return SoftmaxMulK.apply(batch_norm_act(xw, bn_w, slope), batch_norm_act(xy, bn_y, slope), k)


class KMax(Function):
    """max over the k neighbours: [P*k, C] -> [P, C] (torch.max(x, 3), modules.py:794)."""

    @staticmethod
    def forward(ctx, x, k):
        x = _c(_rows2d(x))
        E, C = x.shape
        P = E // k
        out = torch.empty((P, C), device=x.device, dtype=torch.float32)
        arg = torch.empty((P, C), device=x.device, dtype=torch.int32)
        L().kmax(x.data_ptr(), P, k, C, out.data_ptr(), arg.data_ptr(), _stream())
        ctx.k = k
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        g = _c(g)
        P, C = g.shape
        dx = torch.empty((P * ctx.k, C), device=g.device, dtype=torch.float32)
        L().kmax_scatter(g.data_ptr(), arg.data_ptr(), P, ctx.k, C, dx.data_ptr(), _stream())
        return dx, None


class EdgeCombine(Function):
    """out[p*k+r] = pc[p] + pn[nbr(p,r)] - pn[p] + bias: a 1x1 conv over [centre, nbr - centre]
    expressed through per-point projections (pc may be None for a conv on the difference half)."""

    @staticmethod
    def forward(ctx, pc, pn, bias, idx, N, k, zero_bias_grad=False):
        ctx.zero_bias_grad = zero_bias_grad
        pn = _c(_rows2d(pn))
        P, C = pn.shape
        if pc is not None:
            pc = _c(pc)
        if bias is not None:
            bias = _c(bias)
        out = torch.empty((P * k, C), device=pn.device, dtype=torch.float32)
        L().edge_combine(pc.data_ptr() if pc is not None else None, pn.data_ptr(), idx.data_ptr(),
                         bias.data_ptr() if bias is not None else None, P, N, k, C, out.data_ptr(), _stream())
        ctx.dims = (P, C, N, k)
        ctx.has = (pc is not None, bias is not None)
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        P, C, N, k = ctx.dims
        g = _c(g)
        has_pc, has_bias = ctx.has
        dpc = torch.empty((P, C), device=g.device, dtype=torch.float32) if (has_pc and ctx.needs_input_grad[0]) else None
        dpn = torch.empty((P, C), device=g.device, dtype=torch.float32)
        L().edge_combine_bwd(g.data_ptr(), idx.data_ptr(), P, N, k, C, dpc.data_ptr() if dpc is not None else None,
                             dpn.data_ptr(), _stream())
        db = None
        if has_bias and ctx.needs_input_grad[2]:
            db = None if ctx.zero_bias_grad else ColSum.apply(g, g.shape[0]).view(-1)
        return dpc, dpn, db, None, None, None, None


# =========================================================================================
# kNN graph (no gradient: indices)
# =========================================================================================
def knn_indices(x_bcn, k, want_ee=False, main_cols=-1):
    """x [B, C, N] contiguous fp32 -> idx int32 [B, N, k] (ranks 1..k of the reference's sorted
    distance rows, modules.py:695-704); optionally the grouped edge features [B, 2C, N, k]."""
    x = _c(x_bcn.detach())
    B, C, N = x.shape
    xs = torch.empty((B, N), device=x.device, dtype=torch.float32)
    L().sqnorm(x.data_ptr(), B, C, N, main_cols, xs.data_ptr(), _stream())
    idx = torch.empty((B, N, k), device=x.device, dtype=torch.int32)
    ee = torch.empty((B, 2 * C, N, k), device=x.device, dtype=torch.float32) if want_ee else None
    L().knn_group(x.data_ptr(), xs.data_ptr(), B, C, N, k, idx.data_ptr(),
                  ee.data_ptr() if ee is not None else None, _stream())
    return (idx, ee) if want_ee else idx


KNN_TC = _os.environ.get("SPGAN_KNN_TC", "1") != "0"
LAST_KNN_WORKSPACE = None         # int32 view: [1] = queries of the last spgan_knn_rows call ranked by the exact scan


def knn_indices_rows(x_rows, B, N, k):
    """Point-major rows [B*N, C] -> idx int32 [B, N, k], the same neighbour lists as knn_indices on the [B, C, N]
    view (bit for bit).  Runs the tensor-core filter + exact refine kernel when the shape allows, else transposes and
    calls the CUDA-core kernel."""
    global LAST_KNN_WORKSPACE
    x = _c(x_rows.detach())
    C = x.shape[1]
    if KNN_TC and C % 4 != 0 and L().knn_rows_workspace(B, (C + 3) // 4 * 4, N, k) != 0:
        # zero-pad the channels to a multiple of 4 (TMA row pitch): fma(0, 0, acc) = acc and acc + 0*0 = acc exactly,
        # so the recipe's distances -- and the lists -- are unchanged (the xyz sphere of EdgeConv1: C = 3 -> 4)
        pad = (C + 3) // 4 * 4 - C
        xp = torch.empty((x.shape[0], C + pad), device=x.device, dtype=torch.float32)
        zeros = full((1, pad), 0.0, x.device)
        L().concat_cols(x.data_ptr(), C, C, zeros.data_ptr(), 0, 0, x.shape[0], pad, x.shape[0], xp.data_ptr(), _stream())
        x, C = xp, C + pad
    ws_bytes = L().knn_rows_workspace(B, C, N, k) if KNN_TC else 0
    if ws_bytes == 0:
        return knn_indices(RowsToBcn.apply(x, B, C, N), k)
    xs = torch.empty((B, N), device=x.device, dtype=torch.float32)
    L().sqnorm_pm(x.data_ptr(), B, C, N, -1, xs.data_ptr(), _stream())
    idx = torch.empty((B, N, k), device=x.device, dtype=torch.int32)
    ws = torch.empty(ws_bytes // 4, device=x.device, dtype=torch.int32)
    LAST_KNN_WORKSPACE = ws
    L().knn_rows(x.data_ptr(), xs.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
    return idx


def idx_to_int64(idx32):
    out = torch.empty(idx32.shape, device=idx32.device, dtype=torch.int64)
    L().idx32_to_idx64(idx32.data_ptr(), out.data_ptr(), idx32.numel(), _stream())
    return out


def idx_to_int32(idx64):
    idx64 = idx64.contiguous()
    out = torch.empty(idx64.shape, device=idx64.device, dtype=torch.int32)
    L().idx64_to_idx32(idx64.data_ptr(), out.data_ptr(), idx64.numel(), _stream())
    return out


class Group(Function):
    """ee[B, 2C, N, k] = cat(centre, neighbour - centre) for a given neighbour list
    (modules.py:706-720), or cat(neighbour - centre, centre) with diff_first (get_graph_feature,
    modules.py:678); backward scatters with the edge-aggregation kernel."""

    @staticmethod
    def forward(ctx, x, idx32, k, diff_first=False):
        x = _c(x)
        B, C, N = x.shape
        ee = torch.empty((B, 2 * C, N, k), device=x.device, dtype=torch.float32)
        if diff_first:
            L().group_ex(x.data_ptr(), idx32.data_ptr(), B, C, N, k, 1, ee.data_ptr(), _stream())
        else:
            L().group(x.data_ptr(), idx32.data_ptr(), B, C, N, k, ee.data_ptr(), _stream())
        ctx.dims = (B, C, N, k)
        ctx.diff_first = bool(diff_first)
        ctx.save_for_backward(idx32)
        return ee

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (idx32,) = ctx.saved_tensors
        B, C, N, k = ctx.dims
        g = _c(g)
        # d x[b,c,i] = sum_r g_ctr[b,c,i,r] - sum_r g_dif[b,c,i,r] + sum_{(p,r): nbr(p,r)=i} g_dif[b,c,p,r]
        # route through edge-major rows: [B, 2C, N*k] -> rows [B*N*k, 2C]
        rows = torch.empty((B * N * k, 2 * C), device=g.device, dtype=torch.float32)
        g3 = g.view(B, 2 * C, N * k)
        L().bcn_to_rows(g3.data_ptr(), g3.stride(0), g3.stride(1), g3.stride(2), B, 2 * C, N * k, rows.data_ptr(),
                        _stream())
        g_ctr = contiguous(rows[:, C:] if ctx.diff_first else rows[:, :C])
        g_dif = contiguous(rows[:, :C] if ctx.diff_first else rows[:, C:])
        dpc = torch.empty((B * N, C), device=g.device, dtype=torch.float32)
        dpn = torch.empty((B * N, C), device=g.device, dtype=torch.float32)
        junk = torch.empty((B * N, C), device=g.device, dtype=torch.float32)
        L().edge_combine_bwd(g_dif.data_ptr(), idx32.data_ptr(), B * N, N, k, C, None, dpn.data_ptr(), _stream())
        L().colsum(g_ctr.data_ptr(), B * N * k, C, k, dpc.data_ptr(),
                   _ws(B * N * k, C, k, 1, g.device).data_ptr(), _stream())
        L().axpby(1.0, dpc.data_ptr(), 1.0, dpn.data_ptr(), junk.data_ptr(), dpc.numel(), _stream())
        dx = torch.empty((B, C, N), device=g.device, dtype=torch.float32)
        L().rows_to_bcn(junk.data_ptr(), B, C, N, dx.data_ptr(), _stream())
        return dx, None, None, None


# =========================================================================================
# the other kNN / grouping entry points: raw operators, wrapped in pointnet_util.py
# =========================================================================================
def sqnorm_bcn(x_bcn, main_cols=-1):
    """|x[b,:,n]|^2 of a contiguous [B, C, N] tensor in the channel-first reduction order (modules.py:642,697)."""
    B, C, N = x_bcn.shape
    xs = torch.empty((B, N), device=x_bcn.device, dtype=torch.float32)
    L().sqnorm(x_bcn.data_ptr(), B, C, N, main_cols, xs.data_ptr(), _stream())
    return xs


def sqnorm_rows(p_bnc):
    """|p[b,n,:]|^2 of contiguous point-major rows [B, N, C] (pointnet_util.py:38-39)."""
    B, N, C = p_bnc.shape
    xs = torch.empty((B, N), device=p_bnc.device, dtype=torch.float32)
    L().sqnorm_rows(p_bnc.data_ptr(), B * N, C, xs.data_ptr(), _stream())
    return xs


def knn_query(xq_bcn, xsq, xc_bcn, xsc, k, first_rank, cand_norm_first):
    """-> idx int32 [B, Nq, k]: ranks first_rank .. first_rank+k-1 of every query's ascending distance row."""
    B, C, Nq = xq_bcn.shape
    Nc = xc_bcn.shape[2]
    idx = torch.empty((B, Nq, k), device=xq_bcn.device, dtype=torch.int32)
    L().knn_query(xq_bcn.data_ptr(), xsq.data_ptr(), Nq, xc_bcn.data_ptr(), xsc.data_ptr(), Nc, B, C, k,
                  int(first_rank), int(bool(cand_norm_first)), idx.data_ptr(), _stream())
    return idx


def pairwise_sqdist(xq_bcn, xsq, xc_bcn, xsc, cand_norm_first=False):
    """-> dist fp32 [B, Nq, Nc], (-2 q.c + first norm) + second norm."""
    B, C, Nq = xq_bcn.shape
    Nc = xc_bcn.shape[2]
    out = torch.empty((B, Nq, Nc), device=xq_bcn.device, dtype=torch.float32)
    L().pairwise_sqdist(xq_bcn.data_ptr(), xsq.data_ptr(), Nq, xc_bcn.data_ptr(), xsc.data_ptr(), Nc, B, C,
                        int(bool(cand_norm_first)), out.data_ptr(), _stream())
    return out


class GatherRows(Function):
    """out[b, s, :] = points[b, idx[b, s], :] (index_points, pointnet_util.py:43-59); backward scatter-adds."""

    @staticmethod
    def forward(ctx, points, idx):
        points = _c(points, "points")
        if idx.dtype not in (torch.int32, torch.int64) or not idx.is_cuda:
            raise RuntimeError("spgan_b200: idx must be a CUDA int32/int64 tensor")
        idx = idx.contiguous()
        B, N, C = points.shape
        S = idx.numel() // max(B, 1)
        out = torch.empty((B, S, C), device=points.device, dtype=torch.float32)
        status = torch.zeros(1, device=points.device, dtype=torch.int32)
        L().gather_rows(points.data_ptr(), idx.data_ptr(), int(idx.dtype == torch.int64), B, N, S, C, out.data_ptr(),
                        status.data_ptr(), _stream())
        ctx.dims = (B, N, S, C)
        ctx.save_for_backward(idx)
        ctx.mark_non_differentiable(status)
        return out, status          # status[0] != 0: some index fell outside [0, N)

    @staticmethod
    @once_differentiable
    def backward(ctx, g, _gstatus):
        (idx,) = ctx.saved_tensors
        B, N, S, C = ctx.dims
        g = _c(g)
        dp = full((B, N, C), 0.0, g.device)
        L().scatter_add_rows(g.data_ptr(), idx.data_ptr(), int(idx.dtype == torch.int64), B, N, S, C, dp.data_ptr(),
                             _stream())
        return dp, None


# =========================================================================================
# losses / penalty
# =========================================================================================
class MeanScale(Function):
    """scale * mean(x) as a 0-d tensor (wgan loss terms, loss_utils.py:728-730, 859-863)."""

    @staticmethod
    def forward(ctx, x, scale_):
        x = _c(x)
        ctx.n, ctx.scale, ctx.shape = x.numel(), scale_, x.shape
        out = torch.empty((), device=x.device, dtype=torch.float32)
        L().mean(x.data_ptr(), x.numel(), scale_, 0, out.data_ptr(), _stream())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _c(g)
        coef = torch.empty((1,), device=g.device, dtype=torch.float32)
        L().axpby(ctx.scale / ctx.n, g.data_ptr(), 0.0, None, coef.data_ptr(), 1, _stream())
        out = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        L().bcast_segvec(coef.data_ptr(), ctx.n, 1, ctx.n, out.data_ptr(), _stream())
        return out, None


class GradPenalty(Function):
    """lambda * mean_b(((||g_b||_2 - gamma) / gamma)^2) over g [B, D] (gradient_penalty.py:35)."""

    @staticmethod
    def forward(ctx, g, gamma, lam):
        g = _c(g)
        B, D = g.shape
        norms = torch.empty((B,), device=g.device, dtype=torch.float32)
        out = torch.empty((), device=g.device, dtype=torch.float32)
        L().gp_penalty(g.data_ptr(), B, D, gamma, lam, norms.data_ptr(), out.data_ptr(), _stream())
        ctx.consts = (gamma, lam)
        ctx.save_for_backward(g, norms)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        g, norms = ctx.saved_tensors
        gamma, lam = ctx.consts
        gout = _c(gout)
        B, D = g.shape
        dg = torch.empty_like(g)
        L().gp_penalty_bwd(g.data_ptr(), norms.data_ptr(), gout.data_ptr(), B, D, gamma, lam, dg.data_ptr(), _stream())
        return dg, None, None


def gp_interpolate(real, fake, alpha):
    """mix = real + alpha * (fake - real); real/fake [B,C,N] with any strides, alpha [B] (gradient_penalty.py:26)."""
    _chk(real), _chk(fake)
    alpha = _c(alpha.reshape(-1))
    B, C, N = real.shape
    mix = torch.empty((B, C, N), device=real.device, dtype=torch.float32)
    L().gp_interp(real.data_ptr(), real.stride(0), real.stride(1), real.stride(2), fake.data_ptr(), fake.stride(0),
                  fake.stride(1), fake.stride(2), alpha.data_ptr(), B, C, N, mix.data_ptr(), _stream())
    return mix


def fill_(t, v):
    L().fill(t.data_ptr(), t.numel(), float(v), _stream())
    return t


def full(shape, v, device):
    return fill_(torch.empty(shape, device=device, dtype=torch.float32), v)
