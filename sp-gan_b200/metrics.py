"""Pairwise Chamfer-distance evaluation (BASELINE configs[4]): the CD half of
metrics/evaluation_metrics.py:89-204 (`_pairwise_EMD_CD_`, `lgan_mmd_cov`, the 1-NN `knn`) and
Common/GAN_metrics.py:658-684 (`pairwise_CD`).

`pairwise_CD` runs the sm_100a kernel (one CTA per cloud pair) and shards the rows of the S x R matrix over
the ranks of an initialised torch.distributed group (one all_gather of the result rows).  The MMD / COV / 1-NN
statistics are O(S*R) reductions of that matrix and run on the host in numpy, like the reference's own
`.unique()` / `.topk()` post-processing."""
import numpy as np
import torch
import torch.distributed as dist

from . import ops


def pairwise_CD(sample_pcs, ref_pcs, batch_size=None, shard=True):
    """sample_pcs [S, N, 3], ref_pcs [R, M, 3] (CUDA fp32) -> all_cd [S, R] with
    all_cd[i, j] = dl.mean() + dr.mean() of distChamfer(sample_i, ref_j) (evaluation_metrics.py:109-112).
    `batch_size` is accepted for signature compatibility (the kernel needs no batching)."""
    a, b = ops._c(sample_pcs), ops._c(ref_pcs)
    S, N, _ = a.shape
    R, M, _ = b.shape
    world = dist.get_world_size() if (shard and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    rows = (S + world - 1) // world                        # rows of the S x R matrix per rank
    i0, i1 = min(S, rank * rows), min(S, (rank + 1) * rows)
    part = torch.zeros((rows, R), device=a.device, dtype=torch.float32) if world > 1 else \
        torch.empty((S, R), device=a.device, dtype=torch.float32)
    if i1 > i0:
        ops.L().pairwise_chamfer(a.data_ptr(), b.data_ptr(), S, R, N, M, i0 * R, (i1 - i0) * R, part.data_ptr(),
                                 None, None, ops._stream())
    if world == 1:
        return part
    full = torch.empty((world * rows, R), device=a.device, dtype=torch.float32)
    dist.all_gather_into_tensor(full, part)
    return full[:S]


def pairwise_EMD(sample_pcs, ref_pcs, batch_size=None, eps=0.005, iters=300, shard=True):
    """The EMD half of `_pairwise_EMD_CD_` (metrics/evaluation_metrics.py:89-125): sample_pcs [S, N, 3], ref_pcs
    [R, N, 3] -> all_emd [S, R] with all_emd[i, j] = mean over the points of sample_i of the matched L2 distance to
    ref_j (what `emd_approx` = match_cost / N returns per pair, :28-36; here the matching is the auction of
    spgan_b200.emd with the reference's eps / iteration count, GAN_metrics.py:406-407).  Rows are sharded over the
    ranks of an initialised process group like pairwise_CD.  `batch_size` bounds the reference clouds per launch
    (default: all R; each launch expands one sample cloud R-fold like the reference's own loop).
    Round-1 status: composed from validated kernels, itself exercised only by the opt-in test
    (SPGAN_TEST_PAIRWISE_EMD=1 pytest tests/test_gpu_chamfer.py)."""
    from .emd import emdFunction
    a, b = ops._c(sample_pcs), ops._c(ref_pcs)
    S, N, _ = a.shape
    R = b.shape[0]
    if b.shape[1] != N:
        raise AssertionError("pairwise_EMD: clouds must have the same number of points")
    world = dist.get_world_size() if (shard and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    rows = (S + world - 1) // world
    i0, i1 = min(S, rank * rows), min(S, (rank + 1) * rows)
    part = torch.zeros((rows if world > 1 else S, R), device=a.device, dtype=torch.float32)
    bs = R if not batch_size else int(batch_size)
    with torch.no_grad():
        for i in range(i0, i1):
            for r0 in range(0, R, bs):
                rb = b[r0:r0 + bs]
                d, _ = emdFunction.apply(a[i:i + 1].expand(rb.shape[0], N, 3), rb, eps, iters)
                part[i - i0, r0:r0 + rb.shape[0]] = d.sqrt().mean(dim=1)      # O(R*N) post-processing of the result
    if world == 1:
        return part
    full = torch.empty((world * rows, R), device=a.device, dtype=torch.float32)
    dist.all_gather_into_tensor(full, part)
    return full[:S]


def dist_chamfer(a, b):
    """distChamfer-style per-pair call (evaluation_metrics.py:37-49) for equal-length batches:
    a [B, N, 3], b [B, M, 3] -> (mean_m-side, mean_n-side) directed means are not what the reference returns
    per point, so this helper returns the two directed MEANS (dl.mean(1), dr.mean(1)) of each pair i == j."""
    a, b = ops._c(a), ops._c(b)
    B, N, _ = a.shape
    M = b.shape[1]
    dl = torch.empty((B,), device=a.device, dtype=torch.float32)
    dr = torch.empty_like(dl)
    for i in range(B):       # diagonal pairs only: pair index i*B + i
        ops.L().pairwise_chamfer(a.data_ptr(), b.data_ptr(), B, B, N, M, i * B + i, 1, None,
                                 dl[i:].data_ptr(), dr[i:].data_ptr(), ops._stream())
    return dl, dr


def lgan_mmd_cov(all_dist):
    """evaluation_metrics.py:161-173 on an [N_sample, N_ref] matrix (torch or numpy) -> dict of floats."""
    d = all_dist.detach().cpu().numpy() if torch.is_tensor(all_dist) else np.asarray(all_dist)
    min_idx = d.argmin(axis=1)
    return {"lgan_mmd": float(d.min(axis=0).mean()), "lgan_cov": float(np.unique(min_idx).size) / float(d.shape[1]),
            "lgan_mmd_smp": float(d.min(axis=1).mean())}


def one_nn_accuracy(Mxx, Mxy, Myy):
    """The `acc` of the 1-NN two-sample test (evaluation_metrics.py:129-158 with k = 1, sqrt = False)."""
    f = lambda t: t.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(t) else np.asarray(t, np.float64)
    Mxx, Mxy, Myy = f(Mxx), f(Mxy), f(Myy)
    n0, n1 = Mxx.shape[0], Myy.shape[0]
    label = np.concatenate([np.ones(n0), np.zeros(n1)])
    M = np.concatenate([np.concatenate([Mxx, Mxy], 1), np.concatenate([Mxy.T, Myy], 1)], 0)
    M = M + np.diag(np.full(n0 + n1, np.inf))
    idx = M.argmin(axis=0)                                  # topk(k=1, dim 0, smallest)
    pred = (label[idx] >= 0.5).astype(np.float64)
    return float((label == pred).mean())
