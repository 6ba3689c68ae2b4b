"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pairwise Chamfer evaluation.

Follows metrics/evaluation_metrics.py: distChamfer :37-49 (expanded form |x|^2 + |y|^2 - 2 x.y via bmm, fp32),
_pairwise_EMD_CD_ :89-125 (CD half: dl.mean(1) + dr.mean(1) per (sample, ref) pair), lgan_mmd_cov :161-173 and
the 1-NN test knn :129-158.  `pairwise_cd_exact` is the same quantity in float64 with direct differences (the
arithmetic form of the reference's CUDA kernel, metrics/CD_EMD/cd/chamferdist/chamfer.cu:12-134).
Pinned by tests/golden/chamfer.npz, produced by executing the reference's own function bodies
(tests/golden/make_golden_chamfer.py).  Never imported by the product."""
import numpy as np
import torch


def dist_chamfer(a, b):
    """evaluation_metrics.py:37-49 for a [bs, n, 3], b [bs, m, 3]: P[n, m] = (|a_n|^2 + |b_m|^2) - 2 a_n.b_m in
    fp32 with the squared norms read off the diagonals of the Gram matrices (bmm), then the two directional
    minima (over a for each b-point, over b for each a-point)."""
    sq_a = torch.bmm(a, a.transpose(2, 1)).diagonal(dim1=1, dim2=2)
    sq_b = torch.bmm(b, b.transpose(2, 1)).diagonal(dim1=1, dim2=2)
    P = (sq_a[:, :, None] + sq_b[:, None, :]) - 2 * torch.bmm(a, b.transpose(2, 1))
    return P.min(1)[0], P.min(2)[0]


def pairwise_cd(sample_pcs, ref_pcs, batch_size=4):
    """CD half of _pairwise_EMD_CD_ (evaluation_metrics.py:89-125) -> [N_sample, N_ref]."""
    rows = []
    for i in range(sample_pcs.shape[0]):
        cds = []
        for r0 in range(0, ref_pcs.shape[0], batch_size):
            ref = ref_pcs[r0:r0 + batch_size]
            smp = sample_pcs[i].view(1, -1, 3).expand(ref.size(0), -1, -1).contiguous()
            dl, dr = dist_chamfer(smp, ref)
            cds.append((dl.mean(dim=1) + dr.mean(dim=1)).view(1, -1))
        rows.append(torch.cat(cds, dim=1))
    return torch.cat(rows, dim=0)


def pairwise_cd_exact(sample_pcs, ref_pcs):
    """float64, direct differences; clouds may have different point counts."""
    a = np.asarray(sample_pcs, np.float64)
    b = np.asarray(ref_pcs, np.float64)
    out = np.zeros((a.shape[0], b.shape[0]))
    for i in range(a.shape[0]):
        for j in range(b.shape[0]):
            d = ((a[i][:, None, :] - b[j][None, :, :]) ** 2).sum(-1)
            out[i, j] = d.min(1).mean() + d.min(0).mean()
    return out


def lgan_mmd_cov(all_dist):
    """evaluation_metrics.py:161-173."""
    N_sample, N_ref = all_dist.size(0), all_dist.size(1)
    min_val_fromsmp, min_idx = torch.min(all_dist, dim=1)
    min_val, _ = torch.min(all_dist, dim=0)
    return {"lgan_mmd": float(min_val.mean()), "lgan_cov": float(min_idx.unique().view(-1).size(0)) / float(N_ref),
            "lgan_mmd_smp": float(min_val_fromsmp.mean())}


def one_nn_accuracy(Mxx, Mxy, Myy):
    """`acc` of the leave-one-out 1-NN two-sample test, evaluation_metrics.py:129-158 with k = 1, sqrt = False:
    every cloud of the pooled set is labelled by its nearest other cloud."""
    n0, n1 = Mxx.size(0), Myy.size(0)
    label = torch.cat((torch.ones(n0), torch.zeros(n1)))
    M = torch.cat((torch.cat((Mxx, Mxy), 1), torch.cat((Mxy.t(), Myy), 1)), 0).clone()
    M.fill_diagonal_(float("inf"))
    pred = label[M.argmin(dim=0)]
    return float((pred == label).float().mean())
