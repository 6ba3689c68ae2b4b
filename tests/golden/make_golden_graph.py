#!/usr/bin/env python
"""Golden vectors for the reference's other kNN / grouping entry points, produced by the UNMODIFIED
reference functions on CPU.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_graph.py

Two import-time accommodations, neither touches the arithmetic:
  * `get_graph_feature` hard-codes torch.device('cuda') for its index offsets (modules.py:662): the `torch` name
    inside Generation.modules is replaced by a proxy whose `device()` returns the CPU device;
  * Common/pointconv_util.py imports `sklearn.neighbors.kde` (removed from scikit-learn): a stub module with a
    `KernelDensity` name is registered before the import (only `knn_point` is used).
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SPGAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


class _TorchOnCpu:
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


def main():
    stub = types.ModuleType("sklearn.neighbors.kde")
    stub.KernelDensity = object
    sys.modules.setdefault("sklearn.neighbors.kde", stub)
    import Generation.modules as M
    import Common.pointnet_util as PU
    import Common.pointconv_util as PC
    M.torch = _TorchOnCpu()

    g = torch.Generator().manual_seed(81)
    out = {}
    # knn / get_graph_feature: feature space (C=16) and xyz (C=3, clustered so that near ties occur)
    for tag, (B, C, N, k) in {"a": (2, 16, 160, 20), "b": (3, 3, 100, 5), "c": (1, 64, 256, 8)}.items():
        x = torch.randn(B, C, N, generator=g)
        if tag == "b":
            x = (x * 4).round() / 4                       # coarse grid: many exactly tied distances
        inner = -2 * torch.matmul(x.transpose(2, 1), x)   # modules.py:641-643 restated only to STORE the matrix
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        out["knn_%s_x" % tag] = x.numpy()
        out["knn_%s_k" % tag] = np.int32(k)
        out["knn_%s_negdist" % tag] = (-xx - inner - xx.transpose(2, 1)).numpy()
        out["knn_%s_idx" % tag] = M.knn(x, k).numpy()
        if tag != "c":
            out["ggf_%s" % tag] = M.get_graph_feature(x, k=k).contiguous().numpy()
    # square_distance / pairwise_dist / knn_point / index_points: xyz rows and a wider feature case
    for tag, (B, N, S, C, ns) in {"a": (2, 200, 50, 3, 16), "b": (2, 96, 96, 3, 8), "c": (1, 130, 70, 6, 12)}.items():
        xyz = torch.randn(B, N, C, generator=g)
        new = xyz[:, :S].clone() if tag == "b" else torch.randn(B, S, C, generator=g)
        out["pt_%s_xyz" % tag] = xyz.numpy()
        out["pt_%s_new" % tag] = new.numpy()
        out["pt_%s_ns" % tag] = np.int32(ns)
        out["pt_%s_sqdist" % tag] = PU.square_distance(new, xyz).numpy()
        out["pt_%s_pairwise" % tag] = M.pairwise_dist(new, xyz).numpy()
        gi = PC.knn_point(ns, xyz, new)
        out["pt_%s_knn_point" % tag] = gi.numpy()
        out["pt_%s_index_points" % tag] = PU.index_points(xyz, gi).numpy()
    # get_edge_features_xyz
    B, C, N, k = 2, 12, 128, 6
    x = torch.randn(B, C, N, generator=g)
    pc = torch.randn(B, 3, N, generator=g)
    e_fea, e_xyz = M.get_edge_features_xyz(x, pc, k)
    out.update(efx_x=x.numpy(), efx_pc=pc.numpy(), efx_k=np.int32(k), efx_fea=e_fea.numpy(), efx_xyz=e_xyz.numpy())
    np.savez_compressed(os.path.join(HERE, "graph_util.npz"), **out)
    print("wrote graph_util.npz:", {k_: v.shape for k_, v in out.items() if v.ndim})


if __name__ == "__main__":
    main()
