#!/usr/bin/env python
"""Golden vectors for the pairwise Chamfer evaluation, produced by the reference's OWN function
bodies.  metrics/evaluation_metrics.py is not importable (it imports an un-vendored `StructuralLosses`,
:8-10), so the three pure-torch functions needed (distChamfer :37-49, lgan_mmd_cov :161-173, knn :129-158) are
cut out of the reference source with `ast` and executed unmodified.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_chamfer.py
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SPGAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import spgan_ref as R  # noqa: E402  (synthetic chairs)


def reference_functions(names):
    src = open(os.path.join(REF, "metrics", "evaluation_metrics.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "evaluation_metrics.py", "exec"), ns)
    return [ns[n] for n in names]


def main():
    distChamfer, lgan_mmd_cov, knn = reference_functions(["distChamfer", "lgan_mmd_cov", "knn"])
    rng = np.random.default_rng(70)
    S, Rn, N = 6, 5, 256
    smp = torch.from_numpy(R.synthetic_chairs(rng, S, N))
    ref = torch.from_numpy(R.synthetic_chairs(rng, Rn, N))

    def pairwise(a, b):          # the CD half of _pairwise_EMD_CD_ (evaluation_metrics.py:89-125), batch_size 4
        rows = []
        for i in range(a.shape[0]):
            cds = []
            for r0 in range(0, b.shape[0], 4):
                rb = b[r0:r0 + 4]
                sb = a[i].view(1, -1, 3).expand(rb.size(0), -1, -1).contiguous()
                dl, dr = distChamfer(sb, rb)
                cds.append((dl.mean(dim=1) + dr.mean(dim=1)).view(1, -1))
            rows.append(torch.cat(cds, dim=1))
        return torch.cat(rows, dim=0)

    M_sr, M_ss, M_rr = pairwise(smp, ref), pairwise(smp, smp), pairwise(ref, ref)
    mc = lgan_mmd_cov(M_sr)
    nn = knn(M_ss, M_sr, M_rr, 1, sqrt=False)
    np.savez_compressed(os.path.join(HERE, "chamfer.npz"), sample=smp.numpy(), ref=ref.numpy(), cd_sr=M_sr.numpy(),
                        cd_ss=M_ss.numpy(), cd_rr=M_rr.numpy(), lgan_mmd=np.float32(mc["lgan_mmd"]),
                        lgan_cov=np.float32(mc["lgan_cov"]), lgan_mmd_smp=np.float32(mc["lgan_mmd_smp"]),
                        one_nn_acc=np.float32(nn["acc"]))
    print("wrote chamfer.npz", M_sr.shape, float(mc["lgan_cov"]), float(nn["acc"]))


if __name__ == "__main__":
    main()
