"""Where do the tensor-core kNN lists differ from the CUDA-core kernel's?  python scripts/diag_knn_tc.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
C, N, k = 64, 2048, 10
rng = np.random.default_rng(B * 7 + C + N)
x = rng.standard_normal((B, C, N)).astype(np.float32)
xb = torch.from_numpy(x).cuda()
rows = xb.permute(0, 2, 1).contiguous().view(B * N, C)
want = pkg.ops.knn_indices(xb, k).cpu().numpy()
for rep in range(int(os.environ.get("REPS", "3"))):
    got = pkg.ops.knn_indices_rows(rows, B, N, k).cpu().numpy()
    ws = pkg.ops.LAST_KNN_WORKSPACE
    bad = np.argwhere((got != want).any(axis=2))
    print("rep", rep, "queries differing:", len(bad), "fallbacks", int(ws[1]) if ws is not None else None)
    if len(bad):
        q = bad[:, 0] * N + bad[:, 1]
        mt = q // 128
        print("  tiles:", np.unique(mt)[:40], "n tiles", len(np.unique(mt)))
        print("  cta (mt % 148):", np.unique(mt % 148)[:40])
        print("  iteration (mt // 148):", np.bincount(mt // 148))
        print("  row in tile:", np.bincount(q % 128, minlength=128))
        b, n = bad[0]
        print("  first:", b, n, "got", got[b, n], "want", want[b, n])
        d = ((x[b][:, n:n + 1] - x[b]) ** 2).sum(0)
        print("   dist got", d[got[b, n]], "\n   dist want", d[want[b, n]])
