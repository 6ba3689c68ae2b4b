#!/usr/bin/env python
"""Golden vectors of the approximate EMD from the REFERENCE'S OWN CUDA kernels (needs a GPU: run through gpurun).

    gpurun -- 'python tests/golden/make_golden_emd.py gpurun_out/emd_reference.npz'    # then copy into tests/golden/

oracle/_ref/emd_ref_harness is metrics/emd/emd_cuda.cu (the auction kernels :23-226 and the host loop
`emd_cuda_forward` :228-282) compiled unmodified from the reference tree (oracle/Makefile target `ref`).  For every
case the harness runs TWICE (the source has benign-looking write races in Bid/GetMax; identical outputs of repeated
runs are recorded as `deterministic`), next to oracle/emd_recipe.c and the product kernel, and the script prints how
the three relate.  Inputs are regenerated from the seed by the tests: only the reference OUTPUTS are stored.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
HARNESS = os.path.join(ROOT, "oracle", "_ref", "emd_ref_harness")

CASES = [(2, 1024, 0.005, 300, 101), (2, 2048, 0.005, 300, 102), (3, 1024, 0.002, 50, 103), (1, 2048, 0.005, 3000, 104)]


def clouds(seed, B, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((B, n, 3), dtype=np.float32), rng.random((B, n, 3), dtype=np.float32)
    from spgan_b200 import synthetic                      # the evaluation's own data: normalised synthetic chairs
    return synthetic.synthetic_chairs(rng, B, n), synthetic.synthetic_chairs(rng, B, n)


def run_reference(a, b, eps, iters):
    B, n, _ = a.shape
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(np.array([B, n], np.int32).tobytes())
            f.write(np.array([eps], np.float32).tobytes())
            f.write(np.array([iters], np.int32).tobytes())
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
            f.write(np.ascontiguousarray(b, np.float32).tobytes())
        subprocess.run([HARNESS, fin, fout], check=True, timeout=600)
        raw = open(fout, "rb").read()
    dist = np.frombuffer(raw[:4 * B * n], np.float32).reshape(B, n).copy()
    ass = np.frombuffer(raw[4 * B * n:], np.int32).reshape(B, n).copy()
    return dist, ass


def main(out_path):
    import torch
    from oracle import emd_ref
    from spgan_b200 import emd as emd_mod
    arrs = {}
    for ci, (B, n, eps, iters, seed) in enumerate(CASES):
        for kind in ("uniform", "chairs"):
            a, b = clouds(seed, B, n, kind)
            d1, s1 = run_reference(a, b, eps, iters)
            d2, s2 = run_reference(a, b, eps, iters)
            det = bool(np.array_equal(s1, s2) and np.array_equal(d1, d2))
            do, so = emd_ref.emd(a, b, eps, iters)
            dk, sk = emd_mod.emdModule()(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), eps, iters)
            dk, sk = dk.cpu().numpy(), sk.cpu().numpy()
            cost = lambda d: np.sqrt(d).sum(1)
            tag = "c%d_%s" % (ci, kind)
            arrs[tag + ".dist"] = d1
            arrs[tag + ".assignment"] = s1.astype(np.int16 if n <= 32767 else np.int32)
            arrs[tag + ".meta"] = np.array([B, n, iters, seed, int(det)], np.int64)
            arrs[tag + ".eps"] = np.float32(eps)
            print("%-12s B=%d n=%d eps=%g iters=%d | reference deterministic: %s | oracle == reference: assignment %s "
                  "(%.4f%% equal), dist %s | kernel == oracle: %s | cost ref %s oracle %s (n*eps = %.2f)" % (
                      tag, B, n, eps, iters, det, np.array_equal(so, s1), 100.0 * (so == s1).mean(),
                      np.array_equal(do, d1), np.array_equal(sk, so) and np.array_equal(dk, do),
                      np.round(cost(d1), 3), np.round(cost(do), 3), n * eps), flush=True)
    np.savez_compressed(out_path, **arrs)
    print("wrote", out_path, os.path.getsize(out_path) // 1024, "KB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "emd_reference.npz"))
