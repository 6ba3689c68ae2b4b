"""CPU: the Chamfer oracle (oracle/chamfer_ref.py) against the golden vectors produced by the reference's own
function bodies (tests/golden/make_golden_chamfer.py), and the host-side MMD / COV / 1-NN reductions of the
product (numpy) against the oracle's torch restatement."""
import numpy as np
import torch

from conftest import assert_rel, golden
from oracle import chamfer_ref as C


def test_pairwise_cd_matches_reference_golden():
    g = golden("chamfer")
    smp, ref = torch.from_numpy(g["sample"]), torch.from_numpy(g["ref"])
    assert np.array_equal(C.pairwise_cd(smp, ref).numpy(), g["cd_sr"])          # same fp32 arithmetic: bit-exact
    assert np.array_equal(C.pairwise_cd(smp, smp).numpy(), g["cd_ss"])
    # the direct-difference float64 form (what the CUDA kernel computes in fp32) agrees to 1e-3 relative
    assert_rel(C.pairwise_cd_exact(g["sample"], g["ref"]), g["cd_sr"], 1e-3, "exact vs expanded")
    mc = C.lgan_mmd_cov(torch.from_numpy(g["cd_sr"]))
    assert abs(mc["lgan_mmd"] - float(g["lgan_mmd"])) < 1e-7 and abs(mc["lgan_cov"] - float(g["lgan_cov"])) < 1e-7
    acc = C.one_nn_accuracy(torch.from_numpy(g["cd_ss"]), torch.from_numpy(g["cd_sr"]), torch.from_numpy(g["cd_rr"]))
    assert abs(acc - float(g["one_nn_acc"])) < 1e-7


def test_host_reductions_of_the_product_match_the_oracle():
    import importlib.util
    import os
    # metrics.py imports the CUDA binding lazily through ops; the reductions themselves are numpy
    import spgan_b200
    g = golden("chamfer")
    mc = spgan_b200.lgan_mmd_cov(g["cd_sr"])
    for k in ("lgan_mmd", "lgan_cov", "lgan_mmd_smp"):
        assert abs(mc[k] - float(g[k])) < 1e-6, k
    acc = spgan_b200.one_nn_accuracy(g["cd_ss"], g["cd_sr"], g["cd_rr"])
    assert abs(acc - float(g["one_nn_acc"])) < 1e-7
