"""Diagnostic (GPU): decode the operand mapping of the MN-major tcgen05 wgrad kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import spgan_b200 as pkg
ops = pkg.ops
Mo, No, K = 128, 64, 4096
for k0 in (0, 1, 5, 8, 31, 32):
    A = torch.zeros(K, Mo); B = torch.zeros(K, No)
    A[k0] = torch.arange(1, Mo + 1).float(); B[k0] = torch.arange(1, No + 1).float()
    out = ops.gemm_raw(A.cuda(), B.cuda(), None, True, False, engine=1).cpu()
    st = int(ops.LAST_TC_WORKSPACE.view(torch.int32)[0])
    ref = A.t() @ B
    nz = (out != 0).sum().item()
    print("k0=%d status=%d nonzero=%d/%d equal=%s" % (k0, st, nz, Mo * No, torch.equal(out, ref)))
    if not torch.equal(out, ref):
        # decode: out[m,n] = f(m')*g(n') -> find m', n' for a few entries
        for (m, n) in [(0, 0), (1, 0), (0, 1), (4, 0), (0, 4), (32, 0), (0, 32), (33, 5), (127, 63)]:
            v = out[m, n].item()
            print("   out[%d,%d]=%g expected %g" % (m, n, v, ref[m, n].item()))
# all-k test with distinct k weights
A = torch.zeros(K, Mo); B = torch.zeros(K, No)
A[:, 0] = 1.0; B[:, 0] = torch.arange(K).float() % 64
out = ops.gemm_raw(A.cuda(), B.cuda(), None, True, False, engine=1).cpu()
print("sum over k: got", out[0, 0].item(), "expected", (A.t() @ B)[0, 0].item())
