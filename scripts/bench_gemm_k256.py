import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg
ops = pkg.ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
print("SPGAN_TSK256 =", os.environ.get("SPGAN_TSK256", "1"))
for M, N, K, tb in [(131072, 128, 256, True), (131072, 128, 256, False), (131072, 64, 256, True), (1310720, 128, 256, True)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") if tb else torch.randn(K, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ref = (A[:512].double() @ (B.t() if tb else B).double()).float()
    ts = []
    for it in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_raw(A, B, None, False, tb, out=out, engine=3); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[4]
    print("M=%8d N=%4d K=%4d tb=%d: %.3f ms %.1f TF  err %.2e" % (M, N, K, tb, t, 2.0*M*N*K/t/1e9, float((out[:512]-ref).abs().max()/ref.abs().max())), flush=True)
