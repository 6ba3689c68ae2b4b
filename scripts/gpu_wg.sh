#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_fused.py tests/test_gpu_gemm_tc.py -q -x > gpurun_out/pytest_wg.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wg.log
grep -E "^E|passed|failed|rc=" gpurun_out/pytest_wg.log | tail -12
timeout 300 python - > gpurun_out/bench_wg.log 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
import torch, spgan_b200 as pkg
ops = pkg.ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for Mo, No, K in [(1024, 256, 131072), (256, 128, 131072), (128, 1280, 131072), (128, 64, 1310720), (128, 64, 131072), (64, 32, 1310720), (64, 640, 131072), (64, 128, 1310720)]:
    A = torch.randn(K, Mo, device="cuda"); B = torch.randn(K, No, device="cuda"); out = torch.empty(Mo, No, device="cuda")
    row = []
    for eng in (1, 3):
        ts = []
        for it in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm_raw(A, B, None, True, False, out=out, engine=eng); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(sorted(ts)[3])
    fl = 2.0 * Mo * No * K / 1e9; by = 4.0 * K * (Mo + No) / 1e6
    print("wgrad Mo=%-5d No=%-5d K=%-8d  tf32x3(tn) %.3f ms %6.1f TF | wg %.3f ms %6.1f TF  %5.2f TB/s" % (Mo, No, K, row[0], fl / row[0], row[1], fl / row[1], by / row[1] / 1e3), flush=True)
PY
cat gpurun_out/bench_wg.log
