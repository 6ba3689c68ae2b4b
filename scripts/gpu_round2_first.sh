#!/bin/bash
# First GPU visit of round 2: run everything that round 1 drafted after its GPU budget was spent.  Every step is
# under its own `timeout` (the warp-specialised kNN kernel synchronises through named barriers: a protocol bug would
# hang, not fail).  Results land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "default tier rc=$?" >> gpurun_out/pytest_gpu.log
# 1. fp16x3 engine: accuracy, then speed on the big shapes and on the whole step
SPGAN_TEST_ENGINE3=1 timeout 200 python -m pytest tests/test_gpu_gemm_f16s.py -q > gpurun_out/pytest_engine3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_engine3.log
timeout 120 python - > gpurun_out/bench_engine3.log 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
import torch, spgan_b200 as pkg
ops = pkg.ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for M, N, K in [(131072, 1024, 256), (131072, 256, 1024), (131072, 128, 1280), (131072, 256, 128), (1310720, 128, 64)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); out = torch.empty(M, N, device="cuda")
    for eng in (1, 3):
        ts = []
        for it in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm_raw(A, B, None, False, True, out=out, engine=eng); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[2]
        print("M=%d N=%d K=%d engine %d: %.3f ms  %.1f TFLOP/s" % (M, N, K, eng, t, 2.0 * M * N * K / t / 1e9))
PY
SPGAN_GEMM_ENGINE=3 timeout 300 python -m pytest tests/test_gpu_modules.py -q > gpurun_out/pytest_modules_engine3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_modules_engine3.log
timeout 300 python bench.py --engine 3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_engine3.json 2> gpurun_out/bench_engine3.err
# 2. warp-specialised kNN kernel: bit-exactness, then the microbench with and without it
SPGAN_KNN_WS=1 timeout 200 python -m pytest tests/test_gpu_knn_ws.py -q -x > gpurun_out/pytest_knn_ws.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_knn_ws.log
timeout 100 python scripts/bench_knn.py > gpurun_out/bench_knn_default.log 2>&1
SPGAN_KNN_WS=1 timeout 100 python scripts/bench_knn.py > gpurun_out/bench_knn_ws.log 2>&1
# 3. evaluation path: pairwise EMD composition, EMD kernel speed
SPGAN_TEST_PAIRWISE_EMD=1 timeout 200 python -m pytest tests/test_gpu_chamfer.py -q -k pairwise_emd > gpurun_out/pytest_pairwise_emd.log 2>&1
timeout 200 python scripts/bench_emd.py > gpurun_out/bench_emd.log 2>&1
tail -3 gpurun_out/pytest_*.log; cat gpurun_out/bench_engine3.log gpurun_out/bench_knn_default.log gpurun_out/bench_knn_ws.log gpurun_out/bench_emd.log
