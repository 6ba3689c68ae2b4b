#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm_fused.py tests/test_gpu_gemm_tc.py tests/test_gpu_gemm_f16s.py -q -x > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fused.log
tail -30 gpurun_out/pytest_fused.log
timeout 200 python scripts/bench_gemm_fused.py > gpurun_out/bench_gemm_fused.log 2>&1
cat gpurun_out/bench_gemm_fused.log
timeout 200 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1
cat gpurun_out/bench_gemm.log
