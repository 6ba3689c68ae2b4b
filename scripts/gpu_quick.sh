#!/bin/bash
# Quick GPU visit: parity tests, bench line, GEMM/kNN microbenches.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 120 python scripts/bench_knn.py > gpurun_out/bench_knn.log 2>&1
timeout 120 python scripts/bench_chamfer.py > gpurun_out/bench_chamfer.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
