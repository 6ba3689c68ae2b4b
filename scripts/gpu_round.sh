#!/bin/bash
# One full GPU-box visit: parity tests, bench line, microbenches, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 120 python scripts/bench_knn.py > gpurun_out/bench_knn.log 2>&1
timeout 200 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1
timeout 120 python scripts/bench_chamfer.py > gpurun_out/bench_chamfer.log 2>&1
SPGAN_BENCH_MINIMAL=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"knn_group_fast|bn_pool_partial|colmap4_kernel|bn_softmax_mul_k|pairwise_chamfer" -c 8 -f -o gpurun_out/ops_full python scripts/prof_ops.py > gpurun_out/ncu_ops.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/gemm_fc2_full python scripts/prof_gemm_shape.py 131072 1024 256 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
