#!/bin/bash
# Round-2 evidence: ncu launch list of one step, ncu --set full of the dominant kernels, bench line + reference arm.
mkdir -p gpurun_out
SPGAN_BENCH_MINIMAL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/r2_ts_fc2 python scripts/prof_fused_shape.py 131072 1024 256 pro > gpurun_out/ncu_a.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/r2_ts_e128k64 python scripts/prof_fused_shape.py 1310720 128 64 pro > gpurun_out/ncu_b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_wg_kernel -s 2 -c 1 -f -o gpurun_out/r2_wg_fc2 python scripts/prof_wgrad_shape.py 1024 256 131072 > gpurun_out/ncu_c.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_tc_k1280 python scripts/prof_gemm_shape.py 131072 128 1280 > gpurun_out/ncu_d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"knn_group_fast|bn_pool_partial|colmap4_kernel|bn_softmax_mul_k|pairwise_chamfer" -c 8 -f -o gpurun_out/r2_ops python scripts/prof_ops.py > gpurun_out/ncu_ops.log 2>&1
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
ls -la gpurun_out | tail -15; tail -c 600 gpurun_out/bench_r2_reference.json
