#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/ts_fc2 python scripts/prof_fused_shape.py 131072 1024 256 > gpurun_out/ncu_ts_fc2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/ts_k64 python scripts/prof_fused_shape.py 131072 128 64 > gpurun_out/ncu_ts_k64.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/ts_k128 python scripts/prof_fused_shape.py 131072 256 128 > gpurun_out/ncu_ts_k128.log 2>&1
tail -3 gpurun_out/ncu_ts_*.log
