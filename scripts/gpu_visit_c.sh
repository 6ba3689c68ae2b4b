#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_fused.py -x -q > gpurun_out/pytest_edge_fused.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_edge_fused.log
tail -15 gpurun_out/pytest_edge_fused.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/r2c_ts_fc2 python scripts/prof_fused_shape.py 131072 1024 256 pro > gpurun_out/ncu_a.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2c_tc_k1024 python scripts/prof_gemm_shape.py 131072 256 1024 > gpurun_out/ncu_b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_wg_kernel -s 2 -c 1 -f -o gpurun_out/r2c_wg_fc2 python scripts/prof_wgrad_shape.py 1024 256 131072 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/r2c*
