#!/bin/bash
# Round-2 second session, first visit: launch list of the current step, fresh ncu capture of the fc2 GEMM, bench tables.
mkdir -p gpurun_out
SPGAN_BENCH_MINIMAL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_ts_kernel -s 2 -c 1 -f -o gpurun_out/r2b_ts_fc2 python scripts/prof_fused_shape.py 131072 1024 256 pro > gpurun_out/ncu_a.log 2>&1
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
ls -la gpurun_out | tail -5; tail -c 300 gpurun_out/bench_a.json
