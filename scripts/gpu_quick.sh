#!/bin/bash
# Lean GPU-box visit: parity tests + smoke + one bench line.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; grep -v "^BW\|^N\|^T" gpurun_out/bench.err | tail -5
