#!/bin/bash
# Lean GPU-box visit: parity tests + smoke + one bench line (+ the stock-ATen census).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E|passed|failed|rc=" gpurun_out/pytest_gpu.log | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read()); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'], d['roofline']['achieved'], d['roofline_all_gemm']['achieved'], d['kernel_share'])"
grep -v "^BW\|^[NT][NT] " gpurun_out/bench.err | tail -5
timeout 300 python scripts/prof_aten.py > gpurun_out/prof_aten.log 2>&1; grep -A12 "stock kernels" gpurun_out/prof_aten.log
