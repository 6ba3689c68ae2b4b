#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_plain.log; tail -3 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|all ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
BENCH_ARGS="--no-cpu-baseline" bash -c 'SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err'
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read()); print('ms/step', d['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'], d['config']['cuda_graph'])"
