#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_fused.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fused.log
tail -5 gpurun_out/pytest_fused.log
for i in 1 2 3; do
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_d$i.json 2> gpurun_out/bench_d$i.err; echo "bench $i rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_d$i.json'))
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline_all_gemm"]["achieved"])
except Exception as e:
    print("bench failed", e)
PY
done
