#!/bin/bash
# full GPU suite + bench with per-shape tables (after a kernel change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline_all_gemm"]["achieved"], d["submetrics"])
PY
grep -E "^(NT|NN|TN) " gpurun_out/bench_b.err | head -16
