#!/bin/bash
# final evidence of a build whose kernels were already captured: full GPU suite, launch list, bench line + tables, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
SPGAN_BENCH_MINIMAL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
[ "$SKIP_REF" = "1" ] || timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
tail -c 300 gpurun_out/bench_r2_reference.json; echo; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_all_gemm']['achieved'], d['gpu_launches']/d['steps'])"
