#!/bin/bash
# Sweep one environment knob over the bench (device-resident step time only).  usage: gpu_sweep.sh VAR v1 v2 ...
var=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $var=$v SPGAN_BENCH_MINIMAL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/sweep_${var}_$v.json 2> gpurun_out/sweep_${var}_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/sweep_${var}_$v.json").read().strip().splitlines()[-1])
print("$var=$v", round(d["ms_per_step"],3), "ms/step", round(d["value"],1))
PY
done
