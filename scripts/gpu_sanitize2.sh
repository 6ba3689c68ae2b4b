#!/bin/bash
# compute-sanitizer over the small-shape pass of every hand-synchronised kernel (no bench, no full suite)
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_plain.log; tail -4 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|all ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
