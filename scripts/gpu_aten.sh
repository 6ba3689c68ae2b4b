#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/prof_aten.py > gpurun_out/prof_aten.log 2>&1; cat gpurun_out/prof_aten.log | tail -70
