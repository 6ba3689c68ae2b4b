#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3; nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_modules.py -q -x > gpurun_out/pytest_fullsize.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fullsize.log
tail -40 gpurun_out/pytest_fullsize.log
