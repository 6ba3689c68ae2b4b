#!/bin/bash
bash scripts/gpu_wg.sh 2>&1 | tail -12
bash scripts/gpu_quick.sh
