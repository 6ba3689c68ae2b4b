#!/bin/bash
bash scripts/gpu_knn.sh 2>&1 | grep -v "sqnorm\|presplit" | tail -12
bash scripts/gpu_quick.sh
