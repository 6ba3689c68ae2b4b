#!/bin/bash
# A/B of library variants on the same box: scripts/gpu_ab.sh <script.py> variantA variantB ...
mkdir -p gpurun_out
S=$1; shift
cp sp-gan_b200/libspgan_b200.so /tmp/lib_current.so
for v in "$@"; do
  cp sp-gan_b200/variants/$v.so sp-gan_b200/libspgan_b200.so
  echo "=== $v"
  timeout 300 python $S 2>&1 | tee gpurun_out/ab_$v.log | tail -14
done
cp /tmp/lib_current.so sp-gan_b200/libspgan_b200.so
