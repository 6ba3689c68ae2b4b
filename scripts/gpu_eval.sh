#!/bin/bash
# Evaluation path on N GPUs: EMD pin against the reference's own kernels, sharded pairwise tests, config-5 bench.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python tests/golden/make_golden_emd.py gpurun_out/emd_reference.npz > gpurun_out/make_golden_emd.log 2>&1; echo "rc=$?" >> gpurun_out/make_golden_emd.log
cat gpurun_out/make_golden_emd.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -k "two_rank" > gpurun_out/pytest_two_rank.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_two_rank.log
tail -5 gpurun_out/pytest_two_rank.log
timeout 300 python scripts/bench_pairwise.py --clouds 5000 --emd-clouds 128 > gpurun_out/pairwise_n1.json 2> gpurun_out/pairwise_n1.err; tail -c 1200 gpurun_out/pairwise_n1.json
for n in 2 4 8; do
  if [ $NG -ge $n ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/bench_pairwise.py --clouds 5000 --emd-clouds 256 > gpurun_out/pairwise_n$n.json 2> gpurun_out/pairwise_n$n.err
    tail -c 1200 gpurun_out/pairwise_n$n.json
  fi
done
