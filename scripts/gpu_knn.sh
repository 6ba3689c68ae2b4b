#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py -q -x > gpurun_out/pytest_knn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_knn.log
grep -E "^E|passed|failed|rc=" gpurun_out/pytest_knn.log | tail -12
cat > /tmp/bk.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch, spgan_b200 as pkg
ops = pkg.ops
for B, C, N, k in [(64, 64, 2048, 10), (64, 128, 2048, 10)]:
    x = torch.randn(B, C, N, device="cuda")
    rows = x.permute(0, 2, 1).contiguous().view(B * N, C)
    def med(fn):
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return sorted(ts)[3]
    t0 = med(lambda: ops.knn_indices(x, k))
    t1 = med(lambda: ops.knn_indices_rows(rows, B, N, k))
    print("B=%d C=%d N=%d k=%d: CUDA-core exact %.3f ms | tensor-core filter + refine %.3f ms (fallbacks %d)" % (B, C, N, k, t0, t1, int(ops.LAST_KNN_WORKSPACE[1])), flush=True)
PY
timeout 200 python /tmp/bk.py > gpurun_out/bench_knn_tc.log 2>&1
cat gpurun_out/bench_knn_tc.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_|sqnorm" --csv --log-file gpurun_out/knn_tc_launches.csv python /tmp/bk.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/knn_tc_launches.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:][-12:]:
    print(r[ki][:60], r[vi])
PY
