"""BASELINE configs[4]: pairwise CD (and a bounded EMD block) of S x R synthetic chair clouds, N = 2048 points, the
S x R matrix row-sharded over the ranks of one box (metrics/evaluation_metrics.py:89-125, Common/GAN_metrics.py:658-747).

    python scripts/bench_pairwise.py --clouds 5000                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_pairwise.py --clouds 5000 --emd-clouds 256

Rank 0 prints one JSON line: cloud pairs per second (device time, max over ranks, all_gather included), the derived
statistics (MMD / COV / 1-NN accuracy) of the gathered matrix, and the fp32 issue-rate fraction of the Chamfer kernel.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spgan_b200 as pkg  # noqa: E402
from spgan_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=5000)
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--emd-clouds", type=int, default=0, help="S = R of the EMD block (0 = skip)")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = R = args.clouds
    N = args.points
    # two sets of synthetic chairs, seeds 123 / 124; every rank generates the same sets
    gen = lambda seed, cnt: torch.from_numpy(synthetic.synthetic_chairs(np.random.default_rng(seed), cnt, N))
    a, b = gen(123, S).to(dev), gen(124, R).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        best = None
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best = float(ms) if best is None else min(best, float(ms))
        return best, out

    pkg.pairwise_CD(a[:64], b[:64])                                   # warm-up (kernel load, NCCL channels)
    ms_cd, cd = timed(lambda: pkg.pairwise_CD(a, b), args.reps)
    line = {"bench": "pairwise_CD", "config": "configs[4]: %d x %d clouds, N=%d, rows sharded over %d GPU(s)" % (S, R, N, world),
            "n_gpus": world, "ms": ms_cd, "cloud_pairs_per_s": S * R / (ms_cd / 1e3),
            "fp32_inst_per_s": 8.0 * S * R * N * N / (ms_cd / 1e3),
            "frac_of_fp32_issue_peak": 8.0 * S * R * N * N / (ms_cd / 1e3) / (world * 148 * 128 * 1.965e9)}
    if rank == 0:
        line["mmd_cov"] = pkg.lgan_mmd_cov(cd)
        line["checksum"] = float(cd.double().sum())
    if args.emd_clouds:
        E = args.emd_clouds
        ms_emd, emd = timed(lambda: pkg.pairwise_EMD(a[:E], b[:E], eps=0.005, iters=300), 1)
        line["emd"] = {"clouds": [E, E], "eps": 0.005, "iters": 300, "ms": ms_emd, "cloud_pairs_per_s": E * E / (ms_emd / 1e3),
                       "est_5000x5000_s": 25e6 / (E * E / (ms_emd / 1e3))}
        if rank == 0:
            line["emd"]["mmd_cov"] = pkg.lgan_mmd_cov(emd)
            line["emd"]["checksum"] = float(emd.double().sum())
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
