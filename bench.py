#!/usr/bin/env python
"""bench.py -- point-clouds/sec of the full G+D WGAN-GP training step at N=2048, B=64 per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

One "step" = one iteration of Generation/model.py:239-279 with gan='wgan' plus
GradientPenalty(10) (BASELINE.json configs[2]: "Full G+D WGAN-GP training step, Chair synthetic,
N=2048 B=64").  Rank 0 prints ONE JSON line.  `value` is timed with the inputs already resident in
HBM; `e2e` goes through the same public API with pinned-host inputs copied every step and the losses
read back.  Batches shard across ranks (weak scaling, 64 clouds per GPU, BN statistics per replica as
in the reference's DataParallel) with one NCCL all-reduce of a flat gradient buffer per phase.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point-clouds/sec per G+D WGAN-GP step at N=2048 B=64"
UNIT = "clouds/s"
N_POINTS, BATCH, NZ = 2048, 64, 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="spgan_b200", choices=["spgan_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU port timing (N=1 only runs it)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="clouds per step of the CPU sample (0 = the full batch)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed steps of the CPU sample (after one warm-up)")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--engine", type=int, default=None, help="spgan_gemm engine (0 fp32 CUDA cores, 1 tcgen05)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying "
                                                             "the captured CUDA graph of the step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------
def cpu_step_time(batch, points, steps, warmup, seed=123):
    import numpy as np
    import torch
    from oracle import spgan_ref as R
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    o = R.default_opts(np=points)
    st = R.TrainState(R.synth_state(R.generator_spec(o), 1), R.synth_state(R.discriminator_spec(o), 2), o)
    rng = np.random.default_rng(seed)
    from spgan_b200.synthetic import sphere_template
    ball, _ = sphere_template(points)
    x = torch.from_numpy(np.tile(ball[None], (batch, 1, 1)))
    times = []
    for it in range(warmup + steps):
        real = torch.from_numpy(R.synthetic_chairs(rng, batch, points)).transpose(2, 1)
        z_d = torch.from_numpy(R.latent_noise(rng, batch, points, o.nz))
        z_g = torch.from_numpy(R.latent_noise(rng, batch, points, o.nz))
        alpha = torch.rand(batch, 1, 1)
        t0 = time.perf_counter()
        R.wgan_gp_train_step(st, x, z_d, z_g, real, alpha)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args, real_stdout):
    """The reference arm: the full configs[2] batch (B clouds) every step, for the step count asked for
    (about 13 s per step on 16 host cores: 25 steps fit the driver's window).  Only if the projected run
    would pass SPGAN_REF_BUDGET_S (default 1500 s) is the per-step batch cut, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("SPGAN_REF_BUDGET_S", "1500"))
    b = args.batch
    t_probe, cores = cpu_step_time(b, args.points, 1, 0)          # also serves as a first warm-up of the host
    n_more = args.steps + max(0, args.warmup - 1)
    if n_more * t_probe > budget:
        b = max(2, int(b * budget / (n_more * t_probe)))
    if b == args.batch:
        t, cores = cpu_step_time(b, args.points, args.steps, max(0, args.warmup - 1))
    else:
        t, cores = cpu_step_time(b, args.points, args.steps, args.warmup)
    value = b / t
    sample = "each step = one full WGAN-GP step on %d of the %d clouds (N=%d)" % (b, args.batch, args.points)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: full G+D WGAN-GP step, Chair synthetic, N=%d B=%d" % (args.points, args.batch),
                       "parallelism": "cpu", "l2": "n/a (CPU)", "batch_per_step": b, "same_config": b == args.batch},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=real_stdout, flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def gemm_flops(args_):
    return 2.0 * args_[2] * args_[3] * args_[4]


def _claim_stdout():
    """Rank 0 must print exactly ONE line on stdout: park the real stdout and point fd 1 at stderr, so that
    anything a library writes there (NCCL prints its version banner on stdout) cannot end up next to the JSON."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse()
    real_stdout = _claim_stdout()
    try:
        if args.impl == "reference":
            run_reference(args, real_stdout)
        else:
            run_gpu(args, real_stdout)
    finally:
        real_stdout.flush()
        sys.stdout.flush()
        sys.stderr.flush()
    # Normal interpreter exit (the process group was destroyed inside run_gpu after the captured graph was dropped).
    # A watchdog covers the one thing that cannot be allowed: a teardown that blocks after the JSON line is out.
    def _watchdog():
        time.sleep(60.0)
        sys.stderr.write("bench.py: teardown still running after 60 s, leaving\n")
        sys.stderr.flush()
        os._exit(0)
    threading.Thread(target=_watchdog, daemon=True).start()


def run_gpu(args, real_stdout):

    import numpy as np
    import torch
    import torch.distributed as dist
    import spgan_b200 as pkg
    from spgan_b200 import synthetic
    from spgan_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- spgan_b200 has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.engine is not None:
        pkg.ops.GEMM_ENGINE = args.engine
    B, N = args.batch, args.points

    # ---- model: same random-init architecture on every rank (seed 123 as Generation/model.py:38-41)
    torch.manual_seed(123)
    opts = type("Opts", (), dict(np=N, nk=20, nz=NZ, softmax=True, off=False, attn=False, use_head=False,
                                 eql=False, z_norm=False, small_d=False))()
    G, D = pkg.Generator(opts).to(dev).train(), pkg.Discriminator(opts).to(dev).train()
    trainer = pkg.WGANGPTrainer(G, D, lambda_gp=10.0, gamma=1.0)

    # ---- synthetic inputs: different clouds / latents per rank, pinned on the host
    rng = np.random.default_rng(123 + rank)
    ball, ball_src = synthetic.sphere_template(N)
    x = torch.from_numpy(np.tile(ball[None], (B, 1, 1))).to(dev)           # constant of training (model.py:231)
    n_pool = 4
    host = []
    for _ in range(n_pool):
        host.append(dict(real=torch.from_numpy(synthetic.synthetic_chairs(rng, B, N)).pin_memory(),
                         z_d=torch.from_numpy(synthetic.latent_vectors(rng, B, NZ)).pin_memory(),
                         z_g=torch.from_numpy(synthetic.latent_vectors(rng, B, NZ)).pin_memory(),
                         alpha=torch.rand(B, 1, 1).pin_memory()))
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]

    def eager_step(d):
        real = d["real"].transpose(2, 1)                                   # strided view, as model.py:249
        return trainer.step(x, d["z_d"].expand(B, N, NZ), d["z_g"].expand(B, N, NZ), real, alpha=d["alpha"])

    # The whole step (both phases, ~700 launches, the gradient all-reduces, Adam) is captured once into a CUDA
    # graph over static input buffers and replayed: the Python/ctypes enqueue cost (~40 ms/step, as long as the
    # GPU work itself) disappears from the critical path.  Same kernels, same arithmetic.
    graph_note = "disabled (--no-graph)"
    minimal = os.environ.get("SPGAN_BENCH_MINIMAL") == "1"      # under ncu: skip the extra passes
    if not args.no_graph and not minimal:
        try:
            d0 = resident[0]
            trainer.capture(x, d0["z_d"].expand(B, N, NZ), d0["z_g"].expand(B, N, NZ), d0["real"].transpose(2, 1),
                            d0["alpha"], warmup=2)
            graph_note = "whole step replayed from one CUDA graph (%d kernel launches recorded)" % trainer.graph_launches
        except Exception as exc:                                           # noqa: BLE001 -- report and fall back
            graph_note = "capture failed, eager enqueue: %s" % (str(exc).splitlines()[0][:200],)
            trainer._graph = None
            torch.cuda.synchronize()
    use_graph = trainer._graph is not None

    def step_on(d):
        if use_graph:
            return trainer.replay(z_d=d["z_d"], z_g=d["z_g"], real=d["real"].transpose(2, 1), alpha=d["alpha"])
        return eager_step(d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        host_ms.append(1e3 * (time.perf_counter() - t0) / steps)     # host enqueue time per step (no sync inside)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    L = lib()
    for i in range(args.warmup):
        step_on(resident[i % n_pool])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.launches
    ms = timed(lambda i: step_on(resident[i % n_pool]), args.steps)
    launches = trainer.graph_launches * args.steps if use_graph else L.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: pinned host -> device every step, losses read back every step
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    losses = []

    # The three step results go device -> pinned host right behind the step, on the same stream (the graph's output
    # scalars are overwritten by the next replay); the host READS step i's numbers while step i+1 runs, so the
    # read-back costs the GPU no idle time.  Every step's inputs are copied and every step's results are read inside
    # the timed region; the last step's are awaited before it closes.
    loss_host = torch.empty((args.steps, 3), dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(args.steps)]

    def read_back(i):
        loss_ev[i].synchronize()
        losses.append(loss_host[i].tolist())

    def e2e_step(i):
        if use_graph:                           # pinned host -> the graph's static input buffers
            out = step_on(host[i % n_pool])
        else:
            out = step_on({k: v.to(dev, non_blocking=True) for k, v in host[i % n_pool].items()})
        if os.environ.get("SPGAN_BENCH_SYNC_READ") == "1":                    # A/B: blocking read-back every step
            losses.append([float(t) for t in out])
            return
        for j, t in enumerate(out):                                          # D2H of the three step results
            loss_host[i, j].copy_(t, non_blocking=True)
        loss_ev[i].record()
        if i > 0:
            read_back(i - 1)
        if i == args.steps - 1:
            read_back(i)

    if not minimal:
        # warm the end-to-end path itself (pinned staging, events, the async D2H route) before timing it
        for i in range(min(args.warmup, 3)):
            out = step_on(host[i % n_pool])
            for j, t in enumerate(out):
                loss_host[0, j].copy_(t, non_blocking=True)
            loss_ev[0].record()
            loss_ev[0].synchronize()
    ms_e2e = timed(e2e_step, args.steps) if not minimal else float("nan")
    e2e = {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e / args.steps}

    # ---- secondary end-to-end figure, the way the reference's train.py feeds the step: the latent is
    # generated on the host as one vector per cloud and TILED over the N points (model.py:122-131: 67 MB per
    # generator forward, pageable memory), shipped with .cuda() twice per step (model.py:246,270), and the three
    # losses are read with .item() (model.py:282-286).  The compute is the same graph replay.
    e2e_train_py = None
    if not minimal:
        rng_t = np.random.default_rng(7 + rank)

        def noise_generator():                                            # model.py:122-131
            nz_ = rng_t.normal(0, 0.2, (B, 1, NZ)).astype(np.float32)
            return torch.from_numpy(np.tile(nz_, (1, N, 1)))

        def train_py_step(i):
            zd = noise_generator().to(dev)                                # pageable -> device, synchronous
            zg = noise_generator().to(dev)
            d = host[i % n_pool]
            if use_graph:
                out = trainer.replay(z_d=zd, z_g=zg, real=d["real"].transpose(2, 1), alpha=d["alpha"])
            else:
                out = trainer.step(x, zd, zg, d["real"].to(dev).transpose(2, 1), alpha=d["alpha"].to(dev))
            losses.append([t.item() for t in out])

        n_tp = max(3, min(args.steps, 10))
        ms_tp = timed(train_py_step, n_tp)
        e2e_train_py = {"value": world * B * n_tp / (ms_tp / 1e3), "unit": UNIT, "ms_per_step": ms_tp / n_tp,
                        "steps": n_tp, "h2d_bytes_per_step": 2 * B * N * NZ * 4 + h2d - 2 * B * NZ * 4,
                        "d2h_bytes_per_step": 12,
                        "note": "host-side tiled latent (np.tile) + pageable 67 MB H2D twice per step + .item() reads, "
                                "as Generation/model.py:122-131,246,270,282-286 does"}

    # ---- per-kernel device times of one more step (CUDA events around every C-ABI launch)
    roofline, kernel_share, roofline_all, roofline_knn = None, None, None, None
    if not minimal:
        # every rank runs the instrumented step (it contains the gradient all-reduces); rank 0 records it
        if rank == 0:
            L.profile = []
        eager_step(resident[0])                 # eager: every C-ABI call bracketed by CUDA events
        torch.cuda.synchronize()
    if rank == 0 and not minimal:
        prof, L.profile = L.profile, None
        agg = {}
        for name, ia, s, e in prof:
            t = s.elapsed_time(e)
            a = agg.setdefault(name, [0.0, 0, 0.0])
            a[0] += t
            a[1] += 1
            if name == "spgan_gemm":
                a[2] += gemm_flops(ia)
            elif name == "spgan_gemm_fused":                         # (transB, M, N, K, ...)
                a[2] += 2.0 * ia[1] * ia[2] * ia[3]
            elif name == "spgan_gemm_wgrad_fused":                   # (Mo, No, K, ...)
                a[2] += 2.0 * ia[0] * ia[1] * ia[2]
        total = sum(a[0] for a in agg.values())
        if os.environ.get("SPGAN_BENCH_GEMM_TABLE") == "1":        # diagnostic: GEMM time by shape (stderr)
            byshape = {}
            for name, ia, s_, e_ in prof:
                if name == "spgan_gemm":
                    key = (bool(ia[0]), bool(ia[1]), ia[2], ia[3], ia[4])              # transA, transB, M, N, K
                    b = byshape.setdefault(key, [0.0, 0])
                    b[0] += s_.elapsed_time(e_)
                    b[1] += 1
            for (ta, tb, M_, N_, K_), b in sorted(byshape.items(), key=lambda kv: -kv[1][0])[:25]:
                label = ("T" if ta else "N") + ("T" if tb else "N") + " M=%d N=%d K=%d" % (M_, N_, K_)
                sys.stderr.write("%-40s x%-3d %8.3f ms  %6.1f TFLOP/s\n" % (
                    label, b[1], b[0], 2.0 * M_ * N_ * K_ * b[1] / b[0] / 1e9))
        if os.environ.get("SPGAN_BENCH_BW_TABLE") == "1":          # diagnostic: HBM-bound entry points (stderr)
            # algorithmic bytes per call from the integer arguments (R or P, C, k ...) of each C-ABI call
            def nbytes(name, ia):
                if name in ("spgan_colstats", "spgan_colsum"):
                    return 4.0 * ia[0] * ia[1]                                   # R, C: one read
                if name == "spgan_norm_apply":
                    return 8.0 * ia[0] * ia[1]                                   # read x, write y
                if name == "spgan_norm_bwd_reduce":
                    return 8.0 * ia[0] * ia[1]                                   # read g, x
                if name == "spgan_norm_bwd_apply":
                    return 12.0 * ia[0] * ia[1]                                  # read g, x, write dx
                if name in ("spgan_lrelu", "spgan_tanh"):
                    return 8.0 * ia[-2] if len(ia) >= 2 else None                # n: read + write
                if name == "spgan_lrelu_bwd":
                    return 12.0 * ia[-2] if len(ia) >= 2 else None
                if name == "spgan_bn_softmax_mul_k":
                    return 16.0 * ia[0] * ia[1] * ia[2]                          # P, k, C: 2 reads, 2 writes
                if name == "spgan_bn_softmax_mul_k_bwd":
                    return 20.0 * ia[0] * ia[1] * ia[2]                          # 3 reads, 2 writes
                if name == "spgan_edge_combine":
                    return 4.0 * ia[0] * ia[2] * ia[3]                           # P, N, k, C: write E*C (gathers hit L2)
                if name == "spgan_bn_pool_fwd":
                    return 4.0 * ia[0] * ia[1]
                if name == "spgan_bn_pool_bwd":
                    return 8.0 * ia[0] * ia[1]
                if name in ("spgan_bn_dbl_bwd_reduce", "spgan_bn_act_dbl_bwd_reduce"):
                    return 12.0 * ia[0] * ia[1]
                if name in ("spgan_bn_dbl_bwd_apply", "spgan_bn_act_dbl_bwd_apply"):
                    return 20.0 * ia[0] * ia[1]
                if name == "spgan_segmax":
                    return 4.0 * ia[0] * ia[1]
                return None
            bw = {}
            for name, ia, s_, e_ in prof:
                try:
                    nb = nbytes(name, ia)
                except IndexError:
                    nb = None
                if nb:
                    for key in (name, "%s R=%d C=%d" % (name, ia[0], ia[1]) if name in ("spgan_colstats", "spgan_norm_bwd_reduce") else None):
                        if key is None:
                            continue
                        b = bw.setdefault(key, [0.0, 0.0, 0])
                        b[0] += nb
                        b[1] += s_.elapsed_time(e_)
                        b[2] += 1
                    continue
                if nb:
                    b = bw.setdefault(name, [0.0, 0.0, 0])
                    b[0] += nb
                    b[1] += s_.elapsed_time(e_)
                    b[2] += 1
            for name, b in sorted(bw.items(), key=lambda kv: -kv[1][1]):
                sys.stderr.write("BW %-32s x%-3d %8.3f ms %8.2f GB  %7.1f GB/s\n" % (name, b[2], b[1], b[0] / 1e9,
                                                                                     b[0] / 1e6 / b[1]))
        kernel_share = {k[len("spgan_"):]: round(a[0] / total, 4) for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]}
        kernel_share["_sum_of_kernel_ms"] = round(total, 3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        def peak_tf():
            return peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PFLOP/s"
        ncu = {}
        for fn in ("ncu_r1.json", "ncu_r2.json"):                     # later rounds override
            try:
                ncu.update(json.load(open(os.path.join(ROOT, "profiles", fn))))
            except OSError:
                pass
        # dominant kernel: the tcgen05 GEMM on its biggest shape, the critic's fc2 forward
        # (NT, M = B*N points, N = 1024, K = 256); algorithmic FLOPs per launch = 2*M*N*K
        dom = []
        for name, ia, s_, e_ in prof:
            if name == "spgan_gemm" and ia[0] == 0 and ia[1] == 1 and ia[3] == 1024 and ia[4] == 256:
                dom.append((ia[2], s_.elapsed_time(e_)))                    # (transA, transB, M, N, K, ...)
            elif name == "spgan_gemm_fused" and ia[2] == 1024 and ia[3] == 256:
                dom.append((ia[1], s_.elapsed_time(e_)))                    # (transB, M, N, K, ...)
        eng = pkg.ops.GEMM_ENGINE
        if dom and eng in (1, 2, 3):
            t_ms = sum(t for _, t in dom) / len(dom)
            Mrows = dom[0][0]
            fl = 2.0 * Mrows * 1024 * 256
            ach = fl / (t_ms / 1e3) / 1e12
            if eng == 3:
                kname, ckey = "gemm_ts_kernel (tcgen05 TS form, A resident in TMEM, fp16x3 split)", "gemm_ts_kernel"
                note = ("fp32-faithful fp16x3: 3 tcgen05 kind::f16 MMAs per product (hi*hi, hi*lo, lo*hi), so the "
                        "attainable fraction of the bf16 peak is 1/3 = 0.333")
            else:
                kname, ckey = "gemm_tc_kernel<128,TF32x3> (tcgen05)", "gemm_tc_kernel<128, 1"
                note = ("fp32-faithful TF32x3: 3 tcgen05 MMAs per product at half the bf16 rate, so the attainable "
                        "fraction of the bf16 peak is 1/6 = 0.167")
            cap = next((v for kk, v in ncu.items() if kk.startswith(ckey)), {})
            roofline = {"kernel": "%s, critic fc2 forward M=%d N=1024 K=256" % (kname, Mrows),
                        "bound": "tensor", "achieved": ach, "peak": peak_tf(), "unit": "TFLOP/s", "frac": ach / peak_tf(),
                        "traffic": cap.get("dram_traffic_bytes"),
                        "traffic_source": "profiles/ncu_r2.json (ncu --set full, same kernel and shape)" if cap else None,
                        "algorithmic_flops_per_launch": fl, "algorithmic_bytes_per_launch": 4.0 * Mrows * (1024 + 256),
                        "launches_per_step": len(dom), "avg_launch_ms": t_ms,
                        "share_of_step": sum(t for _, t in dom) / total,
                        "tensor_pipe_active_pct_ncu": cap.get("tensor_pipe_pct"),
                        "note": note, "peak_source": peak_src, "mma_per_product": 3,
                        "frac_of_attainable": ach / (peak_tf() / (3.0 if eng == 3 else 6.0))}
        g, gf, gw = agg.get("spgan_gemm"), agg.get("spgan_gemm_fused"), agg.get("spgan_gemm_wgrad_fused")
        if g or gf:
            g = [sum(x) for x in zip(g or [0.0, 0, 0.0], gf or [0.0, 0, 0.0], gw or [0.0, 0, 0.0])]
            ach = g[2] / (g[0] / 1e3) / 1e12
            roofline_all = {"kernel": "spgan_gemm + spgan_gemm_fused + spgan_gemm_wgrad_fused, all %d launches of the step "
                                      "(tensor-core and CUDA-core routes alike)" % g[1], "bound": "tensor",
                            "achieved": ach, "peak": peak_tf(), "unit": "TFLOP/s", "frac": ach / peak_tf(),
                            "share_of_step": g[0] / total}
            if roofline is None:
                roofline = dict(roofline_all, traffic=None, peak_source=peak_src)
        else:
            roofline_all = None
        kn = [(name, ia, s_.elapsed_time(e_)) for name, ia, s_, e_ in prof if name in ("spgan_knn_group", "spgan_knn_rows")]
        roofline_knn = None
        if kn:
            name, ia, t_ms = max(kn, key=lambda kv: kv[1][1])        # (B, C, N, k): the C = 64 launch (EdgeConv2's graph)
            Bk, Ck, Nk, kk = ia[0], ia[1], ia[2], ia[3]
            fl = 2.0 * Bk * Nk * Nk * Ck
            byts = 4.0 * (Bk * Ck * Nk + Bk * Nk * kk)
            ffma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
            tc = name == "spgan_knn_rows"
            roofline_knn = {"kernel": ("knn_tc_filter + knn_tc_refine (tcgen05 filter, exact fp32 refine: bit-exact lists)" if tc else
                                       "knn_group_fast_kernel (bit-exact fp32 recipe: CUDA cores only)") +
                                      " B=%d C=%d N=%d k=%d" % (Bk, Ck, Nk, kk),
                            "bound": "tensor (filter: 2 passes x 3 MMAs per product) + L2 gather (refine)" if tc else "fp32 FFMA + selection",
                            "ms": t_ms, "achieved": fl / (t_ms / 1e3) / 1e12, "peak": ffma_peak, "unit": "TFLOP/s (algorithmic 2 B N^2 C)",
                            "frac": fl / (t_ms / 1e3) / 1e12 / ffma_peak,
                            "peak_source": "nominal fp32 FFMA: 148 SM x 128 lanes x 2 x 1.965 GHz (what an all-pairs fp32 kernel is bound by)",
                            "hbm_algorithmic_bytes": byts, "hbm_gbs": byts / (t_ms / 1e3) / 1e9,
                            "hbm_frac_of_measured": byts / (t_ms / 1e3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                            "traffic": ncu.get("knn_tc_filter_kernel" if tc else "knn_group_fast_kernel", {}).get("dram_traffic_bytes")}
            if tc:
                roofline_knn["tensor_tflops_executed"] = 6.0 * fl / (t_ms / 1e3) / 1e12
                roofline_knn["frac_of_bf16_peak_executed"] = 6.0 * fl / (t_ms / 1e3) / 1e12 / peak_tf()

    # ---- sub-metrics of BASELINE's metric string: "kNN+EdgeConv ms/batch" and configs[1] (generator forward only)
    sub = None
    if rank == 0 and not minimal:
        def ev_ms(fn, reps=5):
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]
        d0 = resident[0]
        with torch.no_grad():
            z = d0["z_d"].expand(B, N, NZ)
            g_fwd = ev_ms(lambda: G(x, z))
            x1 = G._last_x1                                                  # EdgeConv2's input rows [B*N, 64]
            x1_bcn = pkg.ops.RowsToBcn.apply(x1, B, x1.shape[1], N)
            pc_rows = x.reshape(B * N, 3)
            pc_bcn = pkg.ops.RowsToBcn.apply(pc_rows, B, 3, N)
            knn2 = ev_ms(lambda: pkg.ops.knn_indices_rows(x1, B, N, G.nk))          # what the generator calls
            knn2_cuda_core = ev_ms(lambda: pkg.ops.knn_indices(x1_bcn, G.nk))
            knn1 = ev_ms(lambda: pkg.ops.knn_indices_rows(pc_rows, B, N, G.nk))       # what the generator calls (then caches)
            knn1_cuda_core = ev_ms(lambda: pkg.ops.knn_indices(pc_bcn, G.nk))
            idx2 = pkg.ops.knn_indices_rows(x1, B, N, G.nk)
            idx1 = pkg.ops.knn_indices_rows(pc_rows, B, N, G.nk)
            ec2 = ev_ms(lambda: G.EdgeConv2.forward_rows(x1, idx2, B, N))
            ec1 = ev_ms(lambda: G.EdgeConv1.forward_rows(pc_rows, idx1, B, N))
        sub = {"generator_forward_ms_per_batch": g_fwd, "generator_forward_clouds_per_s": B / (g_fwd / 1e3),
               "knn_edgeconv_ms_per_batch": knn1 + knn2 + ec1 + ec2,
               "knn_C3_ms": knn1, "knn_C3_cuda_core_kernel_ms": knn1_cuda_core, "knn_C64_ms": knn2,
               "knn_C64_cuda_core_kernel_ms": knn2_cuda_core, "edgeblock1_fwd_ms": ec1, "edgeblock2_fwd_ms": ec2,
               "note": "forward, train-mode BN, B=%d N=%d k=%d; the C=3 graph of the static sphere is cached inside "
                       "training steps (model.py:231) but counted here" % (B, N, G.nk)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch if args.cpu_batch > 0 else B
        t_cpu, cores = cpu_step_time(cb, N, args.cpu_steps, 1)
        cpu_baseline = {"value": cb / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d timed full WGAN-GP steps (after 1 warm-up) on %d of the %d clouds (N=%d), oracle "
                                  "port of the reference's torch CPU path" % (args.cpu_steps, cb, B, N),
                        "ms_per_step": 1e3 * t_cpu}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[2]: full G+D WGAN-GP step, Chair synthetic, N=%d B=%d per GPU" % (N, B),
                           "global_batch": world * B, "points": N, "k": 10, "parallelism": "dp%d" % world,
                           "sphere": ball_src, "gemm_engine": pkg.ops.GEMM_ENGINE, "cuda_graph": graph_note,
                           "l2": "no flush needed: per-step working set (activations) is several GB >> 126 MB L2"},
                "clocks": clocks, "e2e": e2e, "e2e_train_py": e2e_train_py, "gpu_launches": launches,
                "host_enqueue_ms_per_step": host_ms[0] if host_ms else None, "roofline": roofline,
                "roofline_all_gemm": roofline_all, "roofline_knn": roofline_knn, "submetrics": sub,
                "kernel_share": kernel_share, "cpu_baseline": cpu_baseline,
                "last_losses": losses[-1] if losses else None}
        print(json.dumps(line), file=real_stdout, flush=True)
    # ---- orderly teardown: drop the captured graph (it holds the NCCL all-reduces) before the process group
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    trainer.release_graph()
    torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
