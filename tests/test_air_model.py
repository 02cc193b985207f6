"""AIR-ASR training step: host logic on CPU (oracle ops), data-parallel equivalence over gloo (world size 2),
and -- on the GPU -- the CUDA-op step against the oracle-op step with identical weights and noise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mog_asr_b200 import synth
from mog_asr_b200.air import AIRModel, CudaOps, Trainer, config_from_flags
from oracle.air_ops import OracleOps, SeededNoise


def make_images(B, cs, seed=0):
    canv, num = synth.multi_object_canvases(B, cs, 28, (1, 2, 3), seed=seed)
    return torch.tensor(np.clip(canv, 0, 1).reshape(B, cs * cs)), num


def test_parameter_inventory_matches_reference():
    # 50 variables / 4 433 284 floats at canvas 50; 46 / 6 051 075 with fix_steps at canvas 64
    m = AIRModel(config_from_flags("mnist", "13", gm=100.0, gne=10.0), ops=OracleOps())
    assert len(list(m.parameters())) == 50 and sum(p.numel() for p in m.parameters()) == 4433284
    m = AIRModel(config_from_flags("sprites", "3", ds="bbox20k", gb=1.0, gs=10.0, ga=20.0), ops=OracleOps())
    assert len(list(m.parameters())) == 46 and sum(p.numel() for p in m.parameters()) == 6051075
    assert m.cfg.fix_steps == 3 and tuple(m.cfg.constrains_area_minmax) == (12, 15) and m.cfg.canvas_size == 64


def test_oracle_step_runs_and_updates():
    cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0)
    tr = Trainer(cfg, "cpu", ops=OracleOps())
    images, _ = make_images(8, 50)
    before = [p.detach().clone() for p in tr.params]
    out = tr.step(images, noise=SeededNoise(3, 8))
    assert torch.isfinite(out["loss"]) and 1 <= out["steps"] <= cfg.max_steps
    assert out["z_pres_probs"].shape == (8, out["steps"]) and out["rec_shifts"].shape == (8, out["steps"], 2)
    # per-tensor clip_by_norm(1.0): no gradient tensor may exceed norm 1 after post-processing (:1100-1111)
    assert all(float(g.norm()) <= 1.0 + 1e-5 for g in tr.grads)
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.params))
    # first Adam step moves every touched weight by ~lr
    delta = max(float((a - b).abs().max()) for a, b in zip(before, tr.params))
    assert delta <= cfg.learning_rate * 1.01


def test_fix_steps_prior_and_always_max_steps():
    cfg = config_from_flags("sprites", "3", ds="bbox20k", gb=1.0, gs=10.0, ga=20.0, always_max_steps=True)
    tr = Trainer(cfg, "cpu", ops=OracleOps())
    images, _ = make_images(4, 64)
    out = tr.step(images, noise=SeededNoise(5, 4))
    assert out["steps"] == cfg.max_steps and torch.isfinite(out["loss"])
    assert float(out["margin"]) == 0.0            # -gm unset: count penalties gated off (:973)


@pytest.mark.parametrize("flags", [dict(data="mnist", dn="13", gm=100.0, gne=10.0),
                                   dict(data="sprites", dn="3", ds="bbox20k", gb=1.0, gs=10.0, ga=20.0)])
def test_stacked_kl_equals_per_step_form(flags):
    """The fast form (KL terms evaluated once on [T,B,..] stacks after the loop) against the literal per-step form
    of the reference (:690-787): same loss, same gradients (only the summation order over steps differs)."""
    images, _ = make_images(6, 50 if flags["data"] == "mnist" else 64, seed=2)
    res = []
    for stacked in (False, True):
        cfg = config_from_flags(stacked_kl=stacked, **flags)
        tr = Trainer(cfg, "cpu", ops=OracleOps(), seed=5)
        out = tr.forward_backward(images, noise=SeededNoise(4, 6))
        res.append((float(out["loss"].detach()), out["elbo"].clone(), tr.flat_grad.clone(), out["steps"]))
    (l0, e0, g0, t0), (l1, e1, g1, t1) = res
    assert t0 == t1
    np.testing.assert_allclose(l1, l0, rtol=1e-6)
    assert torch.allclose(e1, e0, rtol=1e-5, atol=1e-4)
    fin = torch.isfinite(g0)
    assert torch.equal(torch.isfinite(g1), fin)
    assert float((g1[fin] - g0[fin]).abs().max()) <= 1e-5 * float(g0[fin].abs().max())


@pytest.mark.parametrize("flags,defer", [(dict(data="mnist", dn="13", gm=100.0, gne=10.0), True),
                                         (dict(data="sprites", dn="3", ds="bbox20k", gb=1.0, gs=10.0, ga=20.0), False)])
def test_batched_tail_equals_loop_form(flags, defer):
    """Dependency-ordered evaluation (generative LSTM, prior / z_pres heads, VAE decoder and canvas writes after the
    inference recurrence, the non-recurrent layers once over all T*B rows) against the step-by-step loop of the
    reference (:393-727) with the same fixed trip count: same loss, same per-image ELBO, same gradients (only GEMM
    row-batching and summation order over steps differ).  -dn 13 has learned z_pres priors, -dn 3 the fixed ones."""
    images, _ = make_images(6, 50 if flags["data"] == "mnist" else 64, seed=2)
    res = []
    for tail in (False, True):   # float64: in fp32 the reference cross-entropy (1e10 slopes at canvas == 0) amplifies the
        cfg = config_from_flags(always_max_steps=True, batched_tail=tail, **flags)      # GEMM re-batching noise to 4e-4
        tr = Trainer(cfg, "cpu", ops=OracleOps(), seed=5, defer_weight_grads=defer, dtype=torch.float64)
        out = tr.forward_backward(images.double(), noise=SeededNoise(4, 6, dtype=torch.float64))
        res.append((float(out["loss"].detach()), out["elbo"].clone(), tr.flat_grad.clone(), out["steps"], out["rec_num_digits"].clone(),
                    out["reconstruction"].clone()))
    (l0, e0, g0, t0, d0, r0), (l1, e1, g1, t1, d1, r1) = res
    assert t0 == t1 and torch.equal(d0, d1)
    np.testing.assert_allclose(l1, l0, rtol=1e-8)
    assert torch.allclose(e1, e0, rtol=1e-8, atol=1e-4)
    assert torch.allclose(r1, r0, atol=1e-6)
    fin = torch.isfinite(g0)
    assert torch.equal(torch.isfinite(g1), fin)
    assert float((g1[fin] - g0[fin]).abs().max()) <= 1e-4 * float(g0[fin].abs().max())   # the sampler itself stays fp32
    assert float(g0[fin].abs().max()) > 0


def test_deferred_weight_gradients_equal_autograd():
    """One GEMM per layer over the concatenated rows of all loop iterations == per-iteration autograd accumulation."""
    cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0)
    images, _ = make_images(6, 50, seed=3)
    grads = []
    for defer in (False, True):
        tr = Trainer(cfg, "cpu", ops=OracleOps(), seed=5, defer_weight_grads=defer)
        tr.forward_backward(images, noise=SeededNoise(4, 6))
        grads.append(tr.flat_grad.clone())
    g0, g1 = grads
    fin = torch.isfinite(g0)
    assert torch.equal(torch.isfinite(g1), fin)
    assert float((g1[fin] - g0[fin]).abs().max()) <= 1e-5 * float(g0[fin].abs().max())
    assert float(g0[fin].abs().max()) > 0


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (same intra-op thread count as the parent: torch's vectorised CPU kernels round the transcendental
        #  ops differently at chunk boundaries, and this loss is ill-conditioned through log(recon + 1e-10))
        cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0, always_max_steps=True)
        Bg = 8
        images, _ = make_images(Bg, 50)
        lo, hi = rank * Bg // world, (rank + 1) * Bg // world
        tr = Trainer(cfg, "cpu", process_group=dist.group.WORLD, global_batch=Bg,
                     ops=OracleOps(process_group=dist.group.WORLD, global_batch=Bg))
        out = tr.forward_backward(images[lo:hi], noise=SeededNoise(7, Bg, lo, hi))
        tr.reduce_gradients()
        if rank == 0:
            ret["flat_grad"] = tr.flat_grad.clone()
            ret["loss_local"] = float(out["loss"])
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradients_equal_single_process_gloo():
    """world_size-2 gloo: sharded forward/backward + all-reduce == the single-process global-batch gradient."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0, always_max_steps=True)
    Bg = 8
    images, _ = make_images(Bg, 50)
    tr = Trainer(cfg, "cpu", ops=OracleOps())
    tr.forward_backward(images, noise=SeededNoise(7, Bg))
    ref, got = tr.flat_grad, ret["flat_grad"]
    finite = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(got), finite)
    scale = ref[finite].abs().max()
    assert float((got[finite] - ref[finite]).abs().max()) <= 2e-3 * float(scale)


def _gloo_dp_check_worker(rank, world, port, ret):
    import torch.distributed as dist
    from mog_asr_b200.air import bench_train
    torch.set_num_threads(1)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = bench_train.dp_check("cpu", dist.group.WORLD, global_batch=8,
                                 ops_factory=lambda pg, gb: OracleOps(process_group=pg, global_batch=gb))
        if rank == 0:
            ret["dp_check"] = r
    finally:
        dist.destroy_process_group()


def test_dp_check_of_the_bench_over_gloo():
    """bench.py's ``train.dp_check`` (gradient equality, loss, trip counts of both loop forms) run on two CPU ranks over gloo
    with the oracle's operators: exercises the code the driver runs under NCCL at N > 1."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_gloo_dp_check_worker, args=(2, port, ret), nprocs=2, join=True)
    r = ret["dp_check"]
    assert r["ok"], r
    for mode in ("fixed_trip_count", "reference_loop"):
        assert r[mode]["loss_rel_diff"] <= 1e-5, r[mode]


def _nccl_dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    from mog_asr_b200.air import bench_train
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        r = bench_train.dp_check(dev, dist.group.WORLD, global_batch=32)
        if rank == 0:
            ret["dp_check"] = r
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_data_parallel_gradients_equal_single_process_nccl(cuda_device):
    """Two ranks on two GPUs over NCCL with the product operators: the all-reduced gradient equals the single-process gradient of
    the whole batch (including the [T] column-sum all-reduce of the marginal count penalty), and the reference's loop form runs
    the same trip count on every rank (the one-flag ``any`` all-reduce).  Needs two devices; bench.py records the same check
    at every N > 1 (``train.dp_check``)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (recorded by bench.py at N > 1 instead)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_nccl_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    r = ret["dp_check"]
    assert r["ok"], r


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [dict(data="mnist", dn="13", gm=100.0, gne=10.0),
                                   dict(data="sprites", dn="3", ds="bbox20k", gb=1.0, gs=10.0, ga=20.0)])
def test_cuda_step_matches_oracle_step(cuda_device, flags):
    """Configs 2 and 3: the training step with the CUDA operators vs the same step with the oracle's operators
    (torch restatements, run on the same device so the dense layers use the same GEMMs and the comparison
    isolates the hot path), identical weights and injected noise: loss, per-image terms, parameter gradients."""
    cfg = config_from_flags(always_max_steps=True, **flags)
    B = 16
    images, _ = make_images(B, cfg.canvas_size, seed=4)
    images = images.to(cuda_device)
    ref = Trainer(cfg, cuda_device, ops=OracleOps(), seed=11)
    got = Trainer(cfg, cuda_device, seed=11)                                     # the product configuration
    exact = Trainer(cfg, cuda_device, ops=CudaOps(fused_pointwise=False), seed=11)   # same kernels, elementwise math in torch ops
    got.model.load_state_dict(ref.model.state_dict())
    exact.model.load_state_dict(ref.model.state_dict())
    o_ref = ref.forward_backward(images, noise=SeededNoise(9, B, device=cuda_device))
    o_fused = got.forward_backward(images, noise=SeededNoise(9, B, device=cuda_device))
    o_got = exact.forward_backward(images, noise=SeededNoise(9, B, device=cuda_device))
    torch.cuda.synchronize()
    assert o_got["steps"] == o_ref["steps"] == o_fused["steps"]
    # with the fused per-step elementwise kernels expf/tanhf/logf differ from torch's by an ulp, the rest follows
    assert torch.allclose(o_fused["reconstruction"], o_ref["reconstruction"], rtol=1e-4, atol=2e-5)
    assert torch.equal(o_fused["rec_num_digits"], o_ref["rec_num_digits"])
    # (the reference cross-entropy amplifies ulp-level changes of the residue pixels by log(r + 1e-10): the loss value
    #  of the fused configuration is compared through the well-conditioned squared-error term instead)
    mse_ = lambda x, r: ((x - r) ** 2).sum(1) * 50.0
    l_ref = ref.forward_backward(images, noise=SeededNoise(9, B, device=cuda_device), recon_loss_fn=mse_)["loss"]
    l_fused = got.forward_backward(images, noise=SeededNoise(9, B, device=cuda_device), recon_loss_fn=mse_)["loss"]
    np.testing.assert_allclose(float(l_fused), float(l_ref), rtol=2e-5)
    # forward: the sampler and the composite are bit-exact; the fused cross-entropy and the ASR kernel are fp32
    # sums in a different order
    assert torch.equal(o_got["reconstruction"], o_ref["reconstruction"])
    assert torch.allclose(o_got["elbo"], o_ref["elbo"], rtol=2e-6, atol=1e-3)
    assert torch.equal(o_got["rec_num_digits"], o_ref["rec_num_digits"])
    np.testing.assert_allclose(o_got["per_image_reg"].cpu().numpy(), o_ref["per_image_reg"].cpu().numpy(), rtol=2e-4, atol=1e-3)
    np.testing.assert_allclose(float(o_got["loss"]), float(o_ref["loss"]), rtol=1e-5)
    # backward: with the reference's cross-entropy the upstream gradient is 1e10 wherever the canvas is exactly 0
    # under an object pixel; multiplied into the +w/-w border taps (|w| up to 1e4) it turns fp32 rounding into
    # O(1e4) noise in the *reference-op* gradients (the CUDA path skips those exactly-cancelling taps).  The
    # gradient comparison therefore uses a well-conditioned reconstruction term (squared error) through the same
    # graph -- every operator's backward is exercised, none of it is amplified.
    # ... evaluated with the oracle's operators in fp64: their fp32 evaluation still carries O(0.3) noise at
    # the window's border pixels (thousands of +w/-w tap pairs with |w| up to 1e4 land on them).
    mse = lambda x, r: ((x - r) ** 2).sum(1) * 50.0
    ref64 = Trainer(cfg, cuda_device, ops=OracleOps(), seed=11, dtype=torch.float64)
    ref64.model.load_state_dict(ref.model.state_dict())
    ref64.forward_backward(images.double(), noise=SeededNoise(9, B, device=cuda_device, dtype=torch.float64), recon_loss_fn=mse)
    got.forward_backward(images, noise=SeededNoise(9, B, device=cuda_device), recon_loss_fn=mse)
    g_ref, g_got = ref64.flat_grad.cpu(), got.flat_grad.cpu().double()
    finite = torch.isfinite(g_ref)
    assert torch.equal(torch.isfinite(g_got), finite)
    off = 0
    for (name, p) in ref.model.named_parameters():
        n = p.numel()
        a, b = g_ref[off:off + n], g_got[off:off + n]
        off += n
        m = torch.isfinite(a)
        if m.any():
            assert float((a[m] - b[m]).abs().max()) <= 2e-3 * float(a[m].abs().max()) + 1e-12, name


@pytest.mark.gpu
@pytest.mark.parametrize("always_max", [False, True])
@pytest.mark.parametrize("dn", ["13", "2"])
def test_loop_form_gradients_with_the_fused_kernels(cuda_device, always_max, dn):
    """The reference's loop form (``while any(stop_sum < thr)``, :386-390) and the fixed-trip-count form, learned z_pres
    prior (``-dn 13``) and fixed step count (``-dn 2``), with the product operators (fused kernels, deferred weight
    gradients) against the oracle's operators in float64: every parameter gradient, including the layers whose input
    does not require grad in the first iteration (the z_pres prior head sees the all-zero initial state, :609-615)."""
    cfg = config_from_flags("mnist", dn, gm=100.0, gne=10.0, always_max_steps=always_max)
    B = 12
    images, _ = make_images(B, 50, seed=8)
    images = images.to(cuda_device)
    got = Trainer(cfg, cuda_device, seed=5)
    ref = Trainer(cfg, cuda_device, ops=OracleOps(), seed=5, dtype=torch.float64)
    ref.model.load_state_dict(got.model.state_dict())
    mse = lambda x, r: ((x - r) ** 2).sum(1) * 50.0
    o_ref = ref.forward_backward(images.double(), noise=SeededNoise(3, B, device=cuda_device, dtype=torch.float64), recon_loss_fn=mse)
    o_got = got.forward_backward(images, noise=SeededNoise(3, B, device=cuda_device), recon_loss_fn=mse)
    assert o_got["steps"] == o_ref["steps"]
    np.testing.assert_allclose(float(o_got["loss"]), float(o_ref["loss"]), rtol=2e-5)
    g_ref, g_got = ref.flat_grad.cpu(), got.flat_grad.cpu().double()
    off = 0
    for (name, p) in got.model.named_parameters():
        n = p.numel()
        a, b = g_ref[off:off + n], g_got[off:off + n]
        off += n
        assert torch.isfinite(b).all(), name
        assert float(a.abs().max()) > 0 or float(b.abs().max()) == 0, name
        assert float((a - b).abs().max()) <= 2e-3 * float(a.abs().max()) + 1e-12, name


@pytest.mark.gpu
def test_graph_step_with_parallel_branches_equals_single_stream_graph(cuda_device, monkeypatch):
    """The captured training step records the generative chain and the weight-gradient flushes as parallel branches of
    the CUDA graph (side streams that fork from and rejoin the capture stream).  Same seeds, same batches: after three
    replays the parameters must equal those of the single-stream capture (up to the atomics of the head kernels), and
    capturing must leave the training state untouched (the warm-up steps are rolled back)."""
    cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0, always_max_steps=True)
    B = 32
    images, _ = make_images(B, 50, seed=12)
    images = images.to(cuda_device)
    finals = []
    for streams in ("1", "0"):
        monkeypatch.setenv("MOG_AIR_STREAMS", streams)
        tr = Trainer(cfg, cuda_device, seed=21)
        before = [p.detach().clone() for p in tr.params]
        tr.capture(B)
        assert all(torch.equal(a, b) for a, b in zip(before, tr.params)) and tr.t == 0 and float(tr.t_dev) == 0.0
        torch.cuda.manual_seed(777)                      # identical noise streams for both captures
        out = tr.step_graph(images)
        torch.cuda.synchronize()
        first_grad = tr.flat_grad.detach().clone()       # the gradient of the first replay: same weights in both captures
        for _ in range(2):
            out = tr.step_graph(images)
        torch.cuda.synchronize()
        assert torch.isfinite(out["loss"]).all()
        finals.append(([p.detach().clone() for p in tr.params], first_grad))
    (pa, ga), (pb, gb) = finals
    # identical up to the order of the head kernels' atomics
    assert float((ga - gb).abs().max()) <= 1e-5 * float(gb.abs().max()), (float((ga - gb).abs().max()), float(gb.abs().max()))
    for a, b in zip(pa, pb):
        # Adam moves every weight by ~lr = 1e-4 per step whatever the gradient's size, so an entry whose gradient is rounding
        # noise may differ by a good part of one step after three; anything systematic would show up as several steps
        assert float((a - b).abs().max()) <= 1e-4, float((a - b).abs().max())


@pytest.mark.gpu
def test_one_step_changes_parameters_like_the_oracle_step(cuda_device):
    """After one full step (clip + TF-style Adam) both implementations hold the same parameters."""
    cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0, always_max_steps=True)
    B = 8
    images, _ = make_images(B, 50, seed=6)
    images = images.to(cuda_device)
    got = Trainer(cfg, cuda_device, seed=3)
    ref = Trainer(cfg, cuda_device, ops=OracleOps(), seed=3, dtype=torch.float64)
    ref.model.load_state_dict(got.model.state_dict())
    mse = lambda x, r: ((x - r) ** 2).sum(1) * 50.0
    for tr, dt in ((ref, torch.float64), (got, torch.float32)):
        tr.forward_backward(images.to(dt), noise=SeededNoise(2, B, device=cuda_device, dtype=dt), recon_loss_fn=mse)
        tr.reduce_gradients()
        tr.postprocess_and_apply()
    # lr = 1e-4: the first Adam step moves each weight by lr * g / (|g| + eps') ~ lr * sign(g); entries whose
    # gradient is ~eps (1e-8) may legitimately land anywhere in [-lr, lr], so: all within 2 lr, 99.9% within 2% of lr
    bad = tot = 0
    for a, b in zip(ref.params, got.params):
        d = (a.detach() - b.detach().double()).abs()
        assert float(d.max()) <= 2.1e-4
        bad += int((d > 2e-6).sum()); tot += d.numel()
    assert bad <= 1e-3 * tot, (bad, tot)


def test_concrete_sample_and_kl_equal_the_reference_source_run_on_the_tf_shim():
    """``tests/golden/graph_concrete.npz``: the reference's own ``air/concrete.py`` executed in float64 on the numpy TF shim
    with injected uniform draws.  Pins the graph of the Concrete pre-sigmoid sample (concrete.py:20-27) and of the Monte
    Carlo KL (:30-64) that ``OracleOps.zpres`` / ``AIRModel._concrete_kl`` restate (and the fused kernels are tested
    against those in tests/test_air_fused_gpu.py)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_concrete.npz"))
    lo, prior, u = (torch.tensor(g[k], dtype=torch.float64) for k in ("log_odds", "prior_lo", "u"))
    for temp in (0.1, 1.0):
        y, z, _, _, _ = OracleOps().zpres(lo, u, torch.zeros_like(lo), temp, 0.9)
        np.testing.assert_allclose(y.numpy(), g[f"y_T{temp}"], rtol=1e-12, atol=1e-12)
        kl = AIRModel._concrete_kl(y, prior, lo, temp)
        np.testing.assert_allclose(kl.numpy(), g[f"kl_T{temp}"], rtol=1e-10, atol=1e-10)
        # concrete_binary_sample (:4-17) returns sigmoid(y / T) of the un-divided y: the same z_pres
        np.testing.assert_allclose(z.numpy(), g[f"sig_T{temp}"], rtol=1e-12, atol=1e-15)


def test_loop_body_kl_masks_and_canvas_equal_the_reference_source_run_on_the_tf_shim():
    """``tests/golden/graph_loop_kl.npz``: lines :683-787 of the reference's loop body exec'd per step on the torch TF shim in
    float64 (``tests/golden/make_golden_loop_kl.py``).  The model's restatement must give the same per-step KL arrays (the
    z_pres term masked by the PREVIOUS stopping sum, the others and the canvas write by the UPDATED one), the same digit
    counts, stopping sums and canvas, and the same gradients of a weighted sum of the KL part of the ELBO and the canvas."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_loop_kl.npz"))
    cv = dict(zip([str(k) for k in g["cfg_keys"]], g["cfg_vals"]))
    cfg = config_from_flags("mnist", "13", zt=float(cv["z_pres_temperature"]))
    assert (cfg.stopping_threshold, cfg.scale_prior_mean, cfg.scale_prior_variance, cfg.vae_prior_mean, cfg.vae_prior_variance) == \
        tuple(float(cv[k]) for k in ("stopping_threshold", "scale_prior_mean", "scale_prior_variance", "vae_prior_mean", "vae_prior_variance"))
    model, ops = AIRModel(cfg, ops=OracleOps()), OracleOps()
    keys = ("y", "prior_lo", "post_lo", "sc_mean", "sc_lv", "sh_mean", "sh_lv", "g_sh_mean", "g_sh_lv", "v_mean", "v_lv", "window")
    t = {k: torch.tensor(g[k], dtype=torch.float64, requires_grad=True) for k in keys}
    T, B = g["y"].shape
    cs = int(cv["canvas_size"])
    stop, canvas = torch.zeros(B, dtype=torch.float64), torch.zeros(B, cs * cs, dtype=torch.float64)
    digits = torch.zeros(B, dtype=torch.int32)
    kls = {k: [] for k in ("z_pres_kl", "scale_kl", "shift_kl", "vae_kl")}
    for s in range(T):
        z = torch.sigmoid(t["y"][s])
        active_prev = stop < cfg.stopping_threshold
        stop = stop + (1.0 - z)                                                                   # OracleOps.zpres, written out
        active = stop < cfg.stopping_threshold
        digits = digits + active.to(torch.int32)
        canvas = canvas + torch.where(active[:, None], z[:, None] * t["window"][s].reshape(B, -1), torch.zeros_like(canvas))
        zk = AIRModel._concrete_kl(t["y"][s], t["prior_lo"][s], t["post_lo"][s], cfg.z_pres_temperature)
        sk, hk, vk = model._gaussian_kls(t["sc_mean"][s], t["sc_lv"][s], t["sh_mean"][s], t["sh_lv"][s], t["g_sh_mean"][s],
                                         t["g_sh_lv"][s], t["v_mean"][s], t["v_lv"][s])
        kls["z_pres_kl"].append(torch.where(active_prev, zk, torch.zeros_like(zk)))
        for k, v in (("scale_kl", sk), ("shift_kl", hk), ("vae_kl", vk)):
            kls[k].append(torch.where(active, v, torch.zeros_like(v)))
    tol = dict(rtol=1e-10, atol=1e-10)
    for k, v in kls.items():
        np.testing.assert_allclose(torch.stack(v, 0).detach().numpy(), g["kl_" + k], **tol)
    np.testing.assert_allclose(stop.detach().numpy(), g["stop_sum"], **tol)
    assert np.array_equal(digits.numpy(), g["digits"])
    np.testing.assert_allclose(canvas.detach().numpy(), g["canvas"], **tol)
    elbo_kl = sum(torch.stack(v, 0).sum(0) for v in kls.values())
    np.testing.assert_allclose(elbo_kl.detach().numpy(), g["elbo_kl"], **tol)
    ((elbo_kl * torch.tensor(g["w"])).sum() + (canvas * torch.tensor(g["wc"])).sum()).backward()
    for k in keys:
        got = t[k].grad.numpy() if t[k].grad is not None else np.zeros_like(g[k])
        np.testing.assert_allclose(got, g["d_" + k], rtol=1e-9, atol=1e-9, err_msg=k)


def _load_reference_graph_run(name):
    """weights / noise / outputs of the reference's whole training graph executed on the torch TF shim
    (tests/golden/make_golden_model.py) and the matching re-hosted model with those weights loaded"""
    from mog_asr_b200.air.model import AIRConfig
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"graph_model_{name}.npz"))
    c = eval(str(g["cfg"]), {"__builtins__": {}})
    cfg = AIRConfig(canvas_size=c["canvas"], windows_size=c["ws"], max_steps=c["max_steps"], rnn_units=c["rnn"], vae_latent_dimensions=c["lat"],
                    vae_recognition_units=c["rec"], vae_generative_units=c["gen"], scale_hidden_units=c["hid"], shift_hidden_units=c["hid"],
                    z_pres_hidden_units=c["hid"], z_pres_temperature=c["zt"], constrains_num=c["counts"], constrains_num_gamma=c["gn"],
                    constrains_margin_gamma=c["gm"], constrains_num_element_gamma=c["gne"], constrains_bbox_gamma=c["gb"],
                    constrains_sharesize_gamma=c["gs"], constrains_area_gamma=c["ga"], constrains_area_minmax=tuple(c["minmax"]),
                    fix_steps=c["counts"][0] if len(c["counts"]) == 1 else None)
    model = AIRModel(cfg, ops=OracleOps()).double()
    P = "air/air_model/"                                           # TF variable names -> parameters (dense kernels are [in, out])
    dense = lambda tfname, lyr, w="kernel", b="bias": {lyr + ".weight": (P + tfname + "/" + w, True), lyr + ".bias": (P + tfname + "/" + b, False)}
    mp = {"infer_cell.kernel": (P + "infer_rnn_running/lstm_cell/kernel", False), "infer_cell.bias": (P + "infer_rnn_running/lstm_cell/bias", False),
          "gen_cell.kernel": (P + "gen_rnn_running/lstm_cell/kernel", False), "gen_cell.bias": (P + "gen_rnn_running/lstm_cell/bias", False)}
    for head in ("inf_shift", "inf_scale", "gen_shift"):           # hidden_m, mean, hidden_v, log-variance: the order the reference creates them
        for tfl, mine in (("dense", "hm"), ("dense_1", "m"), ("dense_2", "hv"), ("dense_3", "v")):
            mp.update(dense(f"{head}/{tfl}", f"{head}.{mine}"))
    for i in (1, 2):
        mp.update(dense(f"vae/recognition_{i}", f"vae_rec.{i - 1}", "weights", "biases"))
        mp.update(dense(f"vae/generative_{i}", f"vae_gen.{i - 1}", "weights", "biases"))
    for tfl, mine in (("rec_mean", "vae_rec_mean"), ("rec_log_variance", "vae_rec_logvar"), ("gen_mean", "vae_gen_mean")):
        mp.update(dense("vae/" + tfl, mine, "weights", "biases"))
    mp.update(dense("z_pres/log_odds/dense", "z_post_h")); mp.update(dense("z_pres/log_odds/dense_1", "z_post"))
    if cfg.fix_steps is None:
        mp.update(dense("z_pres/prior/dense", "z_prior_h")); mp.update(dense("z_pres/prior/dense_1", "z_prior"))
    params = dict(model.named_parameters())
    assert set(params) == set(mp) and len([k for k in g.files if k.startswith("w:")]) == len(mp)     # same variables, nothing left over
    with torch.no_grad():
        for n, (tfn, tr) in mp.items():
            w = torch.tensor(g["w:" + tfn])
            params[n].copy_(w.t() if tr else w)
    return g, cfg, model, params, mp


def _load_generation_run(name):
    """weights / noise / outputs of the reference's generation graph run on the torch TF shim (tests/golden/make_golden_generation.py)"""
    from mog_asr_b200.air.model import AIRConfig
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"graph_generation_{name}.npz"))
    c = eval(str(g["cfg"]), {"__builtins__": {}})
    cfg = AIRConfig(canvas_size=c["canvas"], windows_size=c["ws"], max_steps=c["max_steps"], rnn_units=c["rnn"], vae_latent_dimensions=c["lat"],
                    vae_recognition_units=c["rec"], vae_generative_units=c["gen"], scale_hidden_units=c["hid"], shift_hidden_units=c["hid"],
                    z_pres_hidden_units=c["hid"], z_pres_temperature=c["zt"], constrains_num=c["counts"],
                    constrains_area_minmax=tuple(c["minmax"]), fix_steps=c["counts"][0] if len(c["counts"]) == 1 else None)
    model = AIRModel(cfg, ops=OracleOps()).double()
    P = "air/air_model/"
    dense = lambda tfname, lyr, w="kernel", b="bias": {lyr + ".weight": (P + tfname + "/" + w, True), lyr + ".bias": (P + tfname + "/" + b, False)}
    mp = {"gen_cell.kernel": (P + "gen_rnn_running/lstm_cell/kernel", False), "gen_cell.bias": (P + "gen_rnn_running/lstm_cell/bias", False)}
    for tfl, mine in (("dense", "hm"), ("dense_1", "m"), ("dense_2", "hv"), ("dense_3", "v")):
        mp.update(dense(f"gen_shift/{tfl}", f"gen_shift.{mine}"))
    for i in (1, 2):
        mp.update(dense(f"vae/generative_{i}", f"vae_gen.{i - 1}", "weights", "biases"))
    mp.update(dense("vae/gen_mean", "vae_gen_mean", "weights", "biases"))
    if cfg.fix_steps is None:
        mp.update(dense("z_pres/prior/dense", "z_prior_h")); mp.update(dense("z_pres/prior/dense_1", "z_prior"))
    params = dict(model.named_parameters())
    with torch.no_grad():                                          # (the generation graph uses the generative half of the weights only)
        for n, (tfn, tr) in mp.items():
            w = torch.tensor(g["w:" + tfn])
            params[n].copy_(w.t() if tr else w)
    return g, cfg, model


@pytest.mark.parametrize("name", ["learned_prior", "fix_steps"])
def test_generate_equals_the_reference_generation_graph_run_on_the_tf_shim(name):
    """``AIRModel.generate`` against the reference's own ``_create_generation`` (with its ``vae_generation``, ``concrete.py`` and
    ``transformer.py``) executed on the torch TF shim in float64 with the same weights and injected noise: same trip count of the
    data-dependent loop, same object counts, same write transforms (the constant-scale override of :1203-1205) and the same
    generated canvases (Bernoulli-binarised windows written through the sampler)."""
    g, cfg, model = _load_generation_run(name)
    noise = lambda kind, step, shape: torch.tensor(g["noise_" + kind][step]).reshape(shape)
    out = model.generate(g["samples"].shape[0], noise=noise)
    assert out["steps"] == int(g["steps"])
    assert np.array_equal(out["num"].numpy(), g["num"])
    np.testing.assert_allclose(out["thetas"].numpy(), g["thetas"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["samples"].numpy(), g["samples"], atol=2e-6)      # fp32-rounded sampler constants of the oracle


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["learned_prior", "fix_steps"])
def test_cuda_generate_equals_the_reference_generation_graph(name, cuda_device):
    """The product path of the generation graph (fp32, the fused write + composite kernel) with the golden weights and noise:
    same trip count and counts, canvases within fp32 rounding of the reference run."""
    g, cfg, model64 = _load_generation_run(name)
    model = AIRModel(cfg, ops=CudaOps()).to(cuda_device)
    with torch.no_grad():
        for (n, p), (_, q) in zip(model.named_parameters(), model64.named_parameters()):
            p.copy_(q.to(torch.float32))
    noise = lambda kind, step, shape: torch.tensor(g["noise_" + kind][step], dtype=torch.float32, device=cuda_device).reshape(shape)
    out = model.generate(g["samples"].shape[0], noise=noise)
    assert out["steps"] == int(g["steps"])
    assert np.array_equal(out["num"].cpu().numpy(), g["num"])
    np.testing.assert_allclose(out["thetas"].cpu().numpy(), g["thetas"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(out["samples"].cpu().numpy(), g["samples"], atol=1e-5)


@pytest.mark.parametrize("name", ["c2", "c3", "all"])
def test_rehosted_model_equals_the_reference_graph_run_on_the_tf_shim(name):
    """The reference's WHOLE training graph (``AIRModel._create_model`` with its own vae / concrete / transformer files) was
    executed on the torch TF shim in float64 with seeded weights and injected noise (``tests/golden/make_golden_model.py``;
    regulariser flags of BASELINE configs 2 and 3, shrunken layer widths).  The re-hosted model with the same weights and
    noise must run the same number of loop iterations (the data-dependent ``while`` condition), infer the same counts and
    reproduce the loss, the per-image terms, the canvas and the gradient of the loss w.r.t. every weight.  Library layers
    (dense, LSTMCell) are the shim's definition; what is pinned is the wiring of the whole step."""
    g, cfg, model, params, mp = _load_reference_graph_run(name)
    noise = lambda kind, step, shape: torch.tensor(g["noise_" + kind][step]).reshape(shape)
    out = model(torch.tensor(g["images"]), noise=noise)
    out["loss"].backward()
    assert out["steps"] == int(g["steps"]) and (name == "all" or out["steps"] < cfg.max_steps)   # c2 / c3: the loop really stopped early
    assert np.array_equal(out["rec_num_digits"].numpy(), g["rec_num_digits"])
    np.testing.assert_allclose(float(out["loss"].detach()), float(g["loss"]), rtol=1e-8)
    np.testing.assert_allclose(float(out["margin"]), float(g["margin"]), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["rec_scales"].numpy(), g["rec_scales"], atol=1e-8)
    np.testing.assert_allclose(out["rec_shifts"].numpy(), g["rec_shifts"], atol=1e-8)
    np.testing.assert_allclose(out["z_pres_probs"].numpy(), g["z_pres_probs"], atol=1e-8)
    np.testing.assert_allclose(out["reconstruction"].numpy(), g["reconstruction"], atol=2e-6)    # fp32-rounded sampler constants
    np.testing.assert_allclose(out["recon"].numpy(), g["recon_loss"], rtol=1e-7)
    for n, (tfn, tr) in mp.items():
        ref = torch.tensor(g["g:" + tfn])
        ref = ref.t() if tr else ref
        assert float(ref.abs().max()) > 0, n                                                        # every weight is live in the graph
        assert float((params[n].grad - ref).abs().max()) <= 3e-4 * float(ref.abs().max()), n


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2", "c3"])
def test_cuda_model_forward_equals_the_reference_graph_run_on_the_tf_shim(name, cuda_device):
    """The PRODUCT path (fp32, libmogstn kernels incl. the fused heads / epilogues / KL / loss) with the weights and noise of
    ``tests/golden/graph_model_*.npz`` against the reference's own graph executed on the torch TF shim in float64: same loop
    trip count and counts; scales, shifts, z_pres probabilities, canvas, regularisers and the marginal term within fp32
    rounding.  NOT compared here: the reconstruction cross-entropy and the loss that contains it -- ``log(1e-10 + r)`` turns
    the fp32 border residue of the sampler (|r| <= 1.5e-5 where float64 has 1e-17) into O(1) per pixel;
    that term is pinned in float64 on the CPU (above) and kernel-vs-oracle in fp32 (tests/test_bce_gpu.py)."""
    g, cfg, model64, params, mp = _load_reference_graph_run(name)
    noise64 = lambda kind, step, shape: torch.tensor(g["noise_" + kind][step]).reshape(shape)
    ref = model64(torch.tensor(g["images"]), noise=noise64)                      # float64 re-host == reference graph (test above)
    model = AIRModel(cfg, ops=CudaOps()).to(cuda_device)
    with torch.no_grad():
        for (n, p), (_, q) in zip(model.named_parameters(), model64.named_parameters()):
            p.copy_(q.to(torch.float32))
    model.set_deferred_weight_grads(True)
    noise = lambda kind, step, shape: torch.tensor(g["noise_" + kind][step], dtype=torch.float32, device=cuda_device).reshape(shape)
    out = model(torch.tensor(g["images"], dtype=torch.float32, device=cuda_device), noise=noise)
    assert out["steps"] == int(g["steps"])
    assert np.array_equal(out["rec_num_digits"].cpu().numpy(), g["rec_num_digits"])
    dev = dict(scales=float(np.abs(out["rec_scales"].cpu().numpy() - g["rec_scales"]).max()),
               shifts=float(np.abs(out["rec_shifts"].cpu().numpy() - g["rec_shifts"]).max()),
               probs=float(np.abs(out["z_pres_probs"].cpu().numpy() - g["z_pres_probs"]).max()),
               canvas=float(np.abs(out["reconstruction"].cpu().numpy() - g["reconstruction"]).max()),
               reg=float(np.abs(out["per_image_reg"].cpu().numpy() / ref["per_image_reg"].numpy() - 1).max()),
               margin=abs(float(out["margin"]) - float(g["margin"])) / max(1.0, abs(float(g["margin"]))))
    print("fp32 product path vs float64 reference graph:", dev)
    limits = dict(scales=2e-6, shifts=2e-6, probs=2e-6, canvas=1e-4, reg=1e-4, margin=1e-5)
    assert all(dev[k] <= limits[k] for k in limits), dev


def test_gradient_cleaning_and_clipping_equal_the_reference_lines():
    """``tests/golden/graph_gradpost.npz``: lines :1100-1111 of the reference (inf -> 0, nan -> 0, per-variable clip_by_norm)
    exec'd on the torch TF shim for gradients holding infinities, NaNs, large and small norms."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_gradpost.npz"))
    tr = Trainer(config_from_flags("mnist", "13"), "cpu", ops=OracleOps())
    n = len([k for k in g.files if k.startswith("in")])
    assert tr.cfg.gradient_clipping_norm == float(g["clip"])
    tr.flat_grad.zero_()
    used = []
    for k in range(n):                                   # drop each golden gradient into a variable's slot that is large enough
        src = torch.tensor(g[f"in{k}"]).reshape(-1)
        slot = next(i for i, v in enumerate(tr.grads) if v.numel() >= src.numel() and i not in used)
        used.append(slot)
        tr.grads[slot].reshape(-1)[:src.numel()] = src
    tr.clean_and_clip()
    for k, slot in enumerate(used):
        want = g[f"out{k}"].reshape(-1)
        got = tr.grads[slot].reshape(-1)[:want.size].numpy()
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-12)
        assert np.all(tr.grads[slot].reshape(-1)[want.size:].numpy() == 0)
    assert torch.isfinite(tr.flat_grad).all()


@pytest.mark.parametrize("name", ["c2", "c3", "all"])
def test_rehosted_test_model_equals_the_reference_graph_run_on_the_tf_shim(name):
    """``train=False`` (the reference's test model, z_pres rounded right after the sigmoid, :634-635) against the same
    reference graph executed with ``self.train = False``: trip count, hard counts, scales, shifts, probabilities, canvas --
    the quantities ``detection.evaluation`` consumes."""
    g, cfg, model, params, mp = _load_reference_graph_run(name)
    noise = lambda kind, step, shape: torch.tensor(g["noise_" + kind][step]).reshape(shape)
    with torch.no_grad():
        out = model(torch.tensor(g["images"]), noise=noise, train=False)
    assert out["steps"] == int(g["test_steps"])
    assert np.array_equal(out["rec_num_digits"].numpy(), g["test_rec_num_digits"])
    np.testing.assert_allclose(out["rec_scales"].numpy(), g["test_rec_scales"], atol=1e-8)
    np.testing.assert_allclose(out["rec_shifts"].numpy(), g["test_rec_shifts"], atol=1e-8)
    np.testing.assert_allclose(out["z_pres_probs"].numpy(), g["test_z_pres_probs"], atol=1e-8)
    np.testing.assert_allclose(out["reconstruction"].numpy(), g["test_reconstruction"], atol=2e-6)


@pytest.mark.gpu
def test_device_side_evaluation_pass_matches_the_reference_metrics(cuda_device):
    """train_air_pr.py:314-334 on the device: test model (rounded z_pres) on a labelled batch from the on-device feeder,
    detection metrics by the CUDA kernel -- against the oracle restatement of the reference's ``evaluation`` (itself pinned
    to the reference function) fed with the same inferred boxes."""
    from mog_asr_b200.air import evaluate_detection
    from mog_asr_b200.dataset import DeviceMultiObjectDataset, default_sprites
    from oracle import detection_ref
    cfg = config_from_flags("mnist", "13")
    torch.manual_seed(0)
    model = AIRModel(cfg, ops=CudaOps()).to(cuda_device)
    ds = DeviceMultiObjectDataset(default_sprites(32, 28, seed=2), 50, (1, 2, 3), (11, 15), mode="disjoint", seed=4, device=cuda_device)
    batch = ds.batch(0, 96)
    gen = torch.Generator(device=cuda_device).manual_seed(9)
    bank = {}

    def noise(kind, step, shape):                       # the same draws for both passes
        key = (kind, step)
        if key not in bank:
            bank[key] = (torch.rand(shape, device=cuda_device, generator=gen).clamp(1e-4, 1 - 1e-4) if kind == "concrete"
                         else torch.randn(shape, device=cuda_device, generator=gen))
        return bank[key]
    res = evaluate_detection(model, batch, noise=noise)
    with torch.no_grad():
        out = model(torch.clamp(batch["images"], 0, 1), noise=noise, train=False)
    pos, size, num = (batch[k].cpu().numpy() for k in ("pos", "size", "num"))
    gt_pos = [pos[b, :num[b]].reshape(-1).tolist() for b in range(len(num))]
    gt_size = [size[b, :num[b]].reshape(-1).tolist() for b in range(len(num))]
    want = detection_ref.evaluation(gt_pos, gt_size, out["rec_shifts"].cpu().numpy(), out["rec_scales"].cpu().numpy(),
                                    out["rec_num_digits"].cpu().numpy(), csize=50)
    np.testing.assert_allclose(res["precision"].cpu().numpy(), want[0], atol=1e-12)
    np.testing.assert_allclose(res["recall"].cpu().numpy(), want[1], atol=1e-12)
    for k, w in zip(("gt_max_iou", "detected_max_iou", "global_iou"), want[2:]):
        np.testing.assert_allclose(float(res[k]), w, atol=1e-12)
    assert 0.0 <= float(res["accuracy"]) <= 1.0 and 1 <= res["steps"] <= cfg.max_steps
