"""GPU parity of the fused per-step elementwise kernels (csrc/mog_air.cu) against the op-for-op torch restatement in
oracle/air_ops.py (values and autograd gradients, fp64 yardstick)."""
import numpy as np
import pytest
import torch

from mog_asr_b200.air import fused
from mog_asr_b200.air import fused as fused_mod
from oracle.air_ops import OracleOps

pytestmark = pytest.mark.gpu
REF = OracleOps()


def _check(got, ref, rtol=2e-6, atol=1e-7):
    np.testing.assert_allclose(got.detach().cpu().double().numpy(), ref.detach().cpu().numpy(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("act", [None, "tanh", "sigmoid"])
@pytest.mark.parametrize("shape", [(64, 2), (256, 1), (33, 50)])
def test_gauss_sample(cuda_device, act, shape):
    g = torch.Generator(device=cuda_device).manual_seed(1)
    mean = torch.randn(shape, device=cuda_device, generator=g).requires_grad_(True)
    lv = (torch.randn(shape, device=cuda_device, generator=g) * 0.7).requires_grad_(True)
    eps = torch.randn(shape, device=cuda_device, generator=g)
    gl, gs = torch.randn(shape, device=cuda_device, generator=g), torch.randn(shape, device=cuda_device, generator=g)
    lat, sq = fused.gauss_sample(mean, lv, eps, act)
    (lat * gl).sum().backward() if act is None else ((lat * gl).sum() + (sq * gs).sum()).backward()
    m64, l64 = mean.detach().double().requires_grad_(True), lv.detach().double().requires_grad_(True)
    rlat, rsq = REF.gauss_sample(m64, l64, eps.double(), act)
    (rlat * gl.double()).sum().backward() if act is None else ((rlat * gl.double()).sum() + (rsq * gs.double()).sum()).backward()
    _check(lat, rlat)
    if act is not None:
        _check(sq, rsq)
    else:
        assert sq is None
    _check(mean.grad, m64.grad, rtol=1e-5, atol=1e-6)
    _check(lv.grad, l64.grad, rtol=1e-5, atol=1e-6)


def test_thetas(cuda_device):
    g = torch.Generator(device=cuda_device).manual_seed(2)
    B = 300
    sh = torch.tanh(torch.randn((B, 2), device=cuda_device, generator=g)).requires_grad_(True)
    sc = torch.sigmoid(torch.randn((B, 1), device=cuda_device, generator=g) - 1).requires_grad_(True)
    gr, gw = torch.randn((B, 6), device=cuda_device, generator=g), torch.randn((B, 6), device=cuda_device, generator=g)
    tr, tw = fused.thetas(sh, sc)
    ((tr * gr).sum() + (tw * gw).sum()).backward()
    sh32, sc32 = sh.detach().clone().requires_grad_(True), sc.detach().clone().requires_grad_(True)
    rr, rw = REF.thetas(sh32, sc32)                          # fp32: the forward must be bit-identical (fp32 divides)
    assert torch.equal(tr, rr) and torch.equal(tw, rw)
    sh64, sc64 = sh.detach().double().requires_grad_(True), sc.detach().double().requires_grad_(True)
    r64, w64 = REF.thetas(sh64, sc64)
    ((r64 * gr.double()).sum() + (w64 * gw.double()).sum()).backward()
    _check(sh.grad, sh64.grad, rtol=1e-5, atol=1e-5)
    _check(sc.grad.reshape(-1), sc64.grad.reshape(-1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("temp", [0.1, 1.0])
def test_zpres(cuda_device, temp):
    g = torch.Generator(device=cuda_device).manual_seed(3)
    B = 513
    lo = (torch.randn(B, device=cuda_device, generator=g) * 2).requires_grad_(True)
    u = torch.rand(B, device=cuda_device, generator=g).clamp(1e-4, 1 - 1e-4)
    stop = torch.rand(B, device=cuda_device, generator=g) * 1.5
    gy, gz = torch.randn(B, device=cuda_device, generator=g), torch.randn(B, device=cuda_device, generator=g)
    y, z, s1, ap, ac = fused.zpres(lo, u, stop, temp, 0.9)
    ((y * gy).sum() + (z * gz).sum()).backward()
    lo64 = lo.detach().double().requires_grad_(True)
    ry, rz, rs, rap, rac = REF.zpres(lo64, u.double(), stop.double(), temp, 0.9)
    ((ry * gy.double()).sum() + (rz * gz.double()).sum()).backward()
    _check(y, ry, rtol=1e-5, atol=1e-5)
    _check(z, rz, rtol=1e-5, atol=1e-6)
    _check(s1, rs, rtol=1e-6, atol=1e-6)
    assert torch.equal(ap, rap)
    close = (rs - 0.9).abs() < 1e-5                         # the mask may differ only where stop_sum sits on the threshold
    assert torch.equal(ac[~close], rac[~close])
    _check(lo.grad, lo64.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("split", [False, True])
def test_lstm_pointwise(cuda_device, split):
    """``split``: the gate pre-activations arrive as two addends (per-step GEMM + the step-invariant part)"""
    g = torch.Generator(device=cuda_device).manual_seed(4)
    B, Hh = 37, 256
    gates = torch.randn((B, 4 * Hh), device=cuda_device, generator=g).requires_grad_(True)
    gates2 = torch.randn((B, 4 * Hh), device=cuda_device, generator=g).requires_grad_(True) if split else None
    c = torch.randn((B, Hh), device=cuda_device, generator=g).requires_grad_(True)
    gc, gh = torch.randn((B, Hh), device=cuda_device, generator=g), torch.randn((B, Hh), device=cuda_device, generator=g)
    c2, h2 = fused.lstm_pointwise(gates, c, gates2)
    ((c2 * gc).sum() + (h2 * gh).sum()).backward()
    total = gates.detach().double() + (gates2.detach().double() if split else 0.0)
    g64, c64 = total.requires_grad_(True), c.detach().double().requires_grad_(True)
    i, j, f, o = g64.chunk(4, 1)                                   # LSTMCellTF.forward, written out
    rc = torch.sigmoid(f + 1.0) * c64 + torch.sigmoid(i) * torch.tanh(j)
    rh = torch.sigmoid(o) * torch.tanh(rc)
    ((rc * gc.double()).sum() + (rh * gh.double()).sum()).backward()
    _check(c2, rc, rtol=1e-5, atol=1e-6)
    _check(h2, rh, rtol=1e-5, atol=1e-6)
    _check(gates.grad, g64.grad, rtol=1e-5, atol=1e-6)
    if split:
        _check(gates2.grad, g64.grad, rtol=1e-5, atol=1e-6)
    _check(c.grad, c64.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("T", [6, 3])
def test_kl_terms(cuda_device, T):
    """fused KL kernel vs the model's framework-op formulas (AIRModel._concrete_kl / _gaussian_kls) in fp64"""
    from mog_asr_b200.air import AIRModel, config_from_flags
    cfg = config_from_flags("mnist", "13", gm=100.0, gne=10.0)
    model = AIRModel(cfg, ops=REF)
    g = torch.Generator(device=cuda_device).manual_seed(6)
    B, Ld = 65, cfg.vae_latent_dimensions
    shp = dict(y_pre=(T, B), prior_lo=(T, B), post_lo=(T, B), sc_mean=(T, B, 1), sc_lv=(T, B, 1), sh_mean=(T, B, 2), sh_lv=(T, B, 2),
               g_sh_mean=(T, B, 2), g_sh_lv=(T, B, 2), v_mean=(T, B, Ld), v_lv=(T, B, Ld))
    st = {k: (torch.randn(v, device=cuda_device, generator=g) * (3.0 if k == "y_pre" else 0.8)).requires_grad_(True) for k, v in shp.items()}
    st["active_prev"] = torch.rand((T, B), device=cuda_device, generator=g) < 0.7
    st["active"] = torch.rand((T, B), device=cuda_device, generator=g) < 0.6
    gk = torch.randn(B, device=cuda_device, generator=g)
    kl, comps = fused.kl_terms(st, cfg.z_pres_temperature, cfg.scale_prior_mean, cfg.scale_prior_variance, cfg.vae_prior_mean, cfg.vae_prior_variance)
    (kl * gk).sum().backward()
    s64 = {k: (v.detach().double().requires_grad_(True) if v.dtype != torch.bool else v) for k, v in st.items()}
    z_kl = model._concrete_kl(s64["y_pre"], s64["prior_lo"], s64["post_lo"], cfg.z_pres_temperature)
    sck, shk, vk = model._gaussian_kls(s64["sc_mean"], s64["sc_lv"], s64["sh_mean"], s64["sh_lv"], s64["g_sh_mean"], s64["g_sh_lv"],
                                       s64["v_mean"], s64["v_lv"])
    zero = torch.zeros_like(z_kl)
    parts = [torch.where(s64["active_prev"], z_kl, zero).sum(0)] + [torch.where(s64["active"], t, zero).sum(0) for t in (sck, shk, vk)]
    ref = sum(parts)
    (ref * gk.double()).sum().backward()
    _check(kl, ref, rtol=2e-5, atol=1e-4)
    for i, p_ in enumerate(parts):
        _check(comps[:, i], p_, rtol=2e-5, atol=1e-4)
    for k in shp:
        _check(st[k].grad, s64[k].grad, rtol=2e-5, atol=1e-5)


@pytest.mark.parametrize("skip_dim,out_dim,act,B", [(0, 2, "tanh", 300), (2, 1, "sigmoid", 257), (0, 2, "tanh", 3), (2, 1, "sigmoid", 4096)])
def test_fused_head_matches_framework_ops(cuda_device, skip_dim, out_dim, act, B):
    """The (mean, log-variance) heads with their sample (:424-460): one library GEMM + one kernel each way against the
    layer-by-layer framework form -- outputs, input gradients and every parameter gradient (second layer: accumulated by
    the backward kernel; first layer: from the stashed rows at flush time), over two loop iterations sharing the head."""
    from mog_asr_b200.air.model import _MeanVar, CudaOps
    dev = cuda_device

    def close(a, b, rtol, atol):
        return bool(torch.all((a - b).abs() <= atol + rtol * b.abs()))
    torch.manual_seed(skip_dim * 10 + out_dim)
    head = _MeanVar(256, 64, out_dim, skip_dim=skip_dim).to(dev)
    for prm in head.parameters():                       # zero-initialised biases would hide bias-path errors
        prm.data.add_(0.05 * torch.randn_like(prm))
    ops = CudaOps()
    xs = [torch.randn(B, 256, device=dev, requires_grad=True) for _ in range(2)]
    sk = [torch.randn(B, skip_dim, device=dev, requires_grad=True) if skip_dim else None for _ in range(2)]
    eps = [torch.randn(B, out_dim, device=dev) for _ in range(2)]
    wts = [torch.randn(4, B, out_dim, device=dev) for _ in range(2)]

    def run(fused):
        for t in xs + [s for s in sk if s is not None]:
            t.grad = None
        for prm in head.parameters():
            prm.grad = torch.zeros_like(prm)
        for lyr in (head.hm, head.m, head.hv, head.v):
            lyr.defer = fused
            lyr._stash.clear()
        head._head_stash.clear()
        head._w1cat = None
        if fused:
            assert head.fusable(ops, xs[0])
            head.prepare()
        outs, loss = [], 0.0
        for x, s, e, w in zip(xs, sk, eps, wts):
            o = head.sample(ops, x, e, act, skip=s)
            outs.append([t.detach().clone() for t in o])
            loss = loss + sum((wi * oi).sum() for wi, oi in zip(w, o))
        loss.backward()
        if fused:
            head.flush_head()
            assert not any(lyr._stash for lyr in (head.hm, head.m, head.hv, head.v))
        grads = [x.grad.clone() for x in xs] + [s.grad.clone() for s in sk if s is not None]
        return outs, grads, {n: p.grad.clone() for n, p in head.named_parameters()}

    o0, g0, p0 = run(False)
    o1, g1, p1 = run(True)
    for a, b in zip(o0, o1):
        for x, y in zip(a, b):
            assert close(y, x, 2e-5, 2e-5)
    for x, y in zip(g0, g1):
        assert close(y, x, 1e-4, 1e-4 * float(x.abs().max()))
    for n in p0:
        assert close(p1[n], p0[n], 2e-4, 2e-4 * float(p0[n].abs().max()) + 1e-6), n


@pytest.mark.parametrize("act", [None, "relu", "softplus", "sigmoid"])
def test_fused_dense_epilogue_matches_framework_ops(cuda_device, act):
    """act(x W^T + b) as library GEMM + one kernel each way, weight / bias gradients through the deferred stash."""
    from mog_asr_b200.air.model import StepLinear
    torch.manual_seed(3)
    lyr = StepLinear(96, 40).to(cuda_device)
    lyr.bias.data.normal_(0, 0.3)
    xs = [(3.0 * torch.randn(70, 96, device=cuda_device)).requires_grad_(True) for _ in range(2)]
    ws = [torch.randn(70, 40, device=cuda_device) for _ in range(2)]
    res = []
    for fused in (False, True):
        lyr.fuse, lyr.defer = fused, fused
        lyr._stash.clear()
        lyr.weight.grad, lyr.bias.grad = torch.zeros_like(lyr.weight), torch.zeros_like(lyr.bias)
        for x in xs:
            x.grad = None
        ys = [lyr(x, act) for x in xs]
        sum((y * w).sum() for y, w in zip(ys, ws)).backward()
        if fused:
            assert len(lyr._stash) == 2
            lyr.flush_grads()
        res.append(([y.detach().clone() for y in ys], [x.grad.clone() for x in xs], lyr.weight.grad.clone(), lyr.bias.grad.clone()))
    (y0, g0, w0, b0), (y1, g1, w1, b1) = res
    for a, b in zip(y0 + g0 + [w0, b0], y1 + g1 + [w1, b1]):
        assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max()))


def test_fused_mean_logvar_pair_matches_framework_ops(cuda_device):
    """vae.py:21-31: mean and log-variance layers on one GEMM, biases + sample in one kernel, gradients at flush time."""
    from mog_asr_b200.air.model import StepLinear, _GaussPair, CudaOps
    torch.manual_seed(4)
    lm, lv = StepLinear(256, 50).to(cuda_device), StepLinear(256, 50).to(cuda_device)
    for l in (lm, lv):
        l.bias.data.normal_(0, 0.2)
    pair, ops = _GaussPair(lm, lv), CudaOps()
    xs = [torch.randn(130, 256, device=cuda_device, requires_grad=True) for _ in range(2)]
    eps = [torch.randn(130, 50, device=cuda_device) for _ in range(2)]
    ws = [torch.randn(3, 130, 50, device=cuda_device) for _ in range(2)]
    res = []
    for fused in (False, True):
        for l in (lm, lv):
            l.fuse, l.defer = fused, fused
            l._stash.clear()
            l.weight.grad, l.bias.grad = torch.zeros_like(l.weight), torch.zeros_like(l.bias)
        pair.reset()
        for x in xs:
            x.grad = None
        outs, loss = [], 0.0
        for x, e, w in zip(xs, eps, ws):
            if fused:
                pair.prepare()
                assert pair.usable(x)
                o = fused_mod.linear_gauss(pair, x, e)
            else:
                m, v = lm(x), lv(x)
                o = (m, v, ops.gauss_sample(m, v, e)[0])
            outs.append([t.detach().clone() for t in o])
            loss = loss + sum((wi * oi).sum() for wi, oi in zip(w, o))
        loss.backward()
        if fused:
            pair.flush()
        res.append((outs, [x.grad.clone() for x in xs], [p.grad.clone() for l in (lm, lv) for p in (l.weight, l.bias)]))
    (o0, g0, p0), (o1, g1, p1) = res
    flat = lambda o: [t for step in o for t in step]
    for a, b in zip(flat(o0) + g0 + p0, flat(o1) + g1 + p1):
        assert float((a - b).abs().max()) <= 3e-5 * max(1.0, float(a.abs().max()))


def test_thetas_kernel_vs_the_reference_source_run_on_the_tf_shim(cuda_device):
    """Directly against ``tests/golden/graph_thetas.npz`` (the reference's st_forward / st_backward lines on the torch TF
    shim, float32): both matrices bit for bit, gradients to rounding."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_thetas.npz"))
    s = torch.tensor(g["s"][:, None], device=cuda_device, requires_grad=True)
    xy = torch.tensor(np.stack([g["x"], g["y"]], 1), device=cuda_device, requires_grad=True)
    th_r, th_w = fused.thetas(xy, s)
    assert np.array_equal(th_r.detach().cpu().numpy().reshape(-1, 2, 3).view(np.uint32), g["theta"].view(np.uint32))
    assert np.array_equal(th_w.detach().cpu().numpy().reshape(-1, 2, 3).view(np.uint32), g["theta_recon"].view(np.uint32))
    ((th_r.reshape(-1, 2, 3) * torch.tensor(g["wr"], device=cuda_device)).sum()
     + (th_w.reshape(-1, 2, 3) * torch.tensor(g["ww"], device=cuda_device)).sum()).backward()
    np.testing.assert_allclose(s.grad.cpu().numpy()[:, 0], g["ds"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(xy.grad.cpu().numpy()[:, 0], g["dx"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(xy.grad.cpu().numpy()[:, 1], g["dy"], rtol=2e-5, atol=1e-5)
