"""Fused per-step elementwise math of the AIR loop body (csrc/mog_air.cu): each call is one
kernel forward and one backward instead of ~10 / ~20 framework ops.  See include/mogstn.h for the formulas and the
reference lines they replace."""
from __future__ import annotations

import torch

from .. import _lib
from ..transformer import _need_cuda, _stream

_ACT = {None: 0, "none": 0, "tanh": 1, "sigmoid": 2}


def _p(t):
    return t.data_ptr() if t is not None else None


class _GaussSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, logvar, eps, act):
        L = _lib.load()
        latent = torch.empty_like(mean)
        squashed = torch.empty_like(mean) if act else None
        with torch.cuda.device(mean.device):
            _lib.check(L.mog_air_gauss_sample_forward(_p(mean), _p(logvar), _p(eps), _p(latent), _p(squashed), mean.numel(),
                                                      act, _stream(mean)), "mog_air_gauss_sample_forward")
        ctx.save_for_backward(logvar, eps, squashed)
        ctx.act = act
        ctx.set_materialize_grads(False)     # an unused output's gradient arrives as None, not as a zero-filled tensor (one fill kernel each)
        return (latent, squashed) if act else (latent, latent.new_empty(0))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_latent, g_squashed):
        logvar, eps, squashed = ctx.saved_tensors
        if g_latent is None and (g_squashed is None or not ctx.act):
            return None, None, None, None
        L = _lib.load()
        g_latent = g_latent.contiguous() if g_latent is not None else None
        g_squashed = g_squashed.contiguous() if (ctx.act and g_squashed is not None) else None
        d_mean, d_logvar = torch.empty_like(logvar), torch.empty_like(logvar)
        with torch.cuda.device(logvar.device):
            _lib.check(L.mog_air_gauss_sample_backward(_p(logvar), _p(eps), _p(squashed), _p(g_latent), _p(g_squashed),
                                                       _p(d_mean), _p(d_logvar), logvar.numel(), ctx.act, _stream(logvar)),
                       "mog_air_gauss_sample_backward")
        return d_mean, d_logvar, None, None


def gauss_sample(mean, logvar, eps, act=None):
    """``latent = mean + eps*sqrt(exp(logvar))`` and ``act(latent)`` (``act`` in None/'tanh'/'sigmoid').
    Returns ``(latent, squashed)``; ``squashed`` is None when ``act`` is None."""
    for t, n in ((mean, "mean"), (logvar, "logvar"), (eps, "eps")):
        _need_cuda(t, n)
    a = _ACT[act]
    latent, squashed = _GaussSample.apply(mean.float().contiguous(), logvar.float().contiguous(), eps.float().contiguous(), a)
    return latent, (squashed if a else None)


class _Thetas(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shift, scale):
        L = _lib.load()
        B = shift.shape[0]
        th_r = torch.empty((B, 6), dtype=torch.float32, device=shift.device)
        th_w = torch.empty((B, 6), dtype=torch.float32, device=shift.device)
        with torch.cuda.device(shift.device):
            _lib.check(L.mog_air_thetas_forward(_p(shift), _p(scale), _p(th_r), _p(th_w), B, _stream(shift)), "mog_air_thetas_forward")
        ctx.save_for_backward(shift, scale)
        ctx.set_materialize_grads(False)
        return th_r, th_w

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_r, g_w):
        shift, scale = ctx.saved_tensors
        L = _lib.load()
        B = shift.shape[0]
        d_shift, d_scale = torch.empty_like(shift), torch.empty_like(scale)
        if g_r is None and g_w is None:
            return None, None
        g_r = torch.zeros((B, 6), dtype=torch.float32, device=shift.device) if g_r is None else g_r
        g_w = torch.zeros((B, 6), dtype=torch.float32, device=shift.device) if g_w is None else g_w
        with torch.cuda.device(shift.device):
            _lib.check(L.mog_air_thetas_backward(_p(shift), _p(scale), _p(g_r.contiguous()), _p(g_w.contiguous()), _p(d_shift),
                                                 _p(d_scale), B, _stream(shift)), "mog_air_thetas_backward")
        return d_shift, d_scale


def thetas(inf_shift, inf_scale):
    """``inf_shift [B,2]`` (tanh of the shift latent), ``inf_scale [B,1]`` -> ``(theta_read [B,6], theta_write [B,6])``."""
    _need_cuda(inf_shift, "inf_shift")
    _need_cuda(inf_scale, "inf_scale")
    return _Thetas.apply(inf_shift.float().contiguous(), inf_scale.float().reshape(-1).contiguous())


class _ZPres(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_odds, u, stop_sum, temperature, threshold):
        L = _lib.load()
        B = log_odds.shape[0]
        dev = log_odds.device
        y, z, s1 = torch.empty_like(log_odds), torch.empty_like(log_odds), torch.empty_like(log_odds)
        ap = torch.empty(B, dtype=torch.bool, device=dev)
        ac = torch.empty(B, dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.mog_air_zpres_forward(_p(log_odds), _p(u), _p(stop_sum), float(temperature), float(threshold), _p(y), _p(z),
                                               _p(s1), _p(ap), _p(ac), B, _stream(log_odds)), "mog_air_zpres_forward")
        ctx.save_for_backward(z)
        ctx.temperature = float(temperature)
        ctx.mark_non_differentiable(s1, ap, ac)
        ctx.set_materialize_grads(False)
        return y, z, s1, ap, ac

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_y, g_z, _gs, _gap, _gac):
        (z,) = ctx.saved_tensors
        if g_y is None and g_z is None:
            return None, None, None, None, None
        L = _lib.load()
        d_lo = torch.empty_like(z)
        g_y = g_y.contiguous() if g_y is not None else None
        g_z = g_z.contiguous() if g_z is not None else None
        with torch.cuda.device(z.device):
            _lib.check(L.mog_air_zpres_backward(_p(z), _p(g_y), _p(g_z), ctx.temperature, _p(d_lo), z.shape[0], _stream(z)),
                       "mog_air_zpres_backward")
        return d_lo, None, None, None, None


def zpres(log_odds, u, stop_sum, temperature, threshold):
    """Concrete sample + stopping-sum update: returns ``(y_pre, z_pres, stop_sum_new, active_prev, active)``."""
    for t, n in ((log_odds, "log_odds"), (u, "u"), (stop_sum, "stop_sum")):
        _need_cuda(t, n)
    return _ZPres.apply(log_odds.float().contiguous(), u.float().contiguous(), stop_sum.detach().float().contiguous(), temperature, threshold)


class _LstmPointwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gates, c_prev, gates2):
        L = _lib.load()
        B, H = c_prev.shape
        c_new, h_new = torch.empty_like(c_prev), torch.empty_like(c_prev)
        with torch.cuda.device(gates.device):
            _lib.check(L.mog_air_lstm_pointwise_forward(_p(gates), _p(gates2), _p(c_prev), _p(c_new), _p(h_new), B, H, _stream(gates)),
                       "mog_air_lstm_pointwise_forward")
        ctx.save_for_backward(gates, c_prev, c_new, gates2)
        ctx.set_materialize_grads(False)
        return c_new, h_new

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_c, g_h):
        gates, c_prev, c_new, gates2 = ctx.saved_tensors
        if g_c is None and g_h is None:
            return None, None, None
        L = _lib.load()
        B, H = c_prev.shape
        g_c = g_c.contiguous() if g_c is not None else None
        g_h = g_h.contiguous() if g_h is not None else None
        d_gates, d_c_prev = torch.empty_like(gates), torch.empty_like(c_prev)
        with torch.cuda.device(gates.device):
            _lib.check(L.mog_air_lstm_pointwise_backward(_p(gates), _p(gates2), _p(c_prev), _p(c_new), _p(g_h), _p(g_c), _p(d_gates),
                                                         _p(d_c_prev), B, H, _stream(gates)), "mog_air_lstm_pointwise_backward")
        return d_gates, d_c_prev, (d_gates if gates2 is not None else None)


def lstm_pointwise(gates, c_prev, gates2=None):
    """``gates (+ gates2) [B,4H]`` (i, j, f, o) and ``c_prev [B,H]`` -> ``(c_new, h_new)`` with ``forget_bias = 1``."""
    _need_cuda(gates, "gates")
    _need_cuda(c_prev, "c_prev")
    return _LstmPointwise.apply(gates.float().contiguous(), c_prev.float().contiguous(),
                                None if gates2 is None else gates2.float().contiguous())


_KL_KEYS = ("y_pre", "prior_lo", "post_lo", "sc_mean", "sc_lv", "sh_mean", "sh_lv", "g_sh_mean", "g_sh_lv", "v_mean", "v_lv")


class _KlTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, active_prev, active, consts, *tensors):
        L = _lib.load()
        y_pre = tensors[0]
        T, B = y_pre.shape
        Ld = tensors[9].shape[-1]
        temp, scm, scv, vm, vv = consts
        kl = torch.empty(B, dtype=torch.float32, device=y_pre.device)
        comps = torch.empty((B, 4), dtype=torch.float32, device=y_pre.device)
        t = dict(zip(_KL_KEYS, tensors))
        with torch.cuda.device(y_pre.device):
            _lib.check(L.mog_air_kl_forward(
                _p(t["y_pre"]), _p(t["prior_lo"]), _p(t["post_lo"]), _p(active_prev), _p(active), _p(t["sc_mean"]), _p(t["sc_lv"]),
                _p(t["sh_mean"]), _p(t["sh_lv"]), _p(t["g_sh_mean"]), _p(t["g_sh_lv"]), _p(t["v_mean"]), _p(t["v_lv"]),
                B, T, Ld, temp, scm, scv, vm, vv, _p(kl), _p(comps), _stream(y_pre)), "mog_air_kl_forward")
        ctx.save_for_backward(active_prev, active, *tensors)
        ctx.consts, ctx.dims = consts, (B, T, Ld)
        ctx.mark_non_differentiable(comps)
        return kl, comps

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_kl, _g_comps):
        active_prev, active, *tensors = ctx.saved_tensors
        L = _lib.load()
        B, T, Ld = ctx.dims
        temp, scm, scv, vm, vv = ctx.consts
        t = dict(zip(_KL_KEYS, tensors))
        d = {k: torch.empty_like(v) for k, v in t.items()}
        g_kl = g_kl.to(torch.float32).contiguous()
        with torch.cuda.device(g_kl.device):
            _lib.check(L.mog_air_kl_backward(
                _p(t["y_pre"]), _p(t["prior_lo"]), _p(t["post_lo"]), _p(active_prev), _p(active), _p(t["sc_mean"]), _p(t["sc_lv"]),
                _p(t["sh_mean"]), _p(t["sh_lv"]), _p(t["g_sh_mean"]), _p(t["g_sh_lv"]), _p(t["v_mean"]), _p(t["v_lv"]),
                B, T, Ld, temp, scm, scv, vm, vv, _p(g_kl),
                _p(d["y_pre"]), _p(d["prior_lo"]), _p(d["post_lo"]), _p(d["sc_mean"]), _p(d["sc_lv"]), _p(d["sh_mean"]), _p(d["sh_lv"]),
                _p(d["g_sh_mean"]), _p(d["g_sh_lv"]), _p(d["v_mean"]), _p(d["v_lv"]), _stream(g_kl)), "mog_air_kl_backward")
        return (None, None, None) + tuple(d[k] for k in _KL_KEYS)


def kl_terms(stacks, temperature, scale_prior_mean, scale_prior_var, vae_prior_mean, vae_prior_var):
    """``stacks``: dict of ``[T,B,..]`` tensors with the keys of ``_KL_KEYS`` plus boolean ``active_prev`` / ``active``.
    Returns ``(kl [B], components [B,4])`` -- the masked sum over steps of the four KL terms and the four separate sums."""
    ts = []
    for k in _KL_KEYS:
        v = stacks[k]
        _need_cuda(v, k)
        v = v.to(torch.float32)
        if k in ("sc_mean", "sc_lv"):
            v = v.reshape(v.shape[0], v.shape[1])
        ts.append(v.contiguous())
    ap = stacks["active_prev"].to(torch.bool).contiguous()
    ac = stacks["active"].to(torch.bool).contiguous()
    consts = (float(temperature), float(scale_prior_mean), float(scale_prior_var), float(vae_prior_mean), float(vae_prior_var))
    return _KlTerms.apply(ap, ac, consts, *ts)


class _FusedHead(torch.autograd.Function):
    """One (mean, log-variance) head with its sample: first-layer GEMM by the library, everything else in one kernel
    each way (csrc/mog_air_head.cu).  The parameters are NOT autograd inputs: second-layer gradients are accumulated
    into their ``.grad`` by the backward kernel, first-layer gradients are formed once per training step from the
    stashed ``(x, skip, dpre1)`` rows (``_MeanVar.flush_head``) -- the deferred-weight-gradient scheme of the loop."""

    @staticmethod
    def forward(ctx, x, skip, eps, w1cat, head, act, _anchor):
        # ``_anchor`` (a parameter of the head) only makes sure the node exists when no data input requires grad
        L = _lib.load()
        hm, m, hv, v = head.hm, head.m, head.hv, head.v
        B, K = x.shape
        h, O = hm.weight.shape[0], m.weight.shape[0]
        S = 0 if skip is None else skip.shape[1]
        pre1 = torch.mm(x, w1cat.t())                                   # [B, 2h] = x [W1m | W1v]
        new = lambda: torch.empty((B, O), dtype=torch.float32, device=x.device)
        mean, logvar, latent, squashed = new(), new(), new(), (new() if act else None)
        with torch.cuda.device(x.device):
            _lib.check(L.mog_air_head_forward(_p(pre1), _p(skip), _p(eps), _p(hm.weight), _p(hm.bias), _p(hv.weight), _p(hv.bias),
                                              _p(m.weight), _p(m.bias), _p(v.weight), _p(v.bias), B, h, K, S, O, act,
                                              _p(mean), _p(logvar), _p(latent), _p(squashed), _stream(x)), "mog_air_head_forward")
        ctx.save_for_backward(x, skip, eps, pre1, logvar, squashed, w1cat)
        ctx.head, ctx.act, ctx.dims = head, act, (B, h, K, S, O)
        ctx.set_materialize_grads(False)
        return mean, logvar, latent, (squashed if act else latent.new_empty(0))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_mean, g_logvar, g_latent, g_squashed):
        x, skip, eps, pre1, logvar, squashed, w1cat = ctx.saved_tensors
        head, act, (B, h, K, S, O) = ctx.head, ctx.act, ctx.dims
        hm, m, hv, v = head.hm, head.m, head.hv, head.v
        L = _lib.load()
        c = lambda g: g.contiguous() if g is not None else None
        g_mean, g_logvar, g_latent = c(g_mean), c(g_logvar), c(g_latent)
        g_squashed = c(g_squashed) if act else None
        for prm in (m.weight, m.bias, v.weight, v.bias):
            if prm.grad is None:
                prm.grad = torch.zeros_like(prm)
        dpre1 = torch.empty_like(pre1)
        dskip = torch.empty_like(skip) if skip is not None else None
        with torch.cuda.device(x.device):
            _lib.check(L.mog_air_head_backward(_p(pre1), _p(skip), _p(eps), _p(hm.weight), _p(hm.bias), _p(hv.weight), _p(hv.bias),
                                               _p(m.weight), _p(v.weight), _p(logvar), _p(squashed), _p(g_mean), _p(g_logvar),
                                               _p(g_latent), _p(g_squashed), B, h, K, S, O, act, _p(dpre1), _p(dskip),
                                               _p(m.weight.grad), _p(m.bias.grad), _p(v.weight.grad), _p(v.bias.grad), _stream(x)),
                       "mog_air_head_backward")
        head._head_stash.append((x, skip, dpre1))
        dx = torch.mm(dpre1, w1cat) if ctx.needs_input_grad[0] else None
        return dx, dskip, None, None, None, None, None


def fused_head(head, x, skip, eps, act=None):
    """``(mean, logvar, latent, squashed)`` of a ``_MeanVar`` head (``head.prepare()`` must have run in this step)."""
    _need_cuda(x, "x")
    a = _ACT[act]
    mean, logvar, latent, squashed = _FusedHead.apply(x.float().contiguous(), None if skip is None else skip.float().contiguous(),
                                                      eps.float().contiguous(), head._w1cat, head, a, head.m.weight)
    return mean, logvar, latent, (squashed if a else None)


_ACT_DENSE = {None: 0, "none": 0, "relu": 1, "softplus": 2, "sigmoid": 3}


class _FusedLinearAct(torch.autograd.Function):
    """``act(x W^T + b)``: library GEMM + one bias/activation kernel; backward: one activation kernel + library GEMM.
    The layer's weight / bias gradients go through its deferred stash (``(x, dpre)`` rows, one GEMM per training step)."""

    @staticmethod
    def forward(ctx, x, w, bias, layer, act):
        # ``w`` / ``bias`` are autograd inputs on purpose (their gradients are returned as None and travel through the
        # layer's stash instead): with them the node exists even when ``x`` does not require grad -- the first loop
        # iteration feeds an all-zero state to the z_pres prior head (air_number_bbox_location.py:609-615), and without a
        # node that iteration's weight-gradient rows would be lost.
        L = _lib.load()
        y = torch.mm(x, w.t())
        B, N = y.shape
        with torch.cuda.device(x.device):
            _lib.check(L.mog_air_bias_act_forward(_p(y), _p(bias), _p(y), B, N, act, _stream(x)), "mog_air_bias_act_forward")
        ctx.save_for_backward(x, y)
        ctx.layer, ctx.act = layer, act
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        layer, act = ctx.layer, ctx.act
        g = g.contiguous()
        if act:
            L = _lib.load()
            dpre = torch.empty_like(g)
            with torch.cuda.device(x.device):
                _lib.check(L.mog_air_bias_act_backward(_p(y), _p(g), _p(dpre), g.numel(), act, _stream(x)), "mog_air_bias_act_backward")
        else:
            dpre = g
        layer._stash.append((x, dpre))
        dx = torch.mm(dpre, layer.weight) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None


def linear_act(layer, x, act=None):
    _need_cuda(x, "x")
    return _FusedLinearAct.apply(x.float().contiguous(), layer.weight, layer.bias, layer, _ACT_DENSE[act])


class _FusedLinearGauss(torch.autograd.Function):
    """mean / log-variance layers sharing their input, plus the sample: one GEMM on ``[Wmean | Wlogvar]`` and one kernel."""

    @staticmethod
    def forward(ctx, x, eps, pair, _anchor):
        L = _lib.load()
        pre = torch.mm(x, pair.wcat.t())                                  # [B, 2L]
        B, Ld = x.shape[0], pair.mean_layer.weight.shape[0]
        new = lambda: torch.empty((B, Ld), dtype=torch.float32, device=x.device)
        mean, logvar, latent = new(), new(), new()
        with torch.cuda.device(x.device):
            _lib.check(L.mog_air_bias_gauss_forward(_p(pre), _p(pair.mean_layer.bias), _p(pair.logvar_layer.bias), _p(eps), _p(mean),
                                                    _p(logvar), _p(latent), B, Ld, _stream(x)), "mog_air_bias_gauss_forward")
        ctx.save_for_backward(x, eps, logvar, pair.wcat)
        ctx.pair = pair
        ctx.set_materialize_grads(False)
        return mean, logvar, latent

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_mean, g_logvar, g_latent):
        x, eps, logvar, wcat = ctx.saved_tensors
        L = _lib.load()
        c = lambda g: g.contiguous() if g is not None else None
        B, Ld = logvar.shape
        dpre = torch.empty((B, 2 * Ld), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.mog_air_bias_gauss_backward(_p(logvar), _p(eps), _p(c(g_mean)), _p(c(g_logvar)), _p(c(g_latent)), _p(dpre), B, Ld,
                                                     _stream(x)), "mog_air_bias_gauss_backward")
        ctx.pair.stash.append((x, dpre))
        dx = torch.mm(dpre, wcat) if ctx.needs_input_grad[0] else None
        return dx, None, None, None


def linear_gauss(pair, x, eps):
    _need_cuda(x, "x")
    return _FusedLinearGauss.apply(x.float().contiguous(), eps.float().contiguous(), pair, pair.mean_layer.weight)
