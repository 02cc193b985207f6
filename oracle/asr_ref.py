"""ORACLE (test infrastructure, NOT product code) -- the ASR regularisers of AIR-ASR restated op for op.

Follows ``/root/reference/air/air_number_bbox_location.py``:
  * per-step entropy term ``pr_num``            ``:645-681`` (summed over steps ``:937-943``)
  * marginal + min-element count penalties      ``:970-1015``
  * size-window ("area") penalty                ``:1016-1027``
  * out-of-canvas, pairwise size, pairwise overlap ``:1029-1069``
  * how they enter the loss                     ``:1078-1079``

**Parity pinned to the reference's own lines** (not to a live TensorFlow): ``tests/golden/make_golden_asr_graph.py`` reads
those lines from the reference file, exec's them on a torch-based stand-in for the TF ops they use and differentiates
with autograd; ``asr_terms`` reproduces values, logged components and all gradients to 1e-10
(``tests/test_oracle.py``).  Still assumed ([TF-1.12 assumed]): the sub-gradient tie rules -- ``maximum(a,b)`` routes the
gradient to ``a`` when ``a >= b``; ``abs'(0) = 0``; ``reduce_min`` splits equally among ties.  torch autograd is steered to
the same rules below.

Written with torch-CPU ops so autograd provides the gradients (the role TF autodiff plays at ``:1098``).
"""
from __future__ import annotations

import numpy as np
import torch


def _tf_max0(v: torch.Tensor) -> torch.Tensor:
    """``tf.maximum(v, zeros_like(v))`` with TF's tie rule: gradient goes to ``v`` when ``v >= 0``."""
    return torch.where(v >= 0, v, torch.zeros_like(v))


def _bce_logits(labels: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    """``tf.nn.sigmoid_cross_entropy_with_logits`` as TF-1.12 builds it: ``cond = x >= 0``,
    ``relu = where(cond, x, 0)``, ``neg_abs = where(cond, -x, x)``, ``relu - x*z + log1p(exp(neg_abs))``."""
    cond = logits >= 0
    relu = torch.where(cond, logits, torch.zeros_like(logits))
    neg_abs = torch.where(cond, -logits, logits)
    return relu - logits * labels + torch.log1p(torch.exp(neg_abs))


def asr_terms(log_odds, shifts, scales, *, canvas_size, counts, max_steps, gamma_num=0.0, gamma_margin=0.0,
              gamma_elem=0.0, gamma_bbox=0.0, gamma_size=0.0, gamma_area=0.0, area_minmax=(0.0, 0.0),
              batch_prob_mean=None):
    """Torch tensors in, dict of torch tensors out (differentiable).

    ``log_odds [B,T]`` posterior z_pres log-odds per executed step (``P = sigmoid(log_odds)``, ``:641-643``);
    ``shifts [B,T,2]`` = tanh(shift latent) (``:435-437``); ``scales [B,T]`` = sigmoid(scale latent)
    (``:458-460``).  ``counts`` = digits of ``-dn`` (train_air_pr.py:67,204).  ``area_minmax = (size_min,
    size_max)`` (train_air_pr.py:83-97,211).  ``batch_prob_mean [T]`` overrides ``mean_b P`` (data-parallel
    runs pass the global mean).

    Returns ``per_image [B]`` (= pr_loss + L_elem, the part inside ``reduce_mean`` at ``:1078``),
    ``margin`` scalar (added once, ``:1079``) and the logged components.
    """
    B, T = log_odds.shape
    dt = log_odds.dtype
    cs = float(canvas_size)
    P = torch.sigmoid(log_odds)

    # ---- pr_num (:645-681): entropy of the posterior Bernoulli, gated by gamma_num > 1e-8
    if gamma_num > 1e-8:
        ent = P * torch.nn.functional.softplus(-log_odds) + (1.0 - P) * torch.nn.functional.softplus(log_odds)
        pr_num = (ent * gamma_num).sum(-1)
    else:
        pr_num = torch.zeros(B, dtype=dt, device=log_odds.device)

    # ---- count penalties (:970-1015); both gated by gamma_margin > 1e-8 (:973)
    dev = log_odds.device
    margin = torch.zeros((), dtype=dt, device=dev)
    elem = torch.zeros(B, dtype=dt, device=dev)
    if gamma_margin > 1e-8:
        K = len(counts)
        obj = torch.zeros(K, max_steps, dtype=dt, device=dev)
        for k, c in enumerate(counts):
            obj[k, :c] = 1.0                                               # :974-976
        mobj = obj.mean(0)                                                 # :979
        pbar = P.mean(0) if batch_prob_mean is None else batch_prob_mean   # :982
        logit_bar = torch.log(pbar + 1e-8) - torch.log(1 - pbar + 1e-8)    # :987-988
        margin = (_bce_logits(mobj[:T], logit_bar) * gamma_margin).sum()   # :985-991
        logit_p = torch.log(P + 1e-8) - torch.log(1 - P + 1e-8)            # :999-1000
        bce = _bce_logits(obj[None, :, :T], logit_p[:, None, :]).sum(-1)   # :997-1005  [B,K]
        elem = torch.amin(bce, dim=-1) * gamma_elem                        # :1006-1010 (ties split equally)

    # ---- size window (:1016-1027)
    px = scales * cs                                                       # [B,T]
    area = (_tf_max0(area_minmax[1] - px) + _tf_max0(px - area_minmax[0])).mean(-1)

    # ---- boxes (:1029-1069)
    cx = (shifts[..., 0] + 1.0) * cs / 2.0
    cy = (shifts[..., 1] + 1.0) * cs / 2.0
    min_x, min_y = cx - 0.5 * px, cy - 0.5 * px
    max_x, max_y = cx + 0.5 * px, cy + 0.5 * px
    out = (_tf_max0(-1 * min_x) + _tf_max0(-1 * min_y) + _tf_max0(max_x - cs) + _tf_max0(max_y - cs)).sum(-1)
    size = _tf_max0(torch.abs(px[:, :, None] - px[:, None, :]) - 3).sum((-1, -2))
    xd = torch.abs(cx[:, :, None] - cx[:, None, :])
    yd = torch.abs(cy[:, :, None] - cy[:, None, :])
    # tf.maximum(x_diff, y_diff): gradient to x_diff when x_diff >= y_diff  [TF-1.12 assumed]
    maxd = torch.where(xd >= yd, xd, yd)
    smean = (px[:, :, None] + px[:, None, :]) / 2.0
    over = _tf_max0(smean - maxd) * (1.0 - torch.eye(T, dtype=dt, device=dev))
    overlap = over.sum((-1, -2))

    pr_loss = pr_num + gamma_area * area + gamma_bbox * (overlap + out) + gamma_size * size
    return dict(per_image=pr_loss + elem, margin=margin, pr_num=pr_num, num_min=elem, area=area, out=out,
                size=size, overlap=overlap, P=P)


def asr_numpy(log_odds, shifts, scales, g_per_image=None, g_margin=1.0, dtype=np.float64, **cfg):
    """numpy in / numpy out, values + gradients.

    Gradients are of ``sum_b g_per_image[b]*per_image[b] + g_margin*margin``; the default
    ``g_per_image = 1/B`` reproduces ``reduce_mean(...) + margin`` (``:1078-1079``).
    """
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    lo = torch.tensor(np.asarray(log_odds), dtype=tdt, requires_grad=True)
    sh = torch.tensor(np.asarray(shifts), dtype=tdt, requires_grad=True)
    sc = torch.tensor(np.asarray(scales), dtype=tdt, requires_grad=True)
    B = lo.shape[0]
    r = asr_terms(lo, sh, sc, **cfg)
    gp = torch.full((B,), 1.0 / B, dtype=tdt) if g_per_image is None else torch.tensor(np.asarray(g_per_image), dtype=tdt)
    total = (r["per_image"] * gp).sum() + g_margin * r["margin"]
    total.backward()
    res = {k: v.detach().numpy() for k, v in r.items()}
    zero = lambda t: np.zeros(t.shape, dtype)
    res["d_log_odds"] = lo.grad.numpy() if lo.grad is not None else zero(lo)
    res["d_shifts"] = sh.grad.numpy() if sh.grad is not None else zero(sh)
    res["d_scales"] = sc.grad.numpy() if sc.grad is not None else zero(sc)
    return res
