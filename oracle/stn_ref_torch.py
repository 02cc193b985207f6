"""ORACLE (test infrastructure, NOT product code) -- the reference sampler's op graph rebuilt from
torch-CPU primitives so that ``torch.autograd`` plays the role TF autodiff plays in the reference
(``optimizer.compute_gradients``, air_number_bbox_location.py:1098).

**Parity unpinned** (see ``stn_ref_numpy.py``).  Used only to cross-check the closed-form backward in
``stn_ref_numpy.transformer_backward`` -- the product never imports this.

``torch.nn.functional.grid_sample`` is deliberately NOT used: the ``(W-1.001)/2`` scaling
(transformer.py:75-76) and the clipped-corner weights (``:84-87,:108-115``) match no ``align_corners``
/ ``padding_mode`` combination.
"""
from __future__ import annotations

import torch


def _linspace(num: int, dtype, device="cpu") -> torch.Tensor:
    # transformer.py:126-128, TF-1.12 LinSpace recurrence evaluated in fp32 then widened
    if num == 1:
        return torch.tensor([-1.0], dtype=dtype, device=device)
    step = torch.tensor(2.0, dtype=torch.float32, device=device) / torch.tensor(float(num - 1), dtype=torch.float32, device=device)
    i = torch.arange(num, dtype=torch.float32, device=device)
    return (torch.tensor(-1.0, dtype=torch.float32, device=device) + step * i).to(dtype)


def transformer(U: torch.Tensor, theta: torch.Tensor, out_size, name="SpatialTransformer", **kwargs):
    """Differentiable restatement of transformer.py:18-175 from torch primitives (any float dtype, any device)."""
    dtype = U.dtype
    B, H, W, C = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    theta = theta.reshape(-1, 2, 3).to(dtype)                                   # :144-145
    dev = U.device
    lin_w, lin_h = _linspace(Wo, dtype, dev), _linspace(Ho, dtype, dev)
    x_t = lin_w[None, :].expand(Ho, Wo).reshape(-1)                             # :126-127,131
    y_t = lin_h[:, None].expand(Ho, Wo).reshape(-1)                             # :128-129,132
    x_s = (theta[:, 0, 0:1] * x_t + theta[:, 0, 1:2] * y_t) + theta[:, 0, 2:3]  # :159
    y_s = (theta[:, 1, 0:1] * x_t + theta[:, 1, 1:2] * y_t) + theta[:, 1, 2:3]
    f32 = torch.float32
    wscale = (torch.tensor(float(W), dtype=f32, device=dev) - torch.tensor(1.001, dtype=f32, device=dev)).to(dtype)
    hscale = (torch.tensor(float(H), dtype=f32, device=dev) - torch.tensor(1.001, dtype=f32, device=dev)).to(dtype)
    x = (x_s + 1.0) * wscale / 2.0                                              # :75
    y = (y_s + 1.0) * hscale / 2.0                                              # :76
    x0 = torch.floor(x.detach()).clamp(-1, W).long()                            # :79
    x1 = x0 + 1                                                                 # :80
    y0 = torch.floor(y.detach()).clamp(-1, H).long()                            # :81
    y1 = y0 + 1                                                                 # :82
    x0, x1 = x0.clamp(0, W - 1), x1.clamp(0, W - 1)                             # :84-85
    y0, y1 = y0.clamp(0, H - 1), y1.clamp(0, H - 1)                             # :86-87
    base = (torch.arange(B, device=dev) * (H * W))[:, None]                                 # :88-90
    idx_a, idx_b = base + y0 * W + x0, base + y1 * W + x0                       # :91-94
    idx_c, idx_d = base + y0 * W + x1, base + y1 * W + x1                       # :95-96
    im_flat = U.reshape(-1, C)                                                  # :100
    Ia, Ib, Ic, Id = (im_flat[i.reshape(-1)] for i in (idx_a, idx_b, idx_c, idx_d))  # :102-105
    x0f, x1f, y0f, y1f = (a.to(dtype) for a in (x0, x1, y0, y1))                # :108-111
    wa = ((x1f - x) * (y1f - y)).reshape(-1, 1)                                 # :112
    wb = ((x1f - x) * (y - y0f)).reshape(-1, 1)                                 # :113
    wc = ((x - x0f) * (y1f - y)).reshape(-1, 1)                                 # :114
    wd = ((x - x0f) * (y - y0f)).reshape(-1, 1)                                 # :115
    out = ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id                             # :116
    return out.reshape(B, Ho, Wo, C)                                            # :169-170


def gradients(U, theta, out_size, gout, dtype=torch.float64):
    """(dU, dtheta) by autograd of the restated graph, numpy in / numpy out."""
    Ut = torch.tensor(U, dtype=dtype, requires_grad=True)
    tt = torch.tensor(theta, dtype=dtype).reshape(-1, 2, 3).requires_grad_(True)
    out = transformer(Ut, tt, out_size)
    out.backward(torch.tensor(gout, dtype=dtype).reshape(out.shape))
    return Ut.grad.numpy(), tt.grad.numpy()
