"""The two sampler calls of an AIR step with theta built in the kernel from the model's ``(s, x, y)``.

Reference, per inference step (``air/air_number_bbox_location.py``)::

    theta        = [[s, 0, x], [0, s, y]]                    # :511-531
    window       = transformer(images, theta, [ws, ws])      # :534-542
    theta_recon  = [[1/s, 0, -x/s], [0, 1/s, -y/s]]          # :563-584
    window_recon = transformer(vae_recon, theta_recon, [cs, cs]);  canvas += where(stop < thr, z_pres * window_recon, 0)   # :592-600, :718-727

``read_glimpse_sxy`` and ``write_composite_sxy`` pass ``shift = (x, y)`` and ``scale = s`` straight to the sampler kernels (no
theta tensors, no theta kernels) and get ``d_shift``, ``d_scale`` back.  ``read_glimpse_sxy`` also returns ``shift`` and
``scale`` again: hand THOSE to ``write_composite_sxy`` and the gradient the write call sends to ``(s, x, y)`` is added inside the
read call's backward kernel instead of by two accumulation kernels.  Results equal the theta-taking path (same fp32
expressions for theta; the chain rule of theta is evaluated per image in the sampler's epilogue).
"""
from __future__ import annotations

import torch

from . import _lib
from .transformer import _need_cuda, _stream


def _p(t):
    return t.data_ptr() if t is not None else None


class _ReadSxy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, shift, scale, Ho, Wo):
        ctx.set_materialize_grads(False)
        L = _lib.load()
        B, Hs, Ws, _ = U.shape
        out = torch.empty((B, Ho, Wo, 1), dtype=torch.float32, device=U.device)
        with torch.cuda.device(U.device):
            _lib.check(L.mog_stn_read_sxy_forward(_p(U), _p(shift), _p(scale), _p(out), B, Hs, Ws, Ho, Wo, _stream(U)),
                       "mog_stn_read_sxy_forward")
        ctx.save_for_backward(U, shift, scale)
        ctx.dims = (Ho, Wo)
        return out, shift, scale          # (inputs returned as outputs: autograd hands out views with this node as grad_fn)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout, g_shift, g_scale):
        U, shift, scale = ctx.saved_tensors
        Ho, Wo = ctx.dims
        need_dU = ctx.needs_input_grad[0]
        if gout is None:                  # the glimpse itself was not used: only the pass-through gradients
            return None, g_shift, g_scale, None, None
        L = _lib.load()
        B, Hs, Ws, _ = U.shape
        gout = gout.to(torch.float32).contiguous()
        g_shift = g_shift.contiguous() if g_shift is not None else None
        g_scale = g_scale.contiguous() if g_scale is not None else None
        dU = torch.empty_like(U) if need_dU else None
        d_shift, d_scale = torch.empty_like(shift), torch.empty_like(scale)
        with torch.cuda.device(U.device):
            _lib.check(L.mog_stn_read_sxy_backward(_p(U), _p(shift), _p(scale), _p(gout), _p(g_shift), _p(g_scale), _p(dU),
                                                   _p(d_shift), _p(d_scale), B, Hs, Ws, Ho, Wo, _stream(U)),
                       "mog_stn_read_sxy_backward")
        return dU, d_shift, d_scale, None, None


class _WriteCompositeSxy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, canvas, U, shift, scale, z_pres, stop_sum, threshold):
        L = _lib.load()
        B, Hc, Wc = canvas.shape
        _, Hw, Ww = U.shape
        out = torch.empty_like(canvas)
        with torch.cuda.device(canvas.device):
            _lib.check(L.mog_stn_write_composite_sxy_forward(_p(U), _p(shift), _p(scale), _p(z_pres), _p(stop_sum), float(threshold),
                                                             _p(canvas), _p(out), B, Hw, Ww, Hc, Wc, _stream(canvas)),
                       "mog_stn_write_composite_sxy_forward")
        ctx.save_for_backward(U, shift, scale, z_pres, stop_sum)
        ctx.threshold = float(threshold)
        ctx.cdims = (Hc, Wc)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcanvas):
        U, shift, scale, z_pres, stop_sum = ctx.saved_tensors
        Hc, Wc = ctx.cdims
        need_c, need_U, need_sh, need_sc, need_z = ctx.needs_input_grad[:5]
        L = _lib.load()
        B, Hw, Ww = U.shape
        gcanvas = gcanvas.contiguous()
        dU = torch.empty_like(U) if need_U else None
        dz = torch.empty_like(z_pres) if need_z else None
        d_shift = d_scale = None
        if need_U or need_sh or need_sc or need_z:
            d_shift, d_scale = torch.empty_like(shift), torch.empty_like(scale)
            with torch.cuda.device(U.device):
                _lib.check(L.mog_stn_write_composite_sxy_backward(
                    _p(U), _p(shift), _p(scale), _p(z_pres), _p(stop_sum), ctx.threshold, _p(gcanvas), None, None, _p(dU),
                    _p(d_shift), _p(d_scale), _p(dz), B, Hw, Ww, Hc, Wc, _stream(U)), "mog_stn_write_composite_sxy_backward")
        return ((gcanvas if need_c else None), dU, (d_shift if need_sh else None), (d_scale if need_sc else None), dz, None, None)


def _prep_sxy(shift, scale, B):
    for t, n in ((shift, "shift"), (scale, "scale")):
        _need_cuda(t, n)
    if shift.shape != (B, 2) or scale.numel() != B:
        raise ValueError(f"shift must be [{B}, 2] and scale [{B}] or [{B}, 1], got {tuple(shift.shape)} and {tuple(scale.shape)}")
    return shift.to(torch.float32).contiguous(), scale.to(torch.float32).contiguous()


def read_glimpse_sxy(images, shift, scale, out_size):
    """``transformer(images, [[s,0,x],[0,s,y]], out_size)`` (reference ``:511-542``) without a theta tensor.

    images : ``[B, H, W, 1]`` float32 CUDA;  shift : ``[B, 2]`` = (x, y);  scale : ``[B]`` or ``[B, 1]`` = s.
    Returns ``(window [B, Ho, Wo, 1], shift, scale)``; pass the returned shift / scale on to ``write_composite_sxy``."""
    _need_cuda(images, "images")
    if images.dim() != 4 or images.shape[3] != 1:
        raise ValueError(f"images must be [B, H, W, 1], got {tuple(images.shape)}")
    shift, scale = _prep_sxy(shift, scale, images.shape[0])
    return _ReadSxy.apply(images.to(torch.float32).contiguous(), shift, scale, int(out_size[0]), int(out_size[1]))


def write_composite_sxy(canvas, window, shift, scale, z_pres, stop_sum=None, threshold=0.9):
    """``canvas + where(stop_sum < threshold, z_pres * transformer(window, [[1/s,0,-x/s],[0,1/s,-y/s]], canvas.shape), 0)``
    (reference ``:563-600``, ``:718-727``) without a theta tensor.  canvas ``[B, cs, cs]``, window ``[B, ws, ws]``."""
    for t, n in ((canvas, "canvas"), (window, "window"), (z_pres, "z_pres")):
        _need_cuda(t, n)
    if canvas.dim() != 3 or window.dim() != 3:
        raise ValueError("canvas must be [B, cs, cs] and window [B, ws, ws]")
    shift, scale = _prep_sxy(shift, scale, canvas.shape[0])
    stop = stop_sum.detach().to(torch.float32).contiguous() if stop_sum is not None else None
    return _WriteCompositeSxy.apply(canvas.to(torch.float32).contiguous(), window.to(torch.float32).contiguous(), shift, scale,
                                    z_pres.to(torch.float32).reshape(-1).contiguous(), stop, float(threshold))
