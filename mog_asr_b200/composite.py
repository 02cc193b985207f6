"""Fused write call site + canvas compositing (air/air_number_bbox_location.py:592-600,:718-727).

Reference, per inference step::

    window_recon = transformer(vae_recon[B,28,28,1], theta_recon, [cs, cs])[:, :, :, 0]
    running_recon += tf.where(stopping_sum < threshold, z_pres[:,None] * window_recon, 0)

Here: one kernel forward, one backward (dU, dtheta, dz_pres; the canvas gradient passes through).
"""
from __future__ import annotations

import torch

from . import _lib
from .transformer import _need_cuda, _stream


class _WriteComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, canvas, U, theta, z_pres, stop_sum, threshold, inplace):
        L = _lib.load()
        B, Hc, Wc = canvas.shape
        _, Hw, Ww = U.shape
        out = canvas if inplace else torch.empty_like(canvas)
        if inplace:
            ctx.mark_dirty(canvas)
        with torch.cuda.device(canvas.device):
            _lib.check(L.mog_stn_write_composite_forward(
                U.data_ptr(), theta.data_ptr(), z_pres.data_ptr(),
                stop_sum.data_ptr() if stop_sum is not None else None, float(threshold),
                canvas.data_ptr(), out.data_ptr(), B, Hw, Ww, Hc, Wc, _stream(canvas)),
                "mog_stn_write_composite_forward")
        ctx.save_for_backward(U, theta, z_pres, stop_sum)
        ctx.threshold = float(threshold)
        ctx.cdims = (Hc, Wc)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcanvas):
        U, theta, z_pres, stop_sum = ctx.saved_tensors
        Hc, Wc = ctx.cdims
        need_c, need_U, need_th, need_z = ctx.needs_input_grad[:4]
        L = _lib.load()
        B, Hw, Ww = U.shape
        gcanvas = gcanvas.contiguous()
        dU = torch.empty_like(U) if need_U else None
        dth = torch.empty_like(theta) if need_th else None
        dz = torch.empty_like(z_pres) if need_z else None
        if need_U or need_th or need_z:
            with torch.cuda.device(U.device):
                _lib.check(L.mog_stn_write_composite_backward(
                    U.data_ptr(), theta.data_ptr(), z_pres.data_ptr(),
                    stop_sum.data_ptr() if stop_sum is not None else None, ctx.threshold, gcanvas.data_ptr(),
                    dU.data_ptr() if need_U else None, dth.data_ptr() if need_th else None,
                    dz.data_ptr() if need_z else None, B, Hw, Ww, Hc, Wc, _stream(U)),
                    "mog_stn_write_composite_backward")
        return (gcanvas if need_c else None), dU, dth, dz, None, None, None


def write_composite(canvas, window, theta, z_pres, stop_sum=None, threshold=0.9, inplace=False):
    """``canvas + where(stop_sum < threshold, z_pres * transformer(window, theta, canvas.shape), 0)``.

    canvas : ``[B, cs, cs]`` or ``[B, cs*cs]`` (the reference's ``running_recon``), float32 CUDA
    window : ``[B, ws, ws]`` or ``[B, ws*ws]`` VAE reconstruction of the glimpse (square window assumed
             for the flat form, like the reference's ``windows_size``)
    theta  : ``[B, 6]`` / ``[B, 2, 3]`` write transform (``theta_recon``, :565-584)
    z_pres : ``[B]``;  stop_sum : ``[B]`` *after* the ``+= 1 - z_pres`` update (:712), or None = all active
    Returns the new canvas with the shape of ``canvas``.  ``inplace=True`` updates ``canvas`` itself and
    touches only the pixels that change.
    """
    for t, n in ((canvas, "canvas"), (window, "window"), (theta, "theta"), (z_pres, "z_pres")):
        _need_cuda(t, n)
    shape = canvas.shape
    B = shape[0]
    if canvas.dim() == 2:
        cs = int(round(shape[1] ** 0.5))
        if cs * cs != shape[1]:
            raise ValueError("flat canvas must be square")
        c3 = canvas.view(B, cs, cs) if inplace else canvas.reshape(B, cs, cs)
    else:
        c3 = canvas
    if window.dim() == 2:
        ws = int(round(window.shape[1] ** 0.5))
        if ws * ws != window.shape[1]:
            raise ValueError("flat window must be square")
        window = window.reshape(B, ws, ws)
    if not inplace:
        c3 = c3.to(torch.float32).contiguous()
    elif c3.dtype != torch.float32 or not c3.is_contiguous():
        raise ValueError("inplace compositing needs a contiguous float32 canvas")
    window = window.to(torch.float32).contiguous()
    theta = theta.to(torch.float32).reshape(-1, 6).contiguous()
    z_pres = z_pres.to(torch.float32).reshape(-1).contiguous()
    if stop_sum is not None:
        _need_cuda(stop_sum, "stop_sum")
        stop_sum = stop_sum.to(torch.float32).reshape(-1).contiguous()
    if not (theta.shape[0] == B == window.shape[0] == z_pres.shape[0]):
        raise ValueError("batch sizes of canvas, window, theta and z_pres differ")
    out = _WriteComposite.apply(c3, window, theta, z_pres, stop_sum, threshold, inplace)
    return out.view(shape)
