"""Fused reconstruction loss of AIR (air/air_number_bbox_location.py:945-968): clip to [0,1], cross-entropy with
``epsilon = 1e-10`` summed over pixels, plus the logged squared error -- one kernel forward, one backward."""
from __future__ import annotations

import torch

from . import _lib
from .transformer import _need_cuda, _stream


class _BceRecon(torch.autograd.Function):
    @staticmethod
    def forward(ctx, canvas, images):
        L = _lib.load()
        B, P = canvas.shape
        loss = torch.empty(B, dtype=torch.float32, device=canvas.device)
        mse = torch.empty(B, dtype=torch.float32, device=canvas.device)
        with torch.cuda.device(canvas.device):
            _lib.check(L.mog_bce_recon_forward(canvas.data_ptr(), images.data_ptr(), loss.data_ptr(), mse.data_ptr(), B, P,
                                               _stream(canvas)), "mog_bce_recon_forward")
        ctx.save_for_backward(canvas, images)
        ctx.mark_non_differentiable(mse)
        return loss, mse

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, _g_mse):
        canvas, images = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        L = _lib.load()
        B, P = canvas.shape
        g_loss = g_loss.to(torch.float32).contiguous()
        dcanvas = torch.empty_like(canvas)
        with torch.cuda.device(canvas.device):
            _lib.check(L.mog_bce_recon_backward(canvas.data_ptr(), images.data_ptr(), g_loss.data_ptr(), dcanvas.data_ptr(),
                                                B, P, _stream(canvas)), "mog_bce_recon_backward")
        return dcanvas, None


def reconstruction_loss(canvas, images):
    """``canvas`` (the un-clipped ``running_recon``), ``images``: ``[B, cs*cs]`` (or ``[B, cs, cs]``) float32 CUDA.
    Returns ``(reconstruction_loss [B], mse_loss [B])`` like ``:959-966``; differentiable w.r.t. ``canvas``."""
    _need_cuda(canvas, "canvas")
    _need_cuda(images, "images")
    B = canvas.shape[0]
    c2 = canvas.reshape(B, -1).to(torch.float32).contiguous()
    x2 = images.reshape(B, -1).to(torch.float32).contiguous()
    if c2.shape != x2.shape:
        raise ValueError("canvas and images must have the same number of pixels")
    return _BceRecon.apply(c2, x2)
