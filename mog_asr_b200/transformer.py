"""Host-side mirror of the reference operator surface ``air/transformer.py``.

``transformer(U, theta, out_size, name='SpatialTransformer', **kwargs)``  -- transformer.py:18
``batch_transformer(U, thetas, out_size, name='BatchSpatialTransformer')`` -- transformer.py:178

Same argument order, names and meaning; tensors are torch CUDA tensors instead of TF tensors, and the
call is differentiable w.r.t. ``U`` and ``theta`` through ``torch.autograd`` (the role TF autodiff plays
at air_number_bbox_location.py:1098).  All arithmetic happens in libmogstn's sm_100a kernels; there is
no CPU / eager-PyTorch fallback -- a non-CUDA tensor raises.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the sampler has no CPU fallback "
                           f"(got {type(t).__name__} on {getattr(t, 'device', None)})")


def _prep(U: torch.Tensor, theta: torch.Tensor, out_size, u_div: int):
    _need_cuda(U, "U")
    _need_cuda(theta, "theta")
    if U.dim() != 4:
        raise ValueError(f"U must be [num_batch, height, width, num_channels], got shape {tuple(U.shape)}")
    Ho, Wo = int(out_size[0]), int(out_size[1])
    U = U.to(torch.float32).contiguous()                    # transformer.py:101
    theta = theta.to(torch.float32).reshape(-1, 6).contiguous()  # transformer.py:144-145
    B = theta.shape[0]
    if B != U.shape[0] * u_div:
        raise ValueError(f"theta describes {B} transforms but U holds {U.shape[0]} images x {u_div}")
    if theta.device != U.device:
        raise ValueError("U and theta must live on the same device")
    return U, theta, Ho, Wo, B


class _SpatialTransformer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, theta, Ho, Wo, u_div):
        L = _lib.load()
        B = theta.shape[0]
        _, Hs, Ws, C = U.shape
        out = torch.empty((B, Ho, Wo, C), dtype=torch.float32, device=U.device)
        with torch.cuda.device(U.device):
            _lib.check(L.mog_stn_forward(U.data_ptr(), theta.data_ptr(), out.data_ptr(), B, Hs, Ws, C, Ho, Wo,
                                         u_div, _stream(U)), "mog_stn_forward")
        ctx.save_for_backward(U, theta)
        ctx.dims = (Ho, Wo, u_div)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        U, theta = ctx.saved_tensors
        Ho, Wo, u_div = ctx.dims
        need_dU, need_dth = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_dU or need_dth):
            return None, None, None, None, None
        L = _lib.load()
        B = theta.shape[0]
        _, Hs, Ws, C = U.shape
        gout = gout.to(torch.float32).contiguous()
        dU = torch.empty_like(U) if need_dU else None        # fully overwritten by the kernel
        dth = torch.empty_like(theta) if need_dth else None
        with torch.cuda.device(U.device):
            _lib.check(L.mog_stn_backward(U.data_ptr(), theta.data_ptr(), gout.data_ptr(),
                                          dU.data_ptr() if need_dU else None,
                                          dth.data_ptr() if need_dth else None,
                                          B, Hs, Ws, C, Ho, Wo, u_div, _stream(U)), "mog_stn_backward")
        return dU, dth, None, None, None


def transformer(U, theta, out_size, name="SpatialTransformer", **kwargs):
    """Spatial Transformer Layer -- drop-in for ``air/transformer.py:18``.

    U : float CUDA tensor ``[num_batch, height, width, num_channels]``
    theta : ``[num_batch, 6]`` (or ``[num_batch, 2, 3]``; reshaped like transformer.py:144)
    out_size : ``(out_height, out_width)``
    ``name`` / ``**kwargs`` only label TF graph scopes in the reference and are ignored here.
    Returns ``[num_batch, out_height, out_width, num_channels]`` float32.
    """
    U, theta2, Ho, Wo, _ = _prep(U, theta, out_size, 1)
    return _SpatialTransformer.apply(U, theta2, Ho, Wo, 1)


def batch_transformer(U, thetas, out_size, name="BatchSpatialTransformer"):
    """Batch Spatial Transformer Layer -- drop-in for ``air/transformer.py:178``.

    thetas : ``[num_batch, num_transforms, 6]``.  Returns ``[num_batch*num_transforms, out_height,
    out_width, num_channels]``.  The reference physically repeats every image ``num_transforms`` times
    (transformer.py:192-194); here the kernel indexes ``U[b // num_transforms]`` instead.
    """
    if thetas.dim() < 2:
        raise ValueError("thetas must be [num_batch, num_transforms, 6]")
    T = int(thetas.shape[1])
    U, theta2, Ho, Wo, _ = _prep(U, thetas, out_size, T)
    return _SpatialTransformer.apply(U, theta2, Ho, Wo, T)


def stn_corners(theta, in_size, out_size):
    """Parity probe: the clipped corner indices ``x0, x1, y0, y1`` (transformer.py:79-87) as an int32
    tensor ``[4, B, Ho*Wo]`` computed by the same device code the sampler uses."""
    _need_cuda(theta, "theta")
    theta = theta.to(torch.float32).reshape(-1, 6).contiguous()
    B = theta.shape[0]
    Hs, Ws = int(in_size[0]), int(in_size[1])
    Ho, Wo = int(out_size[0]), int(out_size[1])
    out = torch.empty((4, B, Ho * Wo), dtype=torch.int32, device=theta.device)
    with torch.cuda.device(theta.device):
        _lib.check(_lib.load().mog_stn_corners(theta.data_ptr(), out.data_ptr(), B, Hs, Ws, Ho, Wo, _stream(theta)),
                   "mog_stn_corners")
    return out
