"""Attention-window masks of the visualisation pass (``air/air_number_bbox_location.py:224-382``): the second consumer
of ``transformer`` in the reference.  A one-pixel frame drawn on a ``windows_size`` template (``:253-271``) is written
through the backward spatial-transformer matrices onto a ``zoom * canvas_size`` canvas, clipped to [0, 1] (``:250-276``)
and sharpened with ``> 0.01`` (``:287-288``).  The reference materialises ``num_images * max_steps`` copies of the
template; here ONE template is shared by all transforms (the kernels index ``U[b // u_batch_div]``)."""
from __future__ import annotations

import torch

from .transformer import _need_cuda, batch_transformer


def frame_template(windows_size: int, device, dtype=torch.float32):
    """``tf.image.draw_bounding_boxes(zeros[ws, ws, 1], box [0, 0, 1, 1])`` (``:253-271``): rows 0 and ws-1 and columns
    0 and ws-1 set to 1  [TF-1.12 assumed: first colour of the default table has 1.0 in channel 0]."""
    t = torch.zeros((1, windows_size, windows_size, 1), device=device, dtype=dtype)
    t[:, 0, :, :] = 1.0
    t[:, -1, :, :] = 1.0
    t[:, :, 0, :] = 1.0
    t[:, :, -1, :] = 1.0
    return t


def attention_boxes(st_back, canvas_size: int, zoom: int = 2, windows_size: int = 28, max_steps=None, threshold: float = 0.01):
    """``st_back``: ``[num_images, steps, 2, 3]`` (or ``[num_images, steps, 6]``) backward ST matrices, CUDA.  Steps are
    zero-padded up to ``max_steps`` like ``:243-246`` (a zero matrix samples the template's centre everywhere: an empty mask).
    Returns the 0/1 masks ``[num_images, max_steps, zoom*canvas_size, zoom*canvas_size]`` of ``:250-288``."""
    _need_cuda(st_back, "st_back")
    n, steps = int(st_back.shape[0]), int(st_back.shape[1])
    th = st_back.reshape(n, steps, 6).to(torch.float32)
    T = steps if max_steps is None else int(max_steps)
    if T < steps:
        raise ValueError("max_steps is smaller than the number of matrices per image")
    if T > steps:
        th = torch.cat([th, th.new_zeros(n, T - steps, 6)], 1)                                   # :243-246
    side = zoom * canvas_size
    tmpl = frame_template(windows_size, th.device)
    out = batch_transformer(tmpl, th.reshape(1, n * T, 6), (side, side))                        # :250-274, one shared template
    boxes = torch.clamp(out, 0.0, 1.0).reshape(n, T, side, side)                                # :275-285
    return (boxes > threshold).to(torch.float32)                                                # :287-288
