"""Detection metrics of the evaluation pass (``air/evaluation_detection.py:29-98``) as one CUDA launch per batch.

``evaluation`` keeps the reference's signature and return value (lists of per-image ground-truth boxes in, five
batch means out); ``detection_metrics`` is the tensor form that stays on the device and returns the per-image values.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .transformer import _need_cuda, _stream

MAX_BOXES = 8  # MOG_DET_MAX_BOXES


def detection_metrics(gt_pos, gt_size, gt_num, inf_shifts, inf_scales, inf_num, csize=50):
    """``gt_pos``, ``gt_size``: ``[B, G, 2]`` integer (x, y) / (w, h); ``gt_num``: ``[B]``; ``inf_shifts``: ``[B, T, 2]``;
    ``inf_scales``: ``[B, T]`` or ``[B, T, 1]``; ``inf_num``: ``[B]``; all CUDA.  Returns per-image float64 tensors
    ``(precision [B,11], recall [B,11], gt_max_iou [B], detected_max_iou [B], global_iou [B])``."""
    _need_cuda(inf_shifts, "inf_shifts")
    dev = inf_shifts.device
    B, T = int(inf_shifts.shape[0]), int(inf_shifts.shape[1])
    G = int(gt_pos.shape[1])
    if G > MAX_BOXES or T > MAX_BOXES:
        raise ValueError(f"at most {MAX_BOXES} ground-truth and {MAX_BOXES} inferred boxes per image (got {G}, {T})")
    for t, n in ((gt_pos, "gt_pos"), (gt_size, "gt_size")):
        # the kernel takes integer boxes (the reference's ground truth is integer pixels, multi_mnist.py:200-215); silently
        # truncating a fractional box would change the IoUs, so it is refused
        if t.is_floating_point() and bool((t != torch.floor(t)).any()):
            raise ValueError(f"{n} must hold integer pixel coordinates (got non-integral values)")
    i32 = lambda t: t.to(device=dev, dtype=torch.int32).contiguous()
    gt_pos, gt_size, gt_num, inf_num = i32(gt_pos).reshape(B, G, 2), i32(gt_size).reshape(B, G, 2), i32(gt_num).reshape(B), i32(inf_num).reshape(B)
    shifts = inf_shifts.to(torch.float64).contiguous().reshape(B, T, 2)
    scales = inf_scales.to(device=dev, dtype=torch.float64).contiguous().reshape(B, T)
    f64 = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
    prec, rec, gmax, dmax, glob = f64(B, 11), f64(B, 11), f64(B), f64(B), f64(B)
    L = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(L.mog_detection_eval(gt_pos.data_ptr(), gt_size.data_ptr(), gt_num.data_ptr(), shifts.data_ptr(), scales.data_ptr(),
                                        inf_num.data_ptr(), B, G, T, float(csize), prec.data_ptr(), rec.data_ptr(), gmax.data_ptr(),
                                        dmax.data_ptr(), glob.data_ptr(), _stream(shifts)), "mog_detection_eval")
    return prec, rec, gmax, dmax, glob


def pack_ground_truth(gt_position_xy, gt_scale_xy):
    """The reference's ragged lists (``[x0, y0, x1, y1, ...]`` per image, ``:41-48``) -> padded ``[B, G, 2]`` int arrays + counts."""
    n = len(gt_position_xy)
    num = np.array([len(p) // 2 for p in gt_position_xy], np.int32)
    G = max(1, int(num.max()) if n else 1)
    pos, size = np.zeros((n, G, 2), np.int32), np.zeros((n, G, 2), np.int32)
    for k in range(n):
        if num[k]:
            pos[k, :num[k]] = np.asarray(gt_position_xy[k][:2 * num[k]]).reshape(-1, 2)
            size[k, :num[k]] = np.asarray(gt_scale_xy[k][:2 * num[k]]).reshape(-1, 2)
    return pos, size, num


def evaluation(gt_position_xy, gt_scale_xy, inf_shifts, inf_scales, inf_num, csize=50, device="cuda"):
    """Drop-in for ``evaluation_detection.evaluation`` (``:29``): returns
    ``(precision[11], recall[11], gt_max_iou, detected_max_iou, global_iou_mean)`` as numpy float64."""
    pos, size, num = pack_ground_truth(gt_position_xy, gt_scale_xy)
    dev = torch.device(device)
    t = lambda a, dt=None: torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
    out = detection_metrics(t(pos), t(size), t(num), t(inf_shifts, torch.float64), t(inf_scales, torch.float64),
                            t(np.asarray(inf_num).astype(np.int32)), csize)
    p, r, g, d, m = (o.cpu().numpy() for o in out)
    return np.mean(p, 0), np.mean(r, 0), np.mean(g), np.mean(d), np.mean(m)   # :97-98
