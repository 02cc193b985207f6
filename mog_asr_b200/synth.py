"""Synthetic inputs of the named shapes: Multi-MNIST-like canvases, AIR-like thetas.

MNIST cannot be downloaded here and the reference bundles no Multi-MNIST data, so objects are seeded
random blobs of the same geometry (28x28 crops with values in [0,1], 1-3 per 50x50 canvas, placed with
the rejection rule of multi_mnist.py:171-197).  theta follows the AIR call sites:
read ``[[s,0,x],[0,s,y]]`` (air_number_bbox_location.py:513-531), write ``[[1/s,0,-x/s],[0,1/s,-y/s]]``
(:565-584), with ``s = sigmoid(N(-1, 0.05))`` (the scale prior, :75-76) and ``(x,y) = tanh(N(0,1))``.
"""
from __future__ import annotations

import numpy as np


def _blob(rng: np.random.Generator, size: int) -> np.ndarray:
    """A digit-like stroke pattern: a few thick random strokes, blurred, in [0,1]."""
    img = np.zeros((size, size), np.float32)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    pts = rng.uniform(0.2 * size, 0.8 * size, size=(4, 2))
    for a, b in zip(pts[:-1], pts[1:]):
        for t in np.linspace(0, 1, 12):
            c = a * (1 - t) + b * t
            img += np.exp(-((xx - c[0]) ** 2 + (yy - c[1]) ** 2) / (2 * 1.6 ** 2))
    return np.clip(img / max(img.max(), 1e-6) * 1.2, 0, 1).astype(np.float32)


def multi_object_canvases(batch: int, canvas: int = 50, obj: int = 28, counts=(1, 2, 3), seed: int = 0):
    """``[batch, canvas, canvas]`` float32 canvases and the object count per canvas.

    Placement: uniform position, rejected (<=100 tries) while it overlaps an already placed object's
    pixels (multi_mnist.py:171-197, pixel-overlap mode); overlapping sums may exceed 1 like the
    reference's ``canvas += image`` (multi_mnist.py:200)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, canvas, canvas), np.float32)
    num = rng.choice(np.asarray(counts), size=batch)
    for b in range(batch):
        occ = np.zeros((canvas, canvas), bool)
        for _ in range(int(num[b])):
            side = int(rng.integers(max(8, obj // 2), min(obj, canvas) + 1))
            img = _blob(rng, side)
            for _try in range(100):
                y, x = rng.integers(0, canvas - side + 1, size=2)
                if not (occ[y:y + side, x:x + side] & (img > 0.05)).any():
                    break
            out[b, y:y + side, x:x + side] += img
            occ[y:y + side, x:x + side] |= img > 0.05
    return out, num.astype(np.int32)


def sxy_prior_like(batch: int, seed: int = 1):
    """s = sigmoid(N(-1, var 0.05)), (x, y) = tanh(N(0, 1))."""
    rng = np.random.default_rng(seed)
    s = 1.0 / (1.0 + np.exp(-rng.normal(-1.0, np.sqrt(0.05), batch)))
    x, y = np.tanh(rng.normal(0, 1, batch)), np.tanh(rng.normal(0, 1, batch))
    return s.astype(np.float32), x.astype(np.float32), y.astype(np.float32)


def sxy_full_cover(batch: int, seed: int = 1):
    """s ~ U[0.8, 1], shifts ~ U[-(1-s), 1-s]: the glimpse covers (almost) the whole source."""
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.8, 1.0, batch)
    x, y = rng.uniform(-1, 1, batch) * (1 - s), rng.uniform(-1, 1, batch) * (1 - s)
    return s.astype(np.float32), x.astype(np.float32), y.astype(np.float32)


def theta_read(s, x, y):
    """``[[s,0,x],[0,s,y]]`` as ``[B,6]`` float32 (numpy or torch inputs)."""
    if isinstance(s, np.ndarray):
        z = np.zeros_like(s)
        return np.stack([s, z, x, z, s, y], 1).astype(np.float32)
    import torch
    z = torch.zeros_like(s)
    return torch.stack([s, z, x, z, s, y], 1)


def theta_write(s, x, y):
    """``[[1/s,0,-x/s],[0,1/s,-y/s]]`` as ``[B,6]`` float32 (fp32 divides, like the reference graph)."""
    if isinstance(s, np.ndarray):
        s, x, y = s.astype(np.float32), x.astype(np.float32), y.astype(np.float32)
        z = np.zeros_like(s)
        inv = (np.float32(1.0) / s).astype(np.float32)
        return np.stack([inv, z, (-x / s).astype(np.float32), z, inv, (-y / s).astype(np.float32)], 1)
    import torch
    z = torch.zeros_like(s)
    inv = 1.0 / s
    return torch.stack([inv, z, -x / s, z, inv, -y / s], 1)
