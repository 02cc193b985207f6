"""On-device synthetic Multi-MNIST / Multi-dSprites feeder: the in-memory replacement for the reference's generator +
TFRecord + ``shuffle_batch`` queue (``multi_mnist.py:90-301``, ``train_air_pr.py:144-147``).

Placement follows ``generate_multi_image`` (``multi_mnist.py:110-221``) and runs as one kernel (``mog_synth_place``);
the pixels are then pasted by the sampler itself: every object is one write-direction ``transformer`` call of its
sprite onto the canvas box (bilinear; the reference rescales with an order-5 spline, ``:144-153``), clipped to [0, 1],
values below 0.05 dropped (``:153-154``) and summed onto the canvas (``:200``).  Canvases are a pure function of
``(seed, canvas index)``: any batch size, any rank split and any replay order give the same images.  Labels come in
the layout ``detection_metrics`` takes (positions, sizes, counts)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, synth
from .transformer import _need_cuda, _stream, transformer

MODES = {"bbox": 0, "disjoint": 1}


def default_sprites(num=256, side=28, seed=0):
    """digit-like stroke sprites (MNIST cannot be downloaded here): ``[num, side, side]`` float32 in [0, 1]"""
    rng = np.random.default_rng(seed)
    return np.stack([synth._blob(rng, side) for _ in range(num)])


class DeviceMultiObjectDataset:
    def __init__(self, sprites, canvas_size=50, counts=(1, 3), size_range=(17, 23), gap=0, margin=0, mode="disjoint",
                 share_size=False, seed=0, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceMultiObjectDataset needs a CUDA device: there is no CPU fallback")
        sp = torch.as_tensor(sprites, dtype=torch.float32)
        if sp.dim() != 3 or sp.shape[1] != sp.shape[2]:
            raise ValueError("sprites must be [num, side, side]")
        self.sprites = sp.to(self.device).contiguous()
        self.cs, self.counts = int(canvas_size), [int(c) for c in counts]
        self.size_min, self.size_max = int(size_range[0]), int(size_range[1])
        if self.size_min < 2 or self.size_max < self.size_min or self.size_max > self.cs:
            raise ValueError("size_range must satisfy 2 <= min <= max <= canvas_size")
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        self.gap, self.margin, self.mode, self.share_size, self.seed = int(gap), int(margin), mode, bool(share_size), int(seed)
        self.G = max(1, max(self.counts))
        self._counts_arr = (ctypes.c_int * len(self.counts))(*self.counts)

    # ---- placement (one launch) -----------------------------------------------------------------------------
    def place(self, first_canvas: int, B: int):
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=self.device)
        num, pos, size, sprite = i32(B), i32(B, self.G, 2), i32(B, self.G, 2), i32(B, self.G)
        L = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(L.mog_synth_place(self.seed, int(first_canvas), B, self.cs, self.G, self._counts_arr, len(self.counts),
                                         self.size_min, self.size_max, self.gap, self.margin, MODES[self.mode], int(self.share_size),
                                         int(self.sprites.shape[0]), num.data_ptr(), pos.data_ptr(), size.data_ptr(),
                                         sprite.data_ptr(), _stream(num)), "mog_synth_place")
        return num, pos, size, sprite

    # ---- paste: the sampler's write direction ---------------------------------------------------------------
    @staticmethod
    def box_theta(pos, size, canvas_size):
        """backward ST matrix that maps the sprite's [-1, 1] square onto canvas pixels [x, x+w-1] x [y, y+h-1];
        evaluated in float64 and rounded once, so host and device agree bit for bit"""
        p, s = pos.to(torch.float64), size.to(torch.float64)
        d = float(canvas_size - 1)
        sc = (s - 1.0) / d                                     # scale per axis
        c = -1.0 + (2.0 * p + s - 1.0) / d                      # centre per axis
        sc = torch.where(s > 0, sc, torch.ones_like(sc))       # unused slots: any finite matrix (masked out later)
        z = torch.zeros_like(sc[..., 0])
        th = torch.stack([1.0 / sc[..., 0], z, -c[..., 0] / sc[..., 0], z, 1.0 / sc[..., 1], -c[..., 1] / sc[..., 1]], -1)
        return th.to(torch.float32)

    def paste(self, num, pos, size, sprite):
        B, G, cs = num.shape[0], self.G, self.cs
        U = self.sprites.index_select(0, sprite.reshape(-1).long()).unsqueeze(-1)                      # [B*G, D, D, 1]
        v = transformer(U, self.box_theta(pos, size, cs).reshape(B * G, 6), (cs, cs)).reshape(B, G, cs * cs)
        v = torch.clamp(v, 0.0, 1.0)                                                                    # :153
        live = (torch.arange(G, device=self.device)[None, :] < num[:, None])[:, :, None]
        return torch.where((v >= 0.05) & live, v, torch.zeros_like(v)).sum(1)                           # :154, :200

    def batch(self, index: int, batch_size: int, rank: int = 0, world: int = 1):
        """canvases ``[index*batch_size*world + rank*batch_size, ... + batch_size)`` of the infinite stream"""
        first = (int(index) * world + rank) * batch_size
        num, pos, size, sprite = self.place(first, batch_size)
        return dict(images=self.paste(num, pos, size, sprite), num=num, pos=pos, size=size, sprite=sprite, first_canvas=first)

    def stream(self, batch_size: int, steps: int, rank: int = 0, world: int = 1, start: int = 0):
        """the training loop's ``shuffle_batch`` replacement: ``steps`` resident device batches, generated on demand"""
        for k in range(start, start + steps):
            yield self.batch(k, batch_size, rank, world)
