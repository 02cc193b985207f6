"""ORACLE (test infrastructure, NOT product code) -- the synthetic-dataset placement and paste, restated on the host.

Placement rules: ``/root/reference/multi_mnist.py:110-221`` (count, size / shared size ``:119,:136-142``, <= 100 position
draws inside the margins ``:169-172``, ``bounding_boxes_overlap`` ``:77-87`` as written for mode 0, restart of the
canvas when an object does not fit ``:112-113,:209-210``).  **Parity pinned for the overlap rule only** (``boxes_clash`` mode 0
against the outputs of the reference's own ``bounding_boxes_overlap``, ``tests/golden/synth_overlap_rule.npz``); **unpinned
for the rest**: the reference draws from numpy's global Mersenne Twister after ``np.random.seed(0)`` over a downloaded MNIST and rescales with
scipy's order-5 spline; none of that can be replayed on the device, so this file pins the product's counter-based
draws and its bilinear paste instead (same hash, same arithmetic, scalar Python)."""
import numpy as np

from . import stn_ref_c

M32 = 0xFFFFFFFF


def mix32(x):
    x &= M32
    x ^= x >> 16; x = (x * 0x7feb352d) & M32
    x ^= x >> 15; x = (x * 0x846ca68b) & M32
    x ^= x >> 16
    return x


def draw32(seed, canvas, restart, obj, attempt, field):
    h = mix32((seed & M32) ^ 0x9e3779b9)
    h = mix32(h ^ ((seed >> 32) & M32))
    h = mix32(h ^ (canvas & M32))
    h = mix32(h ^ ((restart * 0x85ebca6b + obj) & M32))
    h = mix32(h ^ ((attempt * 0xc2b2ae35 + field) & M32))
    return h


def draw_int(h, lo, hi):
    return lo + ((h * (hi - lo)) >> 32)


def boxes_clash(mode, x, y, w, h, qx, qy, qw, qh, gap):
    """one pair of the overlap test.  mode 0: ``bounding_boxes_overlap`` (multi_mnist.py:77-87) as written -- pinned by
    ``tests/golden/synth_overlap_rule.npz``, which holds the outputs of the reference function itself; mode 1: true box
    intersection."""
    l1x, l1y, r1x, r1y = x - gap, y - gap, x + w + gap - 1, y + h + gap - 1            # :79
    l2x, l2y, r2x, r2y = qx, qy, qx + qw - 1, qy + qh - 1                              # :80
    xhit = l1x <= r2x and l2x <= r1x
    if mode == 0:
        return bool(xhit or (l1y >= r2y and l2y >= r1y))                               # :82-85
    return bool(xhit and l1y <= r2y and l2y <= r1y)


def place(seed, first_canvas, B, canvas, counts, size_min, size_max, gap=0, margin=0, mode=0, share_size=False, num_sprites=1,
          max_restarts=32):
    G = max(1, max(counts))
    num = np.zeros(B, np.int32)
    pos, size, sprite = np.zeros((B, G, 2), np.int32), np.zeros((B, G, 2), np.int32), np.zeros((B, G), np.int32)
    for b in range(B):
        cid = first_canvas + b
        n = counts[draw_int(draw32(seed, cid, 0, 0, 0, 7), 0, len(counts))]
        placed = []
        for restart in range(max_restarts):
            placed = []
            shared = draw_int(draw32(seed, cid, restart, 0, 0, 1), size_min, size_max + 1)           # :119
            ok = True
            for i in range(n):
                w = shared if share_size else draw_int(draw32(seed, cid, restart, i, 0, 2), size_min, size_max + 1)
                sp = draw_int(draw32(seed, cid, restart, i, 0, 3), 0, num_sprites)
                span = canvas - w - 2 * margin + 1
                found = False
                x = y = 0
                if span > 0:
                    for att in range(100):                                                            # :169
                        x = margin + draw_int(draw32(seed, cid, restart, i, att, 4), 0, span)        # :171
                        y = margin + draw_int(draw32(seed, cid, restart, i, att, 5), 0, span)        # :172
                        found = not any(boxes_clash(mode, x, y, w, w, qx, qy, qw, qw, gap) for (qx, qy, qw, _s) in placed)
                        if found:
                            break
                if not found:
                    ok = False                                                                        # :209-210
                    break
                placed.append((x, y, w, sp))
            if ok:
                break
        num[b] = len(placed)
        for g, (x, y, w, sp) in enumerate(placed):
            pos[b, g], size[b, g], sprite[b, g] = (x, y), (w, w), sp
    return num, pos, size, sprite


def box_theta(pos, size, canvas):
    p, s = pos.astype(np.float64), size.astype(np.float64)
    d = float(canvas - 1)
    sc = np.where(s > 0, (s - 1.0) / d, 1.0)
    c = -1.0 + (2.0 * p + s - 1.0) / d
    z = np.zeros_like(sc[..., 0])
    return np.stack([1.0 / sc[..., 0], z, -c[..., 0] / sc[..., 0], z, 1.0 / sc[..., 1], -c[..., 1] / sc[..., 1]], -1).astype(np.float32)


def paste(sprites, canvas, num, pos, size, sprite):
    B, G = sprite.shape
    U = np.asarray(sprites, np.float32)[sprite.reshape(-1)][..., None]
    v = stn_ref_c.forward(U, box_theta(pos, size, canvas).reshape(B * G, 6), (canvas, canvas)).reshape(B, G, canvas * canvas)
    v = np.clip(v, 0.0, 1.0)
    live = (np.arange(G)[None, :] < num[:, None])[:, :, None]
    out = np.zeros((B, canvas * canvas), np.float32)
    for g in range(G):                                      # same left-to-right sum as the device reduction over G <= 8
        out = out + np.where((v[:, g] >= 0.05) & live[:, g], v[:, g], np.float32(0))
    return out
