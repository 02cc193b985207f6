"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of the reference sampler.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product path (``mog_asr_b200``) never imports anything from ``oracle/``.

**Parity unpinned against a live TensorFlow**: the reference (taufikxu/MOG-ASR) ships no tests, golden
vectors or fixtures for this path, and its arithmetic lives in an un-vendored third-party dependency
(TensorFlow 1.12.0, pinned only by prose in ``README.md:6``) that cannot be installed here (Python 3.12,
no index).  This file therefore *defines* the evaluation order that TF-1.12 is assumed to use; every
assumption is flagged ``[TF-1.12 assumed]``.  The CUDA kernels are checked against this definition.
**The forward op graph IS pinned**: ``tests/golden/make_golden_graph.py`` executes the reference's own
``air/transformer.py`` on a numpy stand-in for the TF ops it uses (``tests/golden/tf_shim.py``, same
per-kernel assumptions) and ``forward`` / ``transformer`` here reproduce its outputs bit for bit
(``tests/test_oracle.py``).  The closed-form backward is checked against ``torch.autograd`` applied to the same reference
file on a torch-based twin of the shim (``tests/golden/make_golden_graph_grad.py``).

Line references are to ``/root/reference/air/transformer.py`` unless another file is named.

Everything is strict IEEE fp32, one rounding per arithmetic op (numpy never contracts to FMA).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
I32 = np.int32


# --------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------
def tf_linspace(start: float, stop: float, num: int) -> np.ndarray:
    """``tf.linspace`` as TF-1.12's LinSpace CPU kernel evaluates it (transformer.py:126-128).

    [TF-1.12 assumed] ``step = (stop-start)/(num-1)`` in fp32, ``out[i] = start + step*i`` in fp32
    (one rounding for the product, one for the sum).  Differs from ``np.linspace`` (up to 129 ulp at
    n=256).  ``num == 1`` yields ``[start]``.
    """
    start = F32(start)
    stop = F32(stop)
    if num == 1:
        return np.array([start], dtype=F32)
    step = F32((stop - start) / F32(num - 1))
    i = np.arange(num, dtype=F32)
    return (start + (step * i).astype(F32)).astype(F32)


def meshgrid(height: int, width: int) -> np.ndarray:
    """``_meshgrid`` (transformer.py:119-136): ``[3, H*W]`` rows (x_t, y_t, 1), row-major n = i*W + j.

    The two K=1 matmuls (``:126-129``) are products with 1.0 and therefore exact.
    """
    lin_w = tf_linspace(-1.0, 1.0, width)
    lin_h = tf_linspace(-1.0, 1.0, height)
    x_t = (np.ones((height, 1), F32) * lin_w[None, :]).astype(F32)   # :126-127
    y_t = (lin_h[:, None] * np.ones((1, width), F32)).astype(F32)    # :128-129
    x_t_flat = x_t.reshape(1, -1)                                     # :131
    y_t_flat = y_t.reshape(1, -1)                                     # :132
    ones = np.ones_like(x_t_flat)                                     # :134
    return np.concatenate([x_t_flat, y_t_flat, ones], axis=0)         # :135


def affine_grid(theta: np.ndarray, out_size) -> tuple[np.ndarray, np.ndarray]:
    """``_transform`` up to the split (transformer.py:144-163): returns ``x_s, y_s`` of shape [B, N].

    [TF-1.12 assumed] the K=3 batched matmul (``:159``) accumulates left to right,
    ``(t0*x_t + t1*y_t) + t2*1``, each product and sum rounded to fp32, no FMA.
    """
    theta = np.asarray(theta).reshape(-1, 2, 3).astype(F32)           # :144-145
    grid = meshgrid(int(out_size[0]), int(out_size[1]))               # :152
    xt, yt, one = grid[0][None, :], grid[1][None, :], grid[2][None, :]

    def row(r):
        p0 = (theta[:, r, 0:1] * xt).astype(F32)
        p1 = (theta[:, r, 1:2] * yt).astype(F32)
        p2 = (theta[:, r, 2:3] * one).astype(F32)
        return ((p0 + p1).astype(F32) + p2).astype(F32)

    return row(0), row(1)                                             # :160-163


def _sat_floor_to_i32(v: np.ndarray, hi: int) -> np.ndarray:
    """``int32(floor(v))`` (transformer.py:79,81) made total.

    The reference leaves |v| >= 2**31 and NaN to the platform's float->int conversion.  Because both
    corners are clipped to ``[0, hi-1]`` right afterwards (``:84-87``), clamping ``floor(v)`` to
    ``[-1, hi]`` first gives the same clipped corners for every finite ``v`` representable in int32 and
    a defined answer for the rest.  NaN maps to -1 (``fmax`` semantics), so corners become (0, 0) and
    the NaN still propagates through the weights.
    """
    f = np.floor(v)
    f = np.fmin(np.fmax(f, F32(-1.0)), F32(hi))
    return f.astype(I32)


def pixel_coords_and_corners(x_s, y_s, height: int, width: int):
    """``_interpolate`` lines 75-87: pixel coordinates and the four clipped corner indices."""
    width_f = F32(width)
    height_f = F32(height)
    wscale = F32(width_f - F32(1.001))
    hscale = F32(height_f - F32(1.001))
    # :75-76   (x + 1.0)*(width_f-1.001) / 2.0   -- multiply, then divide
    x = (((x_s + F32(1.0)).astype(F32) * wscale).astype(F32) / F32(2.0)).astype(F32)
    y = (((y_s + F32(1.0)).astype(F32) * hscale).astype(F32) / F32(2.0)).astype(F32)
    x0 = _sat_floor_to_i32(x, width)                                  # :79
    x1 = x0 + 1                                                       # :80
    y0 = _sat_floor_to_i32(y, height)                                 # :81
    y1 = y0 + 1                                                       # :82
    x0 = np.clip(x0, 0, width - 1).astype(I32)                        # :84
    x1 = np.clip(x1, 0, width - 1).astype(I32)                        # :85
    y0 = np.clip(y0, 0, height - 1).astype(I32)                       # :86
    y1 = np.clip(y1, 0, height - 1).astype(I32)                       # :87
    return x, y, x0, x1, y0, y1


def weights(x, y, x0, x1, y0, y1):
    """transformer.py:108-115 -- weights from the *clipped* corners cast back to fp32."""
    x0f, x1f, y0f, y1f = (a.astype(F32) for a in (x0, x1, y0, y1))
    wa = ((x1f - x).astype(F32) * (y1f - y).astype(F32)).astype(F32)
    wb = ((x1f - x).astype(F32) * (y - y0f).astype(F32)).astype(F32)
    wc = ((x - x0f).astype(F32) * (y1f - y).astype(F32)).astype(F32)
    wd = ((x - x0f).astype(F32) * (y - y0f).astype(F32)).astype(F32)
    return wa, wb, wc, wd


def transformer_full(U, theta, out_size):
    """Forward with every intermediate the parity tests look at.

    Returns a dict: ``out [B,Ho,Wo,C]``; ``x0,x1,y0,y1 [B,N] int32`` (the bit-exact contract);
    ``idx_a..idx_d [B,N] int64`` flat indices (``:88-96``; the reference uses int32 -- int64 here only so
    that the oracle itself never wraps); ``x,y [B,N]`` pixel coords; ``wa..wd [B,N]``.
    """
    U = np.ascontiguousarray(U, dtype=F32)                            # :101
    B, H, W, C = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    N = Ho * Wo
    x_s, y_s = affine_grid(theta, (Ho, Wo))
    assert x_s.shape == (B, N), (x_s.shape, B, N)
    x, y, x0, x1, y0, y1 = pixel_coords_and_corners(x_s, y_s, H, W)
    base = (np.arange(B, dtype=np.int64) * (W * H))[:, None]          # :88-90 (_repeat :48-54)
    base_y0 = base + y0.astype(np.int64) * W                          # :91
    base_y1 = base + y1.astype(np.int64) * W                          # :92
    idx_a = base_y0 + x0                                              # :93
    idx_b = base_y1 + x0                                              # :94
    idx_c = base_y0 + x1                                              # :95
    idx_d = base_y1 + x1                                              # :96
    im_flat = U.reshape(-1, C)                                        # :100
    Ia, Ib, Ic, Id = (im_flat[i] for i in (idx_a, idx_b, idx_c, idx_d))  # :102-105  [B,N,C]
    wa, wb, wc, wd = weights(x, y, x0, x1, y0, y1)
    # :116  tf.add_n sums in list order  [TF-1.12 assumed]
    out = (wa[..., None] * Ia).astype(F32)
    out = (out + (wb[..., None] * Ib).astype(F32)).astype(F32)
    out = (out + (wc[..., None] * Ic).astype(F32)).astype(F32)
    out = (out + (wd[..., None] * Id).astype(F32)).astype(F32)
    return dict(out=out.reshape(B, Ho, Wo, C), x=x, y=y, x0=x0, x1=x1, y0=y0, y1=y1,
                idx_a=idx_a, idx_b=idx_b, idx_c=idx_c, idx_d=idx_d,
                wa=wa, wb=wb, wc=wc, wd=wd, Ia=Ia, Ib=Ib, Ic=Ic, Id=Id)


def transformer(U, theta, out_size, name="SpatialTransformer", **kwargs):
    """Same surface as the reference ``transformer`` (transformer.py:18): returns ``[B,Ho,Wo,C]`` fp32."""
    return transformer_full(U, theta, out_size)["out"]


def batch_transformer(U, thetas, out_size, name="BatchSpatialTransformer"):
    """transformer.py:178-195: repeat each image ``num_transforms`` times, then sample."""
    thetas = np.asarray(thetas)
    B, T = thetas.shape[:2]
    rep = np.repeat(np.arange(B), T)                                  # :192-194
    return transformer(np.asarray(U)[rep], thetas.reshape(B * T, -1), out_size)


def footprint_counts(full: dict, B: int) -> np.ndarray:
    """F[b] = number of distinct source pixels image b addresses (SURVEY 8(d) roofline accounting)."""
    F = np.zeros(B, dtype=np.int64)
    for b in range(B):
        F[b] = np.unique(np.concatenate([full[k][b] for k in ("idx_a", "idx_b", "idx_c", "idx_d")])).size
    return F


# --------------------------------------------------------------------------------------------------
# backward: what TF autodiff of the graph above yields.  dtype selects fp32 or fp64
# accumulation; the fp64 variant is the yardstick the CUDA gradients are compared with.
# --------------------------------------------------------------------------------------------------
def transformer_backward(U, theta, out_size, gout, dtype=np.float64, need_dU=True):
    """Closed-form gradients.

    * gather gradient = scatter-add: ``dU[idx_k] += w_k * g``  (transformer.py:102-105,116)
    * only the explicit ``x``/``y`` terms in the weights (``:112-115``) carry gradient; floor, the int
      casts and the clips (``:79-87``) do not
    * ``dx_s = dx*(W-1.001)/2`` (``:75``), ``dtheta = [dx_s; dy_s] . grid^T`` (``:159``)

    Coordinates, corners and weights are always the fp32 forward values (they are data-dependent
    branch points); only products/sums of the backward itself run in ``dtype``.
    Returns ``(dU [B,H,W,C] or None, dtheta [B,2,3])``.
    """
    full = transformer_full(U, theta, out_size)
    U = np.asarray(U, dtype=F32)
    B, H, W, C = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    N = Ho * Wo
    g = np.asarray(gout, dtype=dtype).reshape(B, N, C)
    x, y = full["x"].astype(dtype), full["y"].astype(dtype)
    x0f, x1f, y0f, y1f = (full[k].astype(dtype) for k in ("x0", "x1", "y0", "y1"))
    wx0, wx1 = x1f - x, x - x0f
    wy0, wy1 = y1f - y, y - y0f
    dU = None
    if need_dU:
        dU = np.zeros((B * H * W, C), dtype=dtype)
        for idx, w in (("idx_a", wx0 * wy0), ("idx_b", wx0 * wy1), ("idx_c", wx1 * wy0), ("idx_d", wx1 * wy1)):
            np.add.at(dU, full[idx].reshape(-1), (w[..., None] * g).reshape(-1, C))
        dU = dU.reshape(B, H, W, C)
    da, db, dc, dd = ((g * full[k].astype(dtype)).sum(-1) for k in ("Ia", "Ib", "Ic", "Id"))
    dx = -wy0 * da - wy1 * db + wy0 * dc + wy1 * dd
    dy = -wx0 * da + wx0 * db - wx1 * dc + wx1 * dd
    wscale = dtype(F32(F32(W) - F32(1.001)))
    hscale = dtype(F32(F32(H) - F32(1.001)))
    dxs = dx * wscale / dtype(2.0)
    dys = dy * hscale / dtype(2.0)
    grid = meshgrid(Ho, Wo).astype(dtype)                              # [3,N]
    dtheta = np.stack([dxs @ grid.T, dys @ grid.T], axis=1)            # [B,2,3]
    return dU, dtheta


def backward_term_magnitudes(U, theta, out_size, gout):
    """Sum of |terms| entering each gradient entry: the scale the fp32 tolerances are stated against.

    Returns ``(absdU [B,H,W,C], absdtheta [B,2,3])`` in fp64.
    """
    full = transformer_full(U, theta, out_size)
    U = np.asarray(U, dtype=F32)
    B, H, W, C = U.shape
    Ho, Wo = int(out_size[0]), int(out_size[1])
    N = Ho * Wo
    g = np.abs(np.asarray(gout, dtype=np.float64).reshape(B, N, C))
    absdU = np.zeros((B * H * W, C))
    for idx, w in (("idx_a", "wa"), ("idx_b", "wb"), ("idx_c", "wc"), ("idx_d", "wd")):
        np.add.at(absdU, full[idx].reshape(-1), (np.abs(full[w].astype(np.float64))[..., None] * g).reshape(-1, C))
    x, y = full["x"].astype(np.float64), full["y"].astype(np.float64)
    x0f, x1f, y0f, y1f = (full[k].astype(np.float64) for k in ("x0", "x1", "y0", "y1"))
    wx0, wx1, wy0, wy1 = np.abs(x1f - x), np.abs(x - x0f), np.abs(y1f - y), np.abs(y - y0f)
    da, db, dc, dd = ((g * np.abs(full[k].astype(np.float64))).sum(-1) for k in ("Ia", "Ib", "Ic", "Id"))
    adx = (wy0 * da + wy1 * db + wy0 * dc + wy1 * dd) * (W - 1.001) / 2
    ady = (wx0 * da + wx0 * db + wx1 * dc + wx1 * dd) * (H - 1.001) / 2
    grid = np.abs(meshgrid(Ho, Wo).astype(np.float64))
    absdtheta = np.stack([adx @ grid.T, ady @ grid.T], axis=1)
    return absdU.reshape(B, H, W, C), absdtheta


# --------------------------------------------------------------------------------------------------
# AIR call sites (air_number_bbox_location.py)
# --------------------------------------------------------------------------------------------------
def theta_read(s, x, y):
    """air_number_bbox_location.py:513-531 -- ``[[s,0,x],[0,s,y]]``."""
    s, x, y = (np.asarray(a, dtype=F32) for a in (s, x, y))
    z = np.zeros_like(s)
    return np.stack([np.stack([s, z, x], 1), np.stack([z, s, y], 1)], 1)


def theta_write(s, x, y):
    """air_number_bbox_location.py:565-584 -- ``[[1/s,0,-x/s],[0,1/s,-y/s]]`` (fp32 divides)."""
    s, x, y = (np.asarray(a, dtype=F32) for a in (s, x, y))
    z = np.zeros_like(s)
    inv = (F32(1.0) / s).astype(F32)
    return np.stack([np.stack([inv, z, (-x / s).astype(F32)], 1),
                     np.stack([z, inv, (-y / s).astype(F32)], 1)], 1)


def write_composite(canvas, U, theta, z_pres, mask):
    """air_number_bbox_location.py:592-600,:722-727.

    ``canvas [B,Hc,Wc]``; ``U [B,Hw,Ww]`` (the VAE reconstruction window); ``mask [B]`` is
    ``stopping_sum < threshold`` evaluated after the stopping_sum update (``:712``).
    ``canvas + where(mask, z*window_recon, 0)`` -- product rounded, then sum rounded.
    """
    canvas = np.asarray(canvas, dtype=F32)
    B, Hc, Wc = canvas.shape
    win = transformer(np.asarray(U, dtype=F32)[..., None], theta, (Hc, Wc))[..., 0]
    contrib = (np.asarray(z_pres, dtype=F32)[:, None, None] * win).astype(F32)
    contrib = np.where(np.asarray(mask, dtype=bool)[:, None, None], contrib, F32(0.0)).astype(F32)
    return (canvas + contrib).astype(F32)


def write_composite_backward(U, theta, z_pres, mask, gcanvas, dtype=np.float64):
    """SURVEY A.2 composite rule: ``g_window = m*z*g_canvas``, ``dz = m*sum(g_canvas*window_recon)``.

    Returns ``(dU [B,Hw,Ww], dtheta [B,2,3], dz [B])``; d canvas_in is ``gcanvas`` itself.
    """
    gcanvas = np.asarray(gcanvas, dtype=dtype)
    B, Hc, Wc = gcanvas.shape
    m = np.asarray(mask, dtype=bool).astype(dtype)
    z = np.asarray(z_pres, dtype=F32).astype(dtype)
    U4 = np.asarray(U, dtype=F32)[..., None]
    win = transformer(U4, theta, (Hc, Wc))[..., 0].astype(dtype)
    gwin = (m * z)[:, None, None] * gcanvas
    dU, dtheta = transformer_backward(U4, theta, (Hc, Wc), gwin[..., None], dtype=dtype)
    dz = m * (gcanvas * win).sum((1, 2))
    return dU[..., 0], dtheta, dz
