"""ORACLE (test infrastructure, NOT product code) -- the reconstruction loss of AIR restated in numpy.

Follows ``/root/reference/air/air_number_bbox_location.py:945-968``: clip by ``tf.maximum(tf.minimum(r, 1), 0)``,
``log(r + 1e-10)``, row sums; the gradient is what TF autodiff yields ([TF-1.12 assumed] ``minimum``/``maximum``
route the gradient to their first argument on ties, so both clip bounds pass it).  Pinned to the reference's own lines
exec'd on the torch TF shim (``tests/golden/make_golden_blocks.py``, ``tests/test_oracle.py``); not to a live TensorFlow."""
import numpy as np


def reconstruction_loss(canvas, images, dtype=np.float64):
    c = np.asarray(canvas, dtype).reshape(len(canvas), -1)
    x = np.asarray(images, dtype).reshape(len(images), -1)
    eps = dtype(1e-10)
    r = np.maximum(np.minimum(c, dtype(1.0)), dtype(0.0))                  # :947-948
    loss = -(x * np.log(r + eps) + (1 - x) * np.log(1 - r + eps)).sum(1)    # :954-959
    mse = ((x - r) ** 2).sum(1)                                             # :960-961
    return loss, mse


def reconstruction_loss_backward(canvas, images, g_loss, dtype=np.float64):
    c = np.asarray(canvas, dtype).reshape(len(canvas), -1)
    x = np.asarray(images, dtype).reshape(len(images), -1)
    eps = dtype(1e-10)
    r = np.maximum(np.minimum(c, dtype(1.0)), dtype(0.0))
    d = -x / (r + eps) + (1 - x) / (1 - r + eps)
    passes = (c >= 0) & (c <= 1)
    return np.where(passes, np.asarray(g_loss, dtype)[:, None] * d, 0)


def term_magnitudes(canvas, images, dtype=np.float64):
    """sum of |terms| of each row sum: the scale the fp32 tolerance is stated against"""
    c = np.asarray(canvas, dtype).reshape(len(canvas), -1)
    x = np.asarray(images, dtype).reshape(len(images), -1)
    eps = dtype(1e-10)
    r = np.maximum(np.minimum(c, dtype(1.0)), dtype(0.0))
    return (np.abs(x * np.log(r + eps)) + np.abs((1 - x) * np.log(1 - r + eps))).sum(1)
