"""The same stand-in as ``tf_shim.py`` built on torch-CPU tensors, so that ``torch.autograd`` differentiates THE
REFERENCE'S OWN GRAPH (``air/transformer.py`` imported on top of this module): gather's gradient is a scatter-add, floor
and integer casts pass no gradient -- what TF autodiff does for the same ops (``optimizer.compute_gradients``,
air_number_bbox_location.py:1098).  Per-kernel forward numerics as in ``tf_shim.py`` ([TF-1.12 assumed]).
Test infrastructure only."""
import contextlib

import numpy as np
import torch

_DT = {"float32": torch.float32, "int32": torch.int32, "float64": torch.float64}


def _t(a, dtype=None):
    return a if isinstance(a, torch.Tensor) and dtype is None else torch.as_tensor(a, dtype=dtype)


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    yield


def _ints(s):
    if isinstance(s, torch.Tensor):
        return tuple(int(v) for v in s.reshape(-1).tolist())
    return tuple(int(v) for v in np.asarray([int(x) for x in s] if isinstance(s, (list, tuple)) else s).reshape(-1))


def ones(shape, dtype="float32"):
    return torch.ones(_ints(shape), dtype=_DT[dtype])


def zeros(shape, dtype="float32"):
    return torch.zeros(_ints(shape), dtype=_DT[dtype])


def ones_like(x):
    return torch.ones_like(x)


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def transpose(x, perm):
    return _t(x).permute(*perm)


def cast(x, dtype):
    return _t(x).to(_DT[dtype])


def reshape(x, shape):
    return _t(x).reshape(_ints(shape))


def shape(x):
    return torch.tensor(list(x.shape), dtype=torch.int32)


def range(n):  # noqa: A001
    return torch.arange(int(n), dtype=torch.int32)


def tile(x, multiples):
    return _t(x).repeat(*_ints(multiples))


def concat(axis, values):
    return torch.cat([_t(v) for v in values], axis)


def slice(x, begin, size):  # noqa: A001
    idx = tuple(np.s_[b:(x.shape[d] if s == -1 else b + s)] for d, (b, s) in enumerate(zip(begin, size)))
    return x[idx]


def floor(x):
    return torch.floor(x)


def clip_by_value(x, lo, hi):
    return torch.minimum(torch.maximum(x, _t(lo).to(x.dtype)), _t(hi).to(x.dtype))


def gather(params, indices):
    return params[_t(indices).long()]


def linspace(start, stop, num):
    num = int(num)
    if num == 1:
        return torch.tensor([start], dtype=torch.float32)
    step = (torch.tensor(stop, dtype=torch.float32) - torch.tensor(start, dtype=torch.float32)) / torch.tensor(float(num - 1), dtype=torch.float32)
    return torch.tensor(start, dtype=torch.float32) + step * torch.arange(num, dtype=torch.float32)


def add_n(values):
    acc = values[0]
    for v in values[1:]:
        acc = acc + v
    return acc


def matmul(a, b):
    K = a.shape[-1]
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for k in np.arange(1, K):
        acc = acc + a[..., :, k:k + 1] * b[..., k:k + 1, :]
    return acc


def Variable(initial_value=None, **k):
    return _t(initial_value)
