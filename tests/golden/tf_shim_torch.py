"""The same stand-in as ``tf_shim.py`` built on torch-CPU tensors, so that ``torch.autograd`` differentiates THE
REFERENCE'S OWN GRAPH (``air/transformer.py`` imported on top of this module): gather's gradient is a scatter-add, floor
and integer casts pass no gradient -- what TF autodiff does for the same ops (``optimizer.compute_gradients``,
air_number_bbox_location.py:1098).  Per-kernel forward numerics as in ``tf_shim.py`` ([TF-1.12 assumed]).
Test infrastructure only."""
import contextlib

import numpy as np
import torch

_DT = {"float32": torch.float32, "int32": torch.int32, "float64": torch.float64}
DEFAULT = {"dtype": torch.float32}       # float dtype of constants; a generator may switch it to float64 for a yardstick run


def _t(a, dtype=None):
    return a if isinstance(a, torch.Tensor) and dtype is None else torch.as_tensor(a, dtype=dtype)


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    yield


def _ints(s):
    if isinstance(s, torch.Tensor):
        return tuple(int(v) for v in s.reshape(-1).tolist())
    return tuple(int(v) for v in np.asarray([int(x) for x in s] if isinstance(s, (list, tuple)) else s).reshape(-1))


def ones(shape, dtype=None):
    return torch.ones(_ints(shape), dtype=_DT[dtype] if dtype else DEFAULT["dtype"])


def zeros(shape, dtype=None):
    return torch.zeros(_ints(shape), dtype=_DT[dtype] if dtype else DEFAULT["dtype"])


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


def concat(values=None, axis=0, **k):
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


# ---- the additional ops of the ASR regulariser block (air_number_bbox_location.py:645-681, :970-1069) ------------------
float32 = "float32"


def convert_to_tensor(x, dtype=None):
    return torch.as_tensor(np.asarray(x), dtype=DEFAULT["dtype"])


def reduce_mean(x, axis=None):
    return x.mean() if axis is None else x.mean(axis)


def reduce_sum(x, axis=None, name=None):
    return x.sum() if axis is None else x.sum(axis)


def reduce_min(x, axis=None):
    return torch.amin(x, dim=axis)                      # ties share the gradient equally, like TF's reduce_min


def log(x):
    return torch.log(x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=DEFAULT["dtype"]))


def maximum(x, y, name=None):
    y = _t(y).to(x.dtype) if not isinstance(y, torch.Tensor) else y
    return torch.where(x >= y, x, y)                    # TF: gradient to x where x >= y, to y elsewhere [TF-1.12 assumed]


def minimum(x, y, name=None):
    y = _t(y).to(x.dtype) if not isinstance(y, torch.Tensor) else y
    return torch.where(x <= y, x, y)                    # TF: gradient to x where x <= y [TF-1.12 assumed]


def square(x):
    return x * x


def abs(x):  # noqa: A001
    return torch.abs(x)


def zeros_like(x):
    return torch.zeros_like(x)


def eye(n):
    return torch.eye(int(n), dtype=DEFAULT["dtype"])


class nn:  # noqa: N801
    @staticmethod
    def sigmoid(x):
        return torch.sigmoid(x)

    @staticmethod
    def softplus(x):
        return torch.nn.functional.softplus(x)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
        # TF's documented formulation: max(x, 0) - x * z + log(1 + exp(-|x|))
        x, z = logits, labels
        return torch.where(x >= 0, x, torch.zeros_like(x)) - x * z + torch.log1p(torch.exp(-torch.abs(x)))


class layers:  # noqa: N801
    @staticmethod
    def flatten(x):
        return x.reshape(x.shape[0], -1)


_tile_plain = tile


def tile(x, multiples):  # noqa: F811  (multiples may hold a 0-d tensor, e.g. self.batch_size)
    return _tile_plain(x, [int(m) for m in multiples])


# ---- the additional ops of the loop-body KL / stopping / canvas lines (air_number_bbox_location.py:683-787) ---------------
int32 = "int32"


def where(cond, a, b):
    # TF 1.x tf.where: a rank-1 condition selects whole ROWS of higher-rank operands
    while cond.dim() < a.dim():
        cond = cond.unsqueeze(-1)
    return torch.where(cond, a, b)


def less(a, b):
    return a < b


def exp(x):
    return torch.exp(x)


def reduce_logsumexp(x, axis=None):
    return torch.logsumexp(x, dim=axis)


class TensorArray:
    """the two methods the loop body uses: ``ta = ta.write(ta.size(), value)``"""

    def __init__(self):
        self.items = []

    def size(self):
        return len(self.items)

    def write(self, index, value):
        assert index == len(self.items)
        self.items.append(value)
        return self


# ---- gradient post-processing (air_number_bbox_location.py:1100-1111) ------------------------------------------------------
def is_inf(x):
    return torch.isinf(x)


def is_nan(x):
    return torch.isnan(x)


def clip_by_norm(t, clip_norm):
    # TF: t * clip_norm / max(||t||_2, clip_norm)  [TF-1.12 assumed]
    return t * clip_norm / torch.clamp(torch.sqrt((t * t).sum()), min=clip_norm)
