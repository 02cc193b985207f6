"""A numpy stand-in for the slice of the TensorFlow 1.x API that ``/root/reference/air/transformer.py`` uses, so that the
REFERENCE'S OWN SOURCE can be executed in the authoring container (TensorFlow 1.12 is not installable here).

What this pins and what it does not.  Running the reference file on this shim pins the GRAPH: which ops, in which
order, wired to which operands -- the part a hand restatement can get wrong (which corner pairs with which weight,
where the -1.001 sits, how indices are flattened).  It does not pin the numerics of individual TF kernels; those are
defined here exactly as ``oracle/stn_ref_numpy.py`` assumes them, each marked [TF-1.12 assumed]:
  * ``linspace``: ``start + step * i`` in float32, ``step = (stop - start) / (num - 1)`` in float32;
  * ``matmul`` (float32): products accumulated left to right over k, one rounding per operation, no FMA;
  * ``add_n``: ``((a + b) + c) + d``;
  * ``cast(float -> int32)`` of a floored value: exact for |v| < 2**31 (inputs outside that range are not used).
Everything else (reshape, tile, gather, clip, ...) has exact semantics.  Test infrastructure only."""
import contextlib

import numpy as np

F32, I32 = np.float32, np.int32
_DT = {"float32": F32, "int32": I32, F32: F32, I32: I32, "float64": np.float64}


class Tensor(np.ndarray):
    """ndarray with the two TF methods batch_transformer touches"""

    def get_shape(self):
        shape = self.shape

        class _S:
            def as_list(self_inner):
                return list(shape)
        return _S()


def _t(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(Tensor)


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    yield


def _shape_arg(s):
    return tuple(int(v) for v in np.asarray(s).reshape(-1)) if not isinstance(s, (tuple, list)) or len(s) else ()


def ones(shape, dtype="float32"):
    return _t(np.ones(_shape_arg(shape), _DT[dtype]))


def zeros(shape, dtype="float32"):
    return _t(np.zeros(_shape_arg(shape), _DT[dtype]))


def ones_like(x):
    return _t(np.ones_like(np.asarray(x)))


def stack(values, axis=0):
    return _t(np.stack([np.asarray(v) for v in values], axis))


def expand_dims(x, axis):
    return _t(np.expand_dims(np.asarray(x), axis))


def transpose(x, perm):
    return _t(np.transpose(np.asarray(x), perm))


def cast(x, dtype):
    return _t(np.asarray(x).astype(_DT[dtype]))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), tuple(int(v) for v in np.asarray(shape).reshape(-1))))


def shape(x):
    return np.asarray(np.asarray(x).shape, I32)


def range(n):  # noqa: A001 (mirrors tf.range)
    return _t(np.arange(int(n), dtype=I32))


def tile(x, multiples):
    return _t(np.tile(np.asarray(x), tuple(int(v) for v in np.asarray(multiples).reshape(-1))))


def concat(values=None, axis=0, **k):
    return _t(np.concatenate([np.asarray(v) for v in values], axis))


def slice(x, begin, size):  # noqa: A001 (mirrors tf.slice)
    x = np.asarray(x)
    idx = tuple(np.s_[b:(x.shape[d] if s == -1 else b + s)] for d, (b, s) in enumerate(zip(begin, size)))
    return _t(x[idx])


def floor(x):
    return _t(np.floor(np.asarray(x)))


def clip_by_value(x, lo, hi):
    return _t(np.clip(np.asarray(x), np.asarray(lo), np.asarray(hi)))


def gather(params, indices):
    return _t(np.take(np.asarray(params), np.asarray(indices), axis=0))


def linspace(start, stop, num):
    num = int(num)
    if num == 1:
        return _t(np.array([start], F32))
    step = F32(F32(stop) - F32(start)) / F32(num - 1)                       # [TF-1.12 assumed]
    return _t(F32(start) + step * np.arange(num, dtype=F32))


def add_n(values):
    acc = np.asarray(values[0])
    for v in values[1:]:
        acc = acc + np.asarray(v)                                            # [TF-1.12 assumed] left to right
    return _t(acc)


def matmul(a, b):
    a, b = np.asarray(a), np.asarray(b)
    K = a.shape[-1]
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for k in np.arange(1, K):                                                # [TF-1.12 assumed] sequential over k, no FMA
        acc = acc + a[..., :, k:k + 1] * b[..., k:k + 1, :]
    return _t(acc)


def Variable(initial_value=None, **k):
    return _t(initial_value)


# ---- the additional ops of air/concrete.py (float64 in, float64 out: used as a yardstick, compared at a tolerance) --------
UNIFORM_QUEUE = []          # random_uniform pops preset draws from here: the reference's noise is injected, not replayed


def random_uniform(shape, minval=0, maxval=1, **k):
    u = np.asarray(UNIFORM_QUEUE.pop(0))
    assert tuple(u.shape) == tuple(int(v) for v in np.asarray(shape).reshape(-1))
    return _t(u)


def log(x):
    return _t(np.log(np.asarray(x)))


def exp(x):
    return _t(np.exp(np.asarray(x)))


def zeros_like(x):
    return _t(np.zeros_like(np.asarray(x)))


def reduce_logsumexp(x, axis=None):
    x = np.asarray(x)
    m = np.max(x, axis=axis, keepdims=True)
    return _t(np.squeeze(m, axis) + np.log(np.sum(np.exp(x - m), axis=axis)))


def round(x):  # noqa: A001
    return _t(np.round(np.asarray(x)))


def stop_gradient(x):
    return x


class nn:  # noqa: N801 (mirrors tf.nn)
    @staticmethod
    def sigmoid(x):
        return _t(1.0 / (1.0 + np.exp(-np.asarray(x))))
