"""``tf_shim_torch.py`` extended to the slice of the TF 1.x API that the reference's WHOLE model graph uses
(``air/air_number_bbox_location.py::AIRModel._create_model``, ``air/vae.py``, ``air/concrete.py``, ``air/transformer.py``):
variable scopes with TF's layer-name uniquification, ``tf.layers.dense``, ``tf.contrib.layers.fully_connected``,
``tf.nn.rnn_cell.LSTMCell``, ``tf.TensorArray``, ``tf.while_loop`` (an eager Python loop), injected noise.
Library layers are DEFINED here the way TF 1.12 documents them ([TF-1.12 assumed]: dense = x W + b with kernel [in, out];
LSTMCell: one kernel [in + units, 4 units], gates i, j, f, o, forget_bias 1.0); what running the reference file on this
shim pins is the WIRING of the model: what feeds what, in which order, masked by which stopping sum, summed into which
loss.  Yardstick runs are float64 (``DEFAULT['dtype']``): casts to 'float32' keep the yardstick dtype.
Test infrastructure only."""
import collections
import contextlib
import sys
import types

import numpy as np
import torch

from tf_shim_torch import *  # noqa: F401,F403
import tf_shim_torch as _base
from tf_shim_torch import DEFAULT, _t, _ints

_base._DT["float32"] = None          # resolved at call time: the yardstick dtype
_orig_cast = _base.cast


def cast(x, dtype):
    if dtype in ("float32", float32):
        return _t(x).to(DEFAULT["dtype"]) if isinstance(x, torch.Tensor) else torch.tensor(float(x), dtype=DEFAULT["dtype"])
    return _t(x).to(torch.int32)


# ---- scopes, variables ------------------------------------------------------------------------------------------------
_SCOPE = []                      # current scope path
VARIABLES = collections.OrderedDict()
_LAYER_COUNTS = collections.Counter()
_RNG = np.random.default_rng(0)


class Scope:
    def __init__(self, name, reuse=None):
        self.name, self.reuse = name, reuse


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None, *a, **k):
    global _SCOPE
    saved = list(_SCOPE)
    if isinstance(name_or_scope, Scope):
        _SCOPE = name_or_scope.name.split("/")          # re-entering a captured scope is absolute, as in TF
    else:
        _SCOPE = _SCOPE + [name_or_scope]
    try:
        yield Scope("/".join(_SCOPE), reuse)
    finally:
        _SCOPE = saved


def _variable(name, shape, kind):
    if name not in VARIABLES:
        if kind == "kernel":
            lim = np.sqrt(6.0 / (shape[0] + shape[-1]))
            v = _RNG.uniform(-lim, lim, shape)
        else:
            v = _RNG.normal(0, 0.05, shape)          # non-zero biases: a mis-wired bias must show
        VARIABLES[name] = torch.tensor(v, dtype=DEFAULT["dtype"], requires_grad=True)
    assert tuple(VARIABLES[name].shape) == tuple(shape), (name, VARIABLES[name].shape, shape)
    return VARIABLES[name]


def _unique(scope_path, base):
    n = _LAYER_COUNTS[(scope_path, base)]
    _LAYER_COUNTS[(scope_path, base)] += 1
    return base if n == 0 else f"{base}_{n}"


class _Layers:
    @staticmethod
    def dense(inputs, units, activation=None, name=None, **k):
        path = "/".join(_SCOPE)
        lname = name or _unique(path, "dense")
        w = _variable(f"{path}/{lname}/kernel", (inputs.shape[-1], units), "kernel")
        b = _variable(f"{path}/{lname}/bias", (units,), "bias")
        y = inputs @ w + b
        return activation(y) if activation is not None else y

    @staticmethod
    def flatten(x):
        return x.reshape(x.shape[0], -1)


layers = _Layers


def fully_connected(inputs, num_outputs, activation_fn=None, scope=None, **k):
    path = scope.name if isinstance(scope, Scope) else "/".join(_SCOPE + [scope])
    w = _variable(f"{path}/weights", (inputs.shape[-1], num_outputs), "kernel")
    b = _variable(f"{path}/biases", (num_outputs,), "bias")
    y = inputs @ w + b
    return activation_fn(y) if activation_fn is not None else y


contrib_layers = types.ModuleType("tensorflow.contrib.layers")
contrib_layers.fully_connected = fully_connected
contrib = types.ModuleType("tensorflow.contrib")
contrib.layers = contrib_layers

LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


class _LSTMCell:
    def __init__(self, num_units, reuse=None, **k):
        self.units = num_units

    def zero_state(self, batch_size, dtype):
        z = torch.zeros((int(batch_size), self.units), dtype=DEFAULT["dtype"])
        return LSTMStateTuple(z, z.clone())

    def __call__(self, inputs, state, scope=None):
        path = scope.name if isinstance(scope, Scope) else "/".join(_SCOPE)
        c, h = state
        kernel = _variable(f"{path}/lstm_cell/kernel", (inputs.shape[-1] + self.units, 4 * self.units), "kernel")
        bias = _variable(f"{path}/lstm_cell/bias", (4 * self.units,), "bias")
        gates = torch.cat([inputs, h], 1) @ kernel + bias
        i, j, f, o = gates.chunk(4, 1)
        c2 = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        return h2, LSTMStateTuple(c2, h2)


class _RnnCell:
    LSTMCell = _LSTMCell


class nn(_base.nn):  # noqa: N801
    rnn_cell = _RnnCell
    relu = staticmethod(torch.relu)
    tanh = staticmethod(torch.tanh)


# ---- control flow, arrays, noise -------------------------------------------------------------------------------------------
STEP = {"t": 0}
NOISE = {"fn": None, "latent": 50}


class TensorArray(_base.TensorArray):
    def __init__(self, dtype=None, size=0, dynamic_size=True, **k):
        super().__init__()

    def stack(self):
        return torch.stack(self.items, 0)


def while_loop(cond, body, loop_vars, **k):
    vars_ = list(loop_vars)
    counts = collections.Counter(_LAYER_COUNTS)
    STEP["t"] = 0
    while bool(cond(*vars_)):
        _LAYER_COUNTS.clear()
        _LAYER_COUNTS.update(counts)              # the body is ONE graph: every iteration names its layers alike
        vars_ = list(body(*vars_))
        STEP["t"] += 1
    return vars_


def random_normal(shape, **k):
    dims = _ints(shape)
    kind = {2: "shift", 1: "scale", NOISE["latent"]: "vae"}.get(dims[1])
    if kind is None:
        return torch.zeros(dims, dtype=DEFAULT["dtype"])          # the decoder's likelihood noise (x likelihood_std = 0)
    return NOISE["fn"](kind, STEP["t"], dims).to(DEFAULT["dtype"])


def random_uniform(shape, minval=0, maxval=1, **k):
    return NOISE["fn"]("concrete", STEP["t"], _ints(shape)).to(DEFAULT["dtype"])


def constant(v, **k):
    return torch.tensor(v)


def logical_and(a, b):
    return torch.logical_and(_t(a), _t(b))


def reduce_any(x):
    return torch.any(x)


def reduce_all(x):
    return torch.all(x)


def greater(a, b):
    return a > b


def equal(a, b):
    return a == b


def sqrt(x):
    return torch.sqrt(x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=DEFAULT["dtype"]))


def round(x):  # noqa: A001
    return torch.round(x)


def less(a, b):
    return _t(a) < (b if isinstance(b, torch.Tensor) else b)


def ones(shape, dtype=None):
    return torch.ones(_ints(shape), dtype=torch.int32 if dtype in ("int32", int32) else DEFAULT["dtype"])


def zeros(shape, dtype=None):
    return torch.zeros(_ints(shape), dtype=torch.int32 if dtype in ("int32", int32) else DEFAULT["dtype"])


class _Adam:
    def __init__(self, *a, **k):
        pass

    def compute_gradients(self, loss):
        return [(None, None)]

    def apply_gradients(self, *a, **k):
        return None


train = types.SimpleNamespace(AdamOptimizer=_Adam)


def install():
    """make ``import tensorflow`` / ``import tensorflow.contrib.layers`` resolve to this shim"""
    me = sys.modules[__name__]
    torch.Tensor.set_shape = lambda self, shape: None          # static-shape hints of the graph compiler: no-ops here
    sys.modules["tensorflow"] = me
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.layers"] = contrib_layers


# ---- name= keyword / default-argument variants the model file uses ---------------------------------------------------------
def log(x, name=None):  # noqa: F811
    return _base.log(x)


def transpose(x, perm=None, name=None):  # noqa: F811
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.permute(*reversed(range(x.dim())))


def reshape(x, shape, name=None):  # noqa: F811
    return _base.reshape(x, shape)


def reduce_mean(x, axis=None, name=None):  # noqa: F811
    x = x.to(DEFAULT["dtype"]) if not x.is_floating_point() else x
    return x.mean() if axis is None else x.mean(axis)
