"""The reference's gradient post-processing lines (``air/air_number_bbox_location.py`` from
``if self.gradient_clipping_norm is not None:`` to just before ``grads_and_vars = ...``, :1100-1111) exec'd on the torch TF shim
for a list of gradient tensors holding infinities, NaNs, norms above and below the clipping norm and a ``None``.
Writes ``tests/golden/graph_gradpost.npz``.  Run from the repo root in the authoring container."""
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim_torch as tf  # noqa: E402

lines = open("/root/reference/air/air_number_bbox_location.py").read().split("\n")
i = next(k for k, l in enumerate(lines) if "if self.gradient_clipping_norm is not None:" in l)
j = next(k for k, l in enumerate(lines) if k > i and "grads_and_vars = list(zip(grads, variables))" in l)
SRC = textwrap.dedent("\n".join(lines[i:j]))
print("lines", (i + 1, j))
rng = np.random.default_rng(0)
shapes = [(7, 5), (5,), (64, 3), (3,), (40, 40), (40,)]
scales = [3.0, 1e-3, 10.0, 0.2, 1e-2, 50.0]
grads_in = [(rng.normal(size=s) * k).astype(np.float32) for s, k in zip(shapes, scales)]
grads_in[0][1, 2] = np.inf; grads_in[0][3, 0] = -np.inf
grads_in[2][5, 1] = np.nan
grads_in[4][0, 0] = np.nan; grads_in[4][7, 7] = np.inf
ns = dict(tf=tf, self=types.SimpleNamespace(gradient_clipping_norm=1.0), grads=tuple([torch.tensor(g) for g in grads_in] + [None]))
exec(SRC, ns)
out = ns["grads"]
assert out[-1] is None
np.savez_compressed(os.path.join(HERE, "graph_gradpost.npz"), clip=np.float32(1.0), **{f"in{k}": g for k, g in enumerate(grads_in)},
                    **{f"out{k}": o.numpy() for k, o in enumerate(out[:-1])})
print([float(torch.sqrt((o * o).sum())) for o in out[:-1]])
