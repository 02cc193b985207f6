"""Gradients of THE REFERENCE'S OWN GRAPH: ``air/transformer.py`` imported on the torch-based TF shim
(``tf_shim_torch.py``) and differentiated by ``torch.autograd`` in float32, for the ``gout`` of the golden cases.
Writes ``tests/golden/graph_grad_<case>.npz``.  Run from the repo root in the authoring container."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim_torch  # noqa: E402

sys.modules["tensorflow"] = tf_shim_torch
spec = importlib.util.spec_from_file_location("ref_transformer_torch", "/root/reference/air/transformer.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

for name in ["read_50_28", "write_28_50", "adversarial_17x23x3_9x31", "fullcover_64_28", "out_1x7"]:
    z = np.load(os.path.join(HERE, name + ".npz"))
    finite = np.isfinite(z["theta"]).all(1) & (np.abs(z["theta"]).max(1) < 1e6)
    rows = np.nonzero(finite)[0]
    U = torch.tensor(z["U"][rows], requires_grad=True)
    th = torch.tensor(z["theta"][rows], requires_grad=True)
    out = ref.transformer(U, th, tuple(int(v) for v in z["out_size"]))
    assert np.array_equal(out.detach().numpy().view(np.uint32), z["out"][rows].view(np.uint32)), name   # same forward, bit for bit
    out.backward(torch.tensor(z["gout"][rows]))
    np.savez_compressed(os.path.join(HERE, "graph_grad_" + name + ".npz"), rows=rows, dU=U.grad.numpy(), dtheta=th.grad.numpy().reshape(-1, 2, 3))
    print(name, "rows", len(rows), "|dU|max", float(U.grad.abs().max()), "|dtheta|max", float(th.grad.abs().max()))
