"""More of the reference's own lines exec'd on the torch-based TF shim (``tf_shim_torch.py``):
  * the reconstruction-loss block ``with tf.variable_scope("loss/reconstruction")`` (air_number_bbox_location.py:944-967),
    float64, with the gradient of ``sum_b w_b * reconstruction_loss_b`` w.r.t. the un-clipped canvas;
  * the two theta constructions ``"st_forward"`` (:511-531) and ``"st_backward"`` (:563-584), float32 (the fused kernel
    is expected to reproduce the fp32 divides bit for bit) with the gradients of a weighted sum of all entries.
Writes ``tests/golden/graph_recon.npz`` and ``tests/golden/graph_thetas.npz``.  Run from the repo root in the authoring
container; the lines are read from the reference file, nothing is copied into the repo."""
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


def block(start_marker, end_marker):
    i = next(k for k, l in enumerate(lines) if start_marker in l)
    j = next(k for k, l in enumerate(lines) if k > i and end_marker in l)
    return textwrap.dedent("\n".join(lines[i:j])), (i + 1, j)


# ---- reconstruction loss ---------------------------------------------------------------------------------------------
src, span = block('with tf.variable_scope("loss/reconstruction"):', "# adding reconstruction loss")
print("reconstruction lines", span)
tf.DEFAULT["dtype"] = torch.float64
rng = np.random.default_rng(11)
B, P = 6, 400
canvas = rng.normal(0.4, 0.5, (B, P))
canvas[0, :50] = 0.0            # exactly on the lower clip bound (gradient 1e10 under an object pixel)
canvas[1, :50] = 1.0            # exactly on the upper bound
canvas[2] = rng.random(P) * 1.4  # sums of overlapping objects exceed 1
images = np.clip(rng.random((B, P)) * 1.2 - 0.1, 0, 1)
images[:, ::7] = 0.0
c = torch.tensor(canvas, requires_grad=True)
self = types.SimpleNamespace(input_images=torch.tensor(images), log_variables={})
ns = dict(tf=tf, np=np, self=self, reconstruction=c, elbo=0.0)
exec(src, ns)
w = torch.tensor(rng.normal(size=B))
(self.reconstruction_loss * w).sum().backward()
np.savez_compressed(os.path.join(HERE, "graph_recon.npz"), canvas=canvas, images=images, w=w.numpy(), loss=self.reconstruction_loss.detach().numpy(),
                    mse=self.mse_loss.detach().numpy(), clipped=self.reconstruction.detach().numpy(), dcanvas=c.grad.numpy())
print("recon loss", self.reconstruction_loss.detach().numpy()[:3], "max |dcanvas|", float(c.grad.abs().max()))

# ---- theta construction ----------------------------------------------------------------------------------------------
fwd_src, s1 = block('with tf.variable_scope("st_forward"):', "# ST forward transformation: canvas -> window")
bwd_src, s2 = block('with tf.variable_scope("st_backward"):', "# collecting backward transformation matrices of ST")
print("theta lines", s1, s2)
tf.DEFAULT["dtype"] = torch.float32
n = 257
s = torch.tensor((1 / (1 + np.exp(-rng.normal(-1, 0.8, n)))).astype(np.float32), requires_grad=True)
x = torch.tensor(np.tanh(rng.normal(0, 1, n)).astype(np.float32), requires_grad=True)
y = torch.tensor(np.tanh(rng.normal(0, 1, n)).astype(np.float32), requires_grad=True)
ns = dict(tf=tf, inf_s=s, inf_x=x, inf_y=y)
exec(fwd_src, ns)
exec(bwd_src, ns)
theta, theta_recon = ns["theta"], ns["theta_recon"]
wr, ww = torch.tensor(rng.normal(size=(n, 2, 3)).astype(np.float32)), torch.tensor(rng.normal(size=(n, 2, 3)).astype(np.float32))
((theta * wr).sum() + (theta_recon * ww).sum()).backward()
np.savez_compressed(os.path.join(HERE, "graph_thetas.npz"), s=s.detach().numpy(), x=x.detach().numpy(), y=y.detach().numpy(), wr=wr.numpy(), ww=ww.numpy(),
                    theta=theta.detach().numpy(), theta_recon=theta_recon.detach().numpy(), ds=s.grad.numpy(), dx=x.grad.numpy(), dy=y.grad.numpy())
print("theta", theta[0].tolist(), "theta_recon", theta_recon[0].tolist())
