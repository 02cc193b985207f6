"""The KL / stopping-sum / canvas lines of the reference's loop body (``air/air_number_bbox_location.py`` from
``with tf.variable_scope("loss/z_pres_kl")`` to just before ``# explicating the shape``, :683-787) exec'd once per step on the
torch-based TF shim in float64, with the reference's own ``air/concrete.py`` imported on the same shim for the Concrete KL.
Pins: which stopping sum masks which term (the previous one for z_pres, the updated one for the rest and for the canvas
write), the three Gaussian KL formulas, the digit count, and how the canvas accumulates ``z_pres * window``.
Writes ``tests/golden/graph_loop_kl.npz``.  Run from the repo root in the authoring container."""
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim_torch as tf  # noqa: E402

tf.DEFAULT["dtype"] = torch.float64
sys.modules["tensorflow"] = tf
spec = importlib.util.spec_from_file_location("ref_concrete_torch", "/root/reference/air/concrete.py")
concrete = importlib.util.module_from_spec(spec)
spec.loader.exec_module(concrete)

lines = open("/root/reference/air/air_number_bbox_location.py").read().split("\n")
i = next(k for k, l in enumerate(lines) if 'with tf.variable_scope("loss/z_pres_kl"):' in l)
j = next(k for k, l in enumerate(lines) if k > i and '# explicating the shape of "batch-sized"' in l)
SRC = textwrap.dedent("\n".join(lines[i:j]))
print("loop-body lines", (i + 1, j))

rng = np.random.default_rng(21)
B, T, L, cs = 40, 6, 50, 12
t64 = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
inp = dict(y=t64(rng.normal(0, 8, (T, B))), prior_lo=t64(rng.normal(0, 2, (T, B))), post_lo=t64(rng.normal(0, 2, (T, B))),
           sc_mean=t64(rng.normal(-1, 0.5, (T, B, 1))), sc_lv=t64(rng.normal(-2, 0.5, (T, B, 1))),
           sh_mean=t64(rng.normal(0, 1, (T, B, 2))), sh_lv=t64(rng.normal(-1, 0.5, (T, B, 2))),
           g_sh_mean=t64(rng.normal(0, 1, (T, B, 2))), g_sh_lv=t64(rng.normal(-1, 0.5, (T, B, 2))),
           v_mean=t64(rng.normal(0, 1, (T, B, L))), v_lv=t64(rng.normal(-1, 0.5, (T, B, L))),
           window=t64(rng.random((T, B, cs, cs))))
cfgv = dict(z_pres_temperature=0.1, stopping_threshold=0.9, canvas_size=cs, vae_prior_mean=0.0, vae_prior_variance=1.0,
            scale_prior_mean=-1.0, scale_prior_variance=0.05)
self = types.SimpleNamespace(**cfgv, vae_prior_log_variance=torch.log(torch.tensor(cfgv["vae_prior_variance"], dtype=torch.float64)))
running_loss = {k: tf.TensorArray() for k in ("z_pres_kl", "scale_kl", "shift_kl", "vae_kl")}
stopping_sum = torch.zeros(B, dtype=torch.float64)
running_digits = torch.zeros(B, dtype=torch.int32)
running_recon = torch.zeros(B, cs * cs, dtype=torch.float64)
gen_scale_variance = torch.tensor(cfgv["scale_prior_variance"], dtype=torch.float64)
for t in range(T):
    ns = dict(tf=tf, self=self, concrete_binary_kl_mc_sample=concrete.concrete_binary_kl_mc_sample, running_loss=running_loss,
              stopping_sum=stopping_sum, running_digits=running_digits, running_recon=running_recon,
              z_pres_pre_sigmoid=inp["y"][t], z_pres_prior_log_odds=inp["prior_lo"][t], z_pres_log_odds=inp["post_lo"][t],
              z_pres=torch.sigmoid(inp["y"][t]),                                                       # :631
              window_recon=inp["window"][t],
              gen_scale_mean=cfgv["scale_prior_mean"], gen_scale_variance=gen_scale_variance,
              gen_scale_log_variance=torch.log(gen_scale_variance),                                    # :505-509 (fixed scale prior)
              scale_mean=inp["sc_mean"][t], scale_log_variance=inp["sc_lv"][t], scale_variance=torch.exp(inp["sc_lv"][t]),
              shift_mean=inp["sh_mean"][t], shift_log_variance=inp["sh_lv"][t], shift_variance=torch.exp(inp["sh_lv"][t]),
              gen_shift_mean=inp["g_sh_mean"][t], gen_shift_log_variance=inp["g_sh_lv"][t], gen_shift_variance=torch.exp(inp["g_sh_lv"][t]),
              vae_mean=inp["v_mean"][t], vae_log_variance=inp["v_lv"][t])
    exec(SRC, ns)
    stopping_sum, running_digits, running_recon = ns["stopping_sum"], ns["running_digits"], ns["running_recon"]
kl = {k: torch.stack(v.items, 0) for k, v in running_loss.items()}                                # [T, B]
elbo_kl = sum(v.sum(0) for v in kl.values())                                                       # :930-935
w = torch.tensor(rng.normal(size=B))
wc = torch.tensor(rng.normal(size=(B, cs * cs)))
((elbo_kl * w).sum() + (running_recon * wc).sum()).backward()
out = {k: v.detach().numpy() for k, v in inp.items()}
out.update({"kl_" + k: v.detach().numpy() for k, v in kl.items()})
out.update({"d_" + k: (v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape))) for k, v in inp.items()})
out.update(stop_sum=stopping_sum.detach().numpy(), digits=running_digits.numpy(), canvas=running_recon.detach().numpy(), w=w.numpy(), wc=wc.numpy(),
           elbo_kl=elbo_kl.detach().numpy(), cfg_keys=np.array(list(cfgv.keys())), cfg_vals=np.array(list(cfgv.values()), np.float64))
np.savez_compressed(os.path.join(HERE, "graph_loop_kl.npz"), **out)
print("digits", np.bincount(running_digits.numpy()), "elbo_kl", elbo_kl[:3].tolist())
