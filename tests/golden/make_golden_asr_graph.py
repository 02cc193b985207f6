"""Execute the reference's own regulariser source -- the entropy lines of the loop body
(``air/air_number_bbox_location.py`` ``with tf.variable_scope("loss/pr_num")``, :645-678) and the whole post-loop block from
``"loss/pr_num_margin"`` to just before ``"accuracy"`` (:970-1069) -- on the torch-based TF shim in float64, and differentiate
``reduce_mean(pr_loss + num_element_min) + num_marginal_loss`` (:1078-1079 without the ELBO) by autograd.
The lines are read from the reference file, dedented and exec'd against a stand-in for ``self``; nothing is copied into
the repo.  Writes ``tests/golden/graph_asr_<config>.npz``.  Run from the repo root in the authoring container."""
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim_torch as tf  # noqa: E402

REF = "/root/reference/air/air_number_bbox_location.py"
lines = open(REF).read().split("\n")


def block(start_marker, end_marker, skip=0):
    i = next(k for k, l in enumerate(lines) if start_marker in l)
    j = next(k for k, l in enumerate(lines) if k > i and end_marker in l)
    return textwrap.dedent("\n".join(lines[i + skip:j])), (i + 1 + skip, j)


ENTROPY_SRC, ent_span = block('with tf.variable_scope("loss/pr_num"):', 'running_pr_loss["pr_num"] = running_pr_loss["pr_num"].write(')
POST_SRC, post_span = block('with tf.variable_scope("loss/pr_num_margin"):', 'with tf.variable_scope("accuracy"):')
print("entropy lines", ent_span, "post-loop lines", post_span)

CONFIGS = {
    "c2": dict(canvas_size=50, max_steps=6, counts=[1, 3], gn=0.0, gm=100.0, gne=10.0, gb=0.0, gs=0.0, ga=0.0, minmax=(17.0, 23.0)),
    "c3": dict(canvas_size=64, max_steps=6, counts=[3], gn=0.0, gm=0.0, gne=0.0, gb=1.0, gs=10.0, ga=20.0, minmax=(12.0, 15.0)),
    "all": dict(canvas_size=50, max_steps=6, counts=[1, 2, 4], gn=0.7, gm=3.0, gne=2.0, gb=1.5, gs=0.5, ga=0.25, minmax=(11.0, 15.0)),
}


def run(cfg, B, T, seed):
    tf.DEFAULT["dtype"] = torch.float64
    rng = np.random.default_rng(seed)
    lo = torch.tensor(rng.normal(0, 2, (B, T)), requires_grad=True)
    sh = torch.tensor(np.tanh(rng.normal(0, 1, (B, T, 2))), requires_grad=True)
    sc = torch.tensor(1 / (1 + np.exp(-rng.normal(-1, 0.7, (B, T, 1)))), requires_grad=True)
    self = types.SimpleNamespace(constrains_num_gamma=cfg["gn"], constrains_margin_gamma=cfg["gm"], constrains_num_element_gamma=cfg["gne"],
                                 constrains_bbox_gamma=cfg["gb"], constrains_sharesize_gamma=cfg["gs"], constrains_area_gamma=cfg["ga"],
                                 constrains_area_minmax=cfg["minmax"], constrains_num_list=cfg["counts"], max_steps=cfg["max_steps"],
                                 canvas_size=cfg["canvas_size"], batch_size=torch.tensor(B), log_variables={})
    # loop body, once per executed step (:641 then :645-678); pr_loss = sum over steps (:936-942)
    pr_loss = 0.0
    probs = []
    for t in range(T):
        ns = dict(tf=tf, np=np, self=self, z_pres_log_odds=lo[:, t])
        ns["z_pres_prob"] = tf.nn.sigmoid(ns["z_pres_log_odds"])                                  # :641
        exec(ENTROPY_SRC, ns)
        pr_loss = pr_loss + ns["pr_num_loss"]
        probs.append(ns["z_pres_prob"])
    pr_num = pr_loss if isinstance(pr_loss, torch.Tensor) else torch.zeros(B, dtype=torch.float64)
    self.z_pres_probs = torch.stack(probs, 1)                                                     # :928  [B, T]
    self.rec_scales, self.rec_shifts = sc, sh                                                     # :922-925  [B, T, 1], [B, T, 2]
    ns = dict(tf=tf, np=np, self=self, pr_loss=pr_num, print=lambda *a, **k: None)
    exec(POST_SRC, ns)
    pr_loss = ns["pr_loss"]
    loss = tf.reduce_mean(pr_loss + self.constrains_num_element_min) + self.constrains_num_marginal_loss      # :1078-1079
    loss.backward()
    elem = self.constrains_num_element_min
    g = lambda t: t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape))
    flt = lambda v: float(v.detach()) if isinstance(v, torch.Tensor) else float(v)
    out = dict(log_odds=lo.detach().numpy(), shifts=sh.detach().numpy(), scales=sc.detach().numpy(),
               per_image=(pr_loss + elem).detach().numpy() * np.ones(B), margin=flt(self.constrains_num_marginal_loss),
               area=self.area_loss.detach().numpy(), out_loss=self.out_loss.detach().numpy(), size=self.size_loss.detach().numpy(),
               overlap=self.overlap_loss.detach().numpy(), loss=flt(loss),
               d_log_odds=g(lo), d_shifts=g(sh), d_scales=g(sc))
    return out


for name, cfg in CONFIGS.items():
    for T in (6, 3):
        r = run(cfg, 48, T, seed=len(name) * 10 + T)
        np.savez_compressed(os.path.join(HERE, f"graph_asr_{name}_T{T}.npz"), cfg_keys=np.array(list(cfg.keys())),
                            cfg_vals=np.array([repr(v) for v in cfg.values()]), **r)
        print(name, T, "loss", r["loss"], "margin", r["margin"], "|dlo|", np.abs(r["d_log_odds"]).max())
