"""GPU parity tests of the fused ASR regulariser kernel vs the oracle (fp32 tolerance: rtol 2e-4 on
values, gradients relative to their largest entry -- transcendental-heavy fp32 vs an fp64 oracle)."""
import numpy as np
import pytest
import torch

import mog_asr_b200 as M
from oracle import asr_ref
from tests import helpers as H

pytestmark = pytest.mark.gpu

VAL_RTOL = 2e-4
GRAD_RTOL = 5e-4


def make_cfg(cfg):
    return M.AsrRegulariser(
        canvas_size=cfg["canvas_size"], max_steps=cfg["max_steps"], constrains_num=cfg["counts"],
        constrains_num_gamma=cfg.get("gamma_num", 0.0), constrains_margin_gamma=cfg.get("gamma_margin", 0.0),
        constrains_num_element_gamma=cfg.get("gamma_elem", 0.0), constrains_bbox_gamma=cfg.get("gamma_bbox", 0.0),
        constrains_sharesize_gamma=cfg.get("gamma_size", 0.0), constrains_area_gamma=cfg.get("gamma_area", 0.0),
        constrains_area_minmax=cfg["area_minmax"])


def close(got, ref, rtol, name):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = np.maximum(np.abs(ref), np.max(np.abs(ref)) * 1e-2 + 1e-6)
    assert np.max(np.abs(got - ref) / scale) <= rtol, (name, np.max(np.abs(got - ref) / scale))


@pytest.mark.parametrize("name", H.ASR_CASES)
def test_asr_golden(cuda_device, name):
    d = H.load(name)
    cfg = H.asr_cfg(d)
    dev = cuda_device
    lo = torch.tensor(d["log_odds"], device=dev, requires_grad=True)
    sh = torch.tensor(d["shifts"], device=dev, requires_grad=True)
    sc = torch.tensor(d["scales"], device=dev, requires_grad=True)
    per_image, margin, comps = M.asr_regularisers(make_cfg(cfg), lo, sh, sc)
    close(per_image.detach().cpu().numpy(), d["per_image"], VAL_RTOL, "per_image")
    close(margin.detach().cpu().numpy(), d["margin"], VAL_RTOL, "margin")
    for i, k in enumerate(("pr_num", "num_min", "area", "out", "size", "overlap")):
        close(comps[:, i].cpu().numpy(), d[k], VAL_RTOL, k)
    total = (per_image * torch.tensor(d["g_per_image"], device=dev)).sum() + float(d["g_margin"]) * margin
    total.backward()
    close(lo.grad.cpu().numpy(), d["d_log_odds"], GRAD_RTOL, "d_log_odds")
    close(sh.grad.cpu().numpy(), d["d_shifts"], GRAD_RTOL, "d_shifts")
    close(sc.grad.cpu().numpy(), d["d_scales"], GRAD_RTOL, "d_scales")


def test_asr_c3_batch_256_seeded(cuda_device):
    """BASELINE config 3: dSprites -dn 3 -ds bbox20k -gb 1 -gs 10 -ga 20, batch 256, fix_steps = 3."""
    rng = np.random.default_rng(33)
    B, T = 256, 3
    cfg = dict(canvas_size=64, counts=[3], max_steps=6, gamma_bbox=1.0, gamma_size=10.0, gamma_area=20.0,
               area_minmax=(12.0, 15.0))
    lo = rng.normal(0, 2, (B, T)).astype(np.float32)
    sh = np.tanh(rng.normal(0, 1, (B, T, 2))).astype(np.float32)
    sc = (1 / (1 + np.exp(-rng.normal(-1, 0.5, (B, T, 1))))).astype(np.float32)
    ref = asr_ref.asr_numpy(lo, sh, sc[..., 0], **cfg)
    t = [torch.tensor(a, device=cuda_device, requires_grad=True) for a in (lo, sh, sc)]
    per_image, margin, _ = M.asr_regularisers(make_cfg(cfg), *t)
    (per_image.mean() + margin).backward()
    close(per_image.detach().cpu().numpy(), ref["per_image"], VAL_RTOL, "per_image")
    close(t[1].grad.cpu().numpy(), ref["d_shifts"], GRAD_RTOL, "d_shifts")
    close(t[2].grad.cpu().numpy()[..., 0], ref["d_scales"], GRAD_RTOL, "d_scales")
    assert float(margin.detach()) == 0.0 and torch.all(t[0].grad == 0)


@pytest.mark.parametrize("name", [f"graph_asr_{c}_T{t}" for c in ("c2", "c3", "all") for t in (6, 3)])
def test_asr_kernel_vs_the_reference_source_run_on_the_tf_shim(cuda_device, name):
    """The fused kernel directly against ``tests/golden/graph_asr_*.npz`` -- the reference's own regulariser lines exec'd on
    the torch TF shim in float64 (tests/golden/make_golden_asr_graph.py); loss = mean(per_image) + margin (:1078-1079)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    cfg = {str(k): eval(str(v), {"__builtins__": {}}) for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    reg = M.AsrRegulariser(canvas_size=cfg["canvas_size"], max_steps=cfg["max_steps"], constrains_num=cfg["counts"],
                           constrains_num_gamma=cfg["gn"], constrains_margin_gamma=cfg["gm"], constrains_num_element_gamma=cfg["gne"],
                           constrains_bbox_gamma=cfg["gb"], constrains_sharesize_gamma=cfg["gs"], constrains_area_gamma=cfg["ga"],
                           constrains_area_minmax=cfg["minmax"])
    dev = cuda_device
    lo = torch.tensor(g["log_odds"], dtype=torch.float32, device=dev, requires_grad=True)
    sh = torch.tensor(g["shifts"], dtype=torch.float32, device=dev, requires_grad=True)
    sc = torch.tensor(g["scales"], dtype=torch.float32, device=dev, requires_grad=True)
    per_image, margin, comps = M.asr_regularisers(reg, lo, sh, sc)
    close(per_image.detach().cpu().numpy(), g["per_image"], VAL_RTOL, "per_image")
    close(margin.detach().cpu().numpy(), g["margin"], VAL_RTOL, "margin")
    for i, k in ((2, "area"), (3, "out_loss"), (4, "size"), (5, "overlap")):
        close(comps[:, i].cpu().numpy(), g[k], VAL_RTOL, k)
    (per_image.mean() + margin).backward()
    close(lo.grad.cpu().numpy() if lo.grad is not None else np.zeros_like(g["d_log_odds"]), g["d_log_odds"], GRAD_RTOL, "d_log_odds")
    close(sh.grad.cpu().numpy(), g["d_shifts"], GRAD_RTOL, "d_shifts")
    close(sc.grad.cpu().numpy(), g["d_scales"], GRAD_RTOL, "d_scales")
