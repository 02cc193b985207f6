"""ORACLE (test infrastructure, NOT product code) -- an ``ops`` object for ``mog_asr_b200.air.AIRModel`` whose
three hot-path operators are the oracle's torch restatements (``oracle/stn_ref_torch.py``,
``oracle/asr_ref.py``) and plain reference ops for the composite (``air/air_number_bbox_location.py:722-727``).
Used by tests/ (parity of the training step) and by bench.py's CPU-baseline leg; the product never imports it.
Pinned to the reference's own graph (not to a live TensorFlow): the re-hosted model driven by these operators reproduces the
reference's ``AIRModel._create_model`` executed on the torch TF shim -- loss, loop trip count, counts, canvas and every weight
gradient (``tests/golden/make_golden_model.py``, ``tests/test_air_model.py``)."""
import torch
import torch.distributed as dist

from oracle import asr_ref, stn_ref_torch


class OracleOps:
    def __init__(self, process_group=None, global_batch=None):
        self.pg, self.global_batch = process_group, global_batch

    def transformer(self, U, theta, out_size):
        return stn_ref_torch.transformer(U, theta, out_size)

    def write_composite(self, canvas, window, theta, z_pres, stop_sum, threshold):
        B, cs = canvas.shape[0], canvas.shape[1]
        win = stn_ref_torch.transformer(window[..., None], theta, (cs, cs))[..., 0]
        return canvas + torch.where((stop_sum < threshold)[:, None, None], z_pres[:, None, None] * win, torch.zeros_like(win))

    def read_sxy(self, images4, inf_shift, inf_scale, out_size):
        """the read call site as the model states it (theta from (s, x, y), :511-542); returns (window, shift, scale)"""
        theta_r, _ = self.thetas(inf_shift, inf_scale)
        return self.transformer(images4, theta_r, out_size), inf_shift, inf_scale

    def write_composite_sxy(self, canvas, window, shift, scale, z_pres, stop_sum, threshold):
        _, theta_w = self.thetas(shift, scale)                                                 # :563-584
        return self.write_composite(canvas, window, theta_w, z_pres, stop_sum, threshold)

    def recon_loss(self, images, canvas):
        r = torch.clamp(canvas, 0.0, 1.0)                                   # air_number_bbox_location.py:947-948
        return -(images * torch.log(r + 1e-10) + (1.0 - images) * torch.log(1.0 - r + 1e-10)).sum(1)   # :954-959

    # per-step elementwise math, written out op for op like the reference
    def gauss_sample(self, mean, logvar, eps, act=None):
        latent = mean + eps * torch.sqrt(torch.exp(logvar))            # _sample_from_mvn, air_number_bbox_location.py:180-184
        if act == "tanh":
            return latent, torch.tanh(latent)                          # :435
        if act == "sigmoid":
            return latent, torch.sigmoid(latent)                       # :458
        return latent, None

    def thetas(self, inf_shift, inf_scale):
        s, x, y = inf_scale[:, 0], inf_shift[:, 0], inf_shift[:, 1]
        zero = torch.zeros_like(s)
        theta_r = torch.stack([s, zero, x, zero, s, y], 1)                                   # :511-531
        theta_w = torch.stack([1.0 / s, zero, -x / s, zero, 1.0 / s, -y / s], 1)             # :563-584
        return theta_r, theta_w

    def zpres(self, log_odds, u, stop_sum, temperature, threshold):
        y_pre = (log_odds + torch.log(u + 10e-10) - torch.log(1.0 - u + 10e-10)) / temperature   # concrete.py:20-27
        z_pres = torch.sigmoid(y_pre)                                                        # :631
        active_prev = stop_sum < threshold                                                   # :698-702 (previous stop_sum)
        stop_new = stop_sum + (1.0 - z_pres)                                                 # :712
        return y_pre, z_pres, stop_new, active_prev, stop_new < threshold

    def asr(self, cfg, log_odds, shifts, scales):
        pbar = None
        if cfg.constrains_margin_gamma > 1e-8:
            psum = torch.sigmoid(log_odds).detach().sum(0)
            n = log_odds.shape[0]
            if self.pg is not None:
                dist.all_reduce(psum, group=self.pg)
                n = self.global_batch or n * dist.get_world_size(self.pg)
            # value from the global sum, gradient through the local rows: d pbar / d P[b,t] = 1 / n_global
            local = torch.sigmoid(log_odds).sum(0)
            pbar = (psum - local.detach() + local) / n
        r = asr_ref.asr_terms(log_odds, shifts, scales[..., 0], canvas_size=cfg.canvas_size, counts=list(cfg.constrains_num),
                              max_steps=cfg.max_steps, gamma_num=cfg.constrains_num_gamma,
                              gamma_margin=cfg.constrains_margin_gamma, gamma_elem=cfg.constrains_num_element_gamma,
                              gamma_bbox=cfg.constrains_bbox_gamma, gamma_size=cfg.constrains_sharesize_gamma,
                              gamma_area=cfg.constrains_area_gamma, area_minmax=tuple(cfg.constrains_area_minmax),
                              batch_prob_mean=pbar)
        comps = torch.stack([r[k] for k in ("pr_num", "num_min", "area", "out", "size", "overlap")], 1).detach()
        return r["per_image"], r["margin"], comps


class SeededNoise:
    """noise(kind, step, shape) drawn on the CPU from a seed per (kind, step) for the *global* batch, then the
    rank's rows are sliced out -- identical draws for any implementation / sharding."""

    def __init__(self, seed, global_batch, lo=0, hi=None, device="cpu", dtype=torch.float32):
        self.seed, self.B, self.lo, self.hi, self.device, self.dtype = seed, global_batch, lo, hi or global_batch, device, dtype

    def __call__(self, kind, step, shape):
        # numpy's generator: the draw must not depend on torch's intra-op thread count
        import numpy as np
        g = np.random.default_rng(self.seed * 1000 + {"shift": 1, "scale": 2, "vae": 3, "concrete": 4}[kind] * 100 + step)
        full = (self.B,) + tuple(shape[1:])
        t = g.random(full, dtype=np.float32) if kind == "concrete" else g.standard_normal(full, dtype=np.float32)
        if kind == "concrete":
            t = np.clip(t, 1e-4, 1 - 1e-4)
        return torch.from_numpy(t[self.lo:self.hi]).to(self.device, self.dtype)
