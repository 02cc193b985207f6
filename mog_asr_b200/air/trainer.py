"""One AIR-ASR training step, optionally data-parallel.

Reference: ``AIRModel.training`` = Adam(lr).apply_gradients after per-variable ``inf -> 0``, ``nan -> 0``,
``clip_by_norm(g, gradient_clipping_norm)`` (``air/air_number_bbox_location.py:1094-1122``); the host loop
``train_air_pr.py:292-295`` runs one such step per ``sess.run``.

Data parallelism (absent from the reference): the batch is sharded across ranks, parameters and
Adam state are replicated, and the step has exactly one exchange: an NCCL all-reduce(sum) of the flat fp32
gradient buffer *before* the per-variable post-processing (``:1100-1111``), plus the ``[T]`` column sums of
the marginal count penalty (``:982``) and a one-flag ``any`` per loop iteration (``:389``).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .model import AIRConfig, AIRModel, CudaOps


class Trainer:
    def __init__(self, cfg: AIRConfig, device, process_group=None, global_batch: Optional[int] = None, seed: int = 1235,
                 ops=None, dtype=torch.float32):
        self.cfg, self.device, self.pg = cfg, torch.device(device), process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.global_batch = global_batch
        torch.manual_seed(seed)                       # identical initial parameters on every rank
        if ops is None:
            ops = CudaOps(process_group=process_group, global_batch=global_batch)
        self.model = AIRModel(cfg, ops).to(self.device, dtype)   # fp32 is the product; fp64 only for test oracles
        self.params = [p for p in self.model.parameters()]
        # flat gradient bucket: every p.grad is a view into it -> one all-reduce per step
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, dtype=dtype, device=self.device)
        off = 0
        for p in self.params:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.grads = [p.grad for p in self.params]
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-8   # tf.train.AdamOptimizer defaults
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed + 1 + (dist.get_rank(process_group) if process_group is not None else 0))

    def num_gradient_floats(self) -> int:
        return self.flat_grad.numel()

    def _noise(self, kind, step, shape):
        if kind == "concrete":
            return torch.rand(shape, device=self.device, generator=self.gen)
        return torch.randn(shape, device=self.device, generator=self.gen)

    def _any_reduce(self, flag):
        if self.pg is None:
            return flag
        f = flag.to(torch.int32)
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.pg)
        return f

    def forward_backward(self, images, noise=None, recon_loss_fn=None):
        B = images.shape[0]
        nglobal = self.global_batch if self.global_batch is not None else B * self.world
        self.flat_grad.zero_()
        out = self.model(images, noise=noise or self._noise, any_reduce=self._any_reduce if self.pg is not None else None,
                         global_batch=nglobal, recon_loss_fn=recon_loss_fn)
        out["loss"].backward()
        return out

    def reduce_gradients(self):
        if self.pg is not None and self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)

    def postprocess_and_apply(self):
        """:1100-1111 per-variable inf/nan -> 0 and clip_by_norm, then TF's Adam update."""
        g = self.flat_grad
        torch.nan_to_num_(g, nan=0.0, posinf=0.0, neginf=0.0)
        c = self.cfg.gradient_clipping_norm
        if c is not None:
            norms = torch._foreach_norm(self.grads)
            # tf.clip_by_norm: g * clip / max(norm, clip)
            scales = [c / torch.clamp(n, min=c) for n in norms]
            torch._foreach_mul_(self.grads, scales)
        self.t += 1
        b1, b2 = self.beta1, self.beta2
        lr_t = self.cfg.learning_rate * (1.0 - b2 ** self.t) ** 0.5 / (1.0 - b1 ** self.t)
        torch._foreach_mul_(self.m, b1); torch._foreach_add_(self.m, self.grads, alpha=1.0 - b1)
        torch._foreach_mul_(self.v, b2); torch._foreach_addcmul_(self.v, self.grads, self.grads, value=1.0 - b2)
        denom = torch._foreach_sqrt(self.v)
        torch._foreach_add_(denom, self.eps)
        with torch.no_grad():
            torch._foreach_addcdiv_(self.params, self.m, denom, value=-lr_t)

    def step(self, images, noise=None):
        out = self.forward_backward(images, noise)
        self.reduce_gradients()
        self.postprocess_and_apply()
        return out
