"""AIR-ASR training-step throughput (BASELINE configs 2-4) on synthetic canvases: images/sec."""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .. import synth
from .model import config_from_flags
from .trainer import Trainer

CONFIGS = {
    # name: (flags, global batch)
    "C2": (dict(data="mnist", dn="13", gm=100.0, gne=10.0), 64),
    "C3": (dict(data="sprites", dn="3", ds="bbox20k", gb=1.0, gs=10.0, ga=20.0), 256),
    "C4": (dict(data="mnist", dn="24"), 4096),
}


_DATASETS = {}


def synthetic_batch(cfg, batch, seed, device):
    """``[batch, canvas*canvas]`` canvases in [0, 1].  On a GPU they come from the on-device feeder
    (``mog_asr_b200/dataset.py``: placement kernel + the sampler's own write direction); the numpy generator of
    ``synth.py`` serves the CPU tests."""
    dev = torch.device(device)
    if dev.type == "cuda":
        from ..dataset import DeviceMultiObjectDataset, default_sprites
        key = (str(dev), cfg.canvas_size, tuple(cfg.constrains_num), tuple(cfg.constrains_area_minmax))
        if key not in _DATASETS:
            lo, hi = (int(v) for v in cfg.constrains_area_minmax)
            _DATASETS[key] = DeviceMultiObjectDataset(default_sprites(256, 28, seed=0), cfg.canvas_size, tuple(cfg.constrains_num),
                                                      (lo, hi), mode="disjoint", seed=1234, device=dev)
        return torch.clamp(_DATASETS[key].batch(seed, batch)["images"], 0.0, 1.0)
    canv, _ = synth.multi_object_canvases(batch, cfg.canvas_size, 28, tuple(cfg.constrains_num), seed=seed)
    return torch.tensor(np.clip(canv, 0.0, 1.0).reshape(batch, -1), device=device)


def run(name, device, steps=20, warmup=5, process_group=None, always_max_steps=False, graph=False, per_rank_batch=None):
    """Returns dict(images_per_sec, ms_per_step, global_batch, per_rank_batch, mean loop steps).  ``per_rank_batch``
    overrides the config's fixed global batch with ``per_rank_batch * world`` (weak scaling)."""
    flags, gbatch = CONFIGS[name]
    world = dist.get_world_size(process_group) if process_group is not None else 1
    rank = dist.get_rank(process_group) if process_group is not None else 0
    if per_rank_batch is not None:
        gbatch = per_rank_batch * world
    cfg = config_from_flags(always_max_steps=always_max_steps, **flags)
    local = gbatch // world
    tr = Trainer(cfg, device, process_group=process_group, global_batch=gbatch)
    # a few distinct resident batches, cycled (the reference feeds from a shuffle queue)
    batches = [synthetic_batch(cfg, local, 100 * rank + k, device) for k in range(4)]
    step_fn = tr.step
    if graph:
        tr.capture(local)
        step_fn = tr.step_graph
    for k in range(warmup):
        step_fn(batches[k % 4])
    torch.cuda.synchronize(device)
    if process_group is not None:
        dist.barrier(process_group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T = 0
    e0.record()
    for k in range(steps):
        T += step_fn(batches[k % 4])["steps"]
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    if process_group is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=process_group)
        ms = float(t.item())
    return dict(config=name, images_per_sec=gbatch / (ms * 1e-3), ms_per_step=ms, global_batch=gbatch,
                per_rank_batch=local, n_gpus=world, mean_loop_steps=T / steps, grad_floats=tr.num_gradient_floats(),
                mode=("CUDA graph, fixed max_steps" if graph else "eager, fixed max_steps") if always_max_steps
                else "eager, reference loop condition (host-checked any)")
