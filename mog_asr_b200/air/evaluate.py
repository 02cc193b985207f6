"""The periodic test pass of the reference trainer (``train_air_pr.py:314-334``): run the TEST model (z_pres rounded,
``air_number_bbox_location.py:634-635``) on a labelled batch and score the inferred boxes with the detection metrics
(``air/evaluation_detection.py``), everything on the device."""
from __future__ import annotations

import torch

from .. import detection


@torch.no_grad()
def evaluate_detection(model, batch, noise=None):
    """``batch``: a dict from ``DeviceMultiObjectDataset.batch`` (``images``, ``pos``, ``size``, ``num``).  Returns the five
    batch means the reference logs -- precision[11], recall[11], gt IoU, detection IoU, global IoU (:326-334) -- as
    float64 tensors, plus the count accuracy (:1073-1077)."""
    out = model(torch.clamp(batch["images"], 0.0, 1.0), noise=noise, train=False)
    cs = model.cfg.canvas_size
    # the loop may have stopped early: only the executed steps exist (inf_num never exceeds them)
    p, r, g, d, m = detection.detection_metrics(batch["pos"], batch["size"], batch["num"], out["rec_shifts"], out["rec_scales"],
                                                out["rec_num_digits"], cs)
    accuracy = (out["rec_num_digits"] == batch["num"]).double().mean()
    return dict(precision=p.mean(0), recall=r.mean(0), gt_max_iou=g.mean(), detected_max_iou=d.mean(), global_iou=m.mean(),
                accuracy=accuracy, steps=out["steps"])
