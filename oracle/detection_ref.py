"""ORACLE (test infrastructure, NOT product code) -- detection metrics of the reference, restated.

Follows ``/root/reference/air/evaluation_detection.py``: ``IoU_evaluation`` ``:5-26`` and ``evaluation`` ``:29-98``
(IoU matrix between ground-truth and inferred boxes, precision/recall at the 11 thresholds 0.5 .. 1.0, mean max-IoU
both ways, Hungarian-matched mean IoU).

**Parity PINNED**: unlike the sampler, this reference file needs only numpy and scipy and is importable in the
authoring container, so ``tests/golden/make_golden_detection.py`` runs the reference function itself and commits its
inputs and outputs (``tests/golden/detection_*.npz``); this restatement is checked against those vectors
(``tests/test_detection.py``).  All arithmetic is float64 (the reference's numpy 1.16 promotes its float32 inputs to
float64 in these scalar expressions; feeding float64 makes every numpy version agree).

The per-instance arrays are returned as well (the reference only returns their means).
"""
import numpy as np
from scipy.optimize import linear_sum_assignment


def iou(boxA, boxB):
    """evaluation_detection.py:5-26 (inclusive pixel boxes: the ``+ 1`` terms)"""
    xA, yA = max(boxA[0], boxB[0]), max(boxA[1], boxB[1])
    xB, yB = min(boxA[2], boxB[2]), min(boxA[3], boxB[3])
    inter = max(0, xB - xA + 1) * max(0, yB - yA + 1)
    areaA = (boxA[2] - boxA[0] + 1) * (boxA[3] - boxA[1] + 1)
    areaB = (boxB[2] - boxB[0] + 1) * (boxB[3] - boxB[1] + 1)
    return inter / float(areaA + areaB - inter)


def evaluation_per_instance(gt_position_xy, gt_scale_xy, inf_shifts, inf_scales, inf_num, csize=50):
    n = len(gt_position_xy)
    inf_shifts = np.asarray(inf_shifts, np.float64)
    inf_scales = np.asarray(inf_scales, np.float64)
    precision, recall = np.zeros([n, 11]), np.zeros([n, 11])
    gt_max, det_max, glob = np.zeros(n), np.zeros(n), np.zeros(n)
    csize_2 = csize / 2                                                             # :55
    for k in range(n):
        pos, size = gt_position_xy[k], gt_scale_xy[k]
        num_gt, num_inf = len(pos) // 2, int(inf_num[k])                            # :40
        M = np.zeros([num_gt, num_inf])
        for a in range(num_gt):
            gt_box = [pos[2 * a], pos[2 * a + 1], pos[2 * a] + size[2 * a], pos[2 * a + 1] + size[2 * a + 1]]   # :45-48
            for b in range(num_inf):
                cx, cy, s = inf_shifts[k, b, 0], inf_shifts[k, b, 1], inf_scales[k, b, 0]        # :50-53
                box = [(cx + 1) * csize_2 - s * csize_2, (cy + 1) * csize_2 - s * csize_2,       # :57-61
                       (cx + 1) * csize_2 + s * csize_2, (cy + 1) * csize_2 + s * csize_2]
                M[a, b] = iou(gt_box, box)                                                       # :63
        if num_gt == 0 and num_inf == 0:                                            # :66-71
            precision[k, :] = 1; recall[k, :] = 1; gt_max[k] = 1; det_max[k] = 1; glob[k] = 1
        elif num_gt == 0:                                                           # :72-73
            recall[k, :] = 1
        elif num_inf == 0:                                                          # :74-75
            pass
        else:
            for i in range(11):                                                     # :77-85
                hit = (M > i * 0.05 + 0.5).astype(np.int32)
                tp = np.sum(np.max(hit, 0))
                precision[k, i] = tp / num_inf
                recall[k, i] = tp / num_gt
            gt_max[k] = np.mean(np.max(M, 1))                                       # :86
            det_max[k] = np.mean(np.max(M, 0))                                      # :87
            r, c = linear_sum_assignment(-1 * M)                                    # :89
            glob[k] = np.sum(M[r, c]) / max([num_inf, num_gt])                      # :90-91
    return precision, recall, gt_max, det_max, glob


def evaluation(gt_position_xy, gt_scale_xy, inf_shifts, inf_scales, inf_num, csize=50):
    """same return value as the reference (``:97-98``)"""
    p, r, g, d, m = evaluation_per_instance(gt_position_xy, gt_scale_xy, inf_shifts, inf_scales, inf_num, csize)
    return np.mean(p, 0), np.mean(r, 0), np.mean(g), np.mean(d), np.mean(m)
