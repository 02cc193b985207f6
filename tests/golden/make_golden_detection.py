"""Mint the detection golden vectors by running THE REFERENCE ITSELF (run from the repo root in the authoring
container: ``python tests/golden/make_golden_detection.py``).  ``/root/reference/air/evaluation_detection.py`` needs
only numpy + scipy, so it is imported from where it lies (never copied) and its outputs are committed together with
the seeded inputs.  This is the one part of the path whose parity is pinned by the reference's own code."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/air/evaluation_detection.py"
spec = importlib.util.spec_from_file_location("ref_evaluation_detection", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

MAXG = 4


def make_case(seed, n, csize, max_gt, max_inf):
    rng = np.random.default_rng(seed)
    gt_pos, gt_size = [], []
    gt_num = rng.integers(0, max_gt + 1, n)
    inf_num = rng.integers(0, max_inf + 1, n)
    inf_shifts = np.tanh(rng.normal(0, 0.6, (n, max_inf, 2)))          # float64 on purpose (see oracle/detection_ref.py)
    inf_scales = 1 / (1 + np.exp(-rng.normal(-1, 0.5, (n, max_inf, 1))))
    for k in range(n):
        pos, size = [], []
        for a in range(gt_num[k]):
            w, h = int(rng.integers(8, 24)), int(rng.integers(8, 24))
            x, y = int(rng.integers(0, csize - w + 1)), int(rng.integers(0, csize - h + 1))
            pos += [x, y]; size += [w, h]
            # make some inferred boxes land on the ground truth so that high-IoU thresholds are exercised
            if a < inf_num[k] and rng.random() < 0.7:
                s = max(w, h) / csize * rng.uniform(0.9, 1.1)
                inf_scales[k, a, 0] = s
                inf_shifts[k, a, 0] = (x + w / 2) / (csize / 2) - 1 + rng.normal(0, 0.02)
                inf_shifts[k, a, 1] = (y + h / 2) / (csize / 2) - 1 + rng.normal(0, 0.02)
        gt_pos.append(pos); gt_size.append(size)
    return gt_pos, gt_size, gt_num, inf_shifts, inf_scales, inf_num


def main():
    for name, seed, n, csize, max_gt, max_inf in (("detection_mnist", 1, 400, 50, 3, 6), ("detection_sprites", 2, 300, 64, 3, 6),
                                                   ("detection_edge", 3, 64, 50, 2, 3)):
        gt_pos, gt_size, gt_num, sh, sc, inf_num = make_case(seed, n, csize, max_gt, max_inf)
        if name == "detection_edge":       # every branch of :66-75 is present
            gt_pos[0], gt_size[0], gt_num[0], inf_num[0] = [], [], 0, 0
            gt_pos[1], gt_size[1], gt_num[1], inf_num[1] = [], [], 0, 2
            inf_num[2] = 0
            if gt_num[2] == 0:
                gt_pos[2], gt_size[2], gt_num[2] = [3, 4], [10, 12], 1
        prec, rec, g, d, m = ref.evaluation(gt_pos, gt_size, sh, sc, inf_num, csize=csize)
        P = np.full((n, MAXG, 2), -1, np.int64); S = np.full((n, MAXG, 2), -1, np.int64)
        for k in range(n):
            for a in range(gt_num[k]):
                P[k, a] = gt_pos[k][2 * a:2 * a + 2]; S[k, a] = gt_size[k][2 * a:2 * a + 2]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), gt_pos=P, gt_size=S, gt_num=gt_num.astype(np.int64), inf_shifts=sh,
                            inf_scales=sc, inf_num=inf_num.astype(np.int64), csize=np.int64(csize), precision=prec, recall=rec,
                            gt_max_iou=np.float64(g), detected_max_iou=np.float64(d), global_iou_mean=np.float64(m))
        print(name, "precision", np.round(prec, 4), "recall@0.5", round(rec[0], 4), g, d, m)


if __name__ == "__main__":
    main()
