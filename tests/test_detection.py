"""Detection metrics.  Parity here is PINNED: the golden vectors were produced by running the
reference's own ``evaluation_detection.evaluation`` (``tests/golden/make_golden_detection.py``)."""
import os

import numpy as np
import pytest

from oracle import detection_ref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = ["detection_mnist", "detection_sprites", "detection_edge"]


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    n = len(z["gt_num"])
    pos = [z["gt_pos"][k, :z["gt_num"][k]].reshape(-1).tolist() for k in range(n)]
    size = [z["gt_size"][k, :z["gt_num"][k]].reshape(-1).tolist() for k in range(n)]
    return z, pos, size


def random_case(seed, n, csize, max_gt, max_inf):
    rng = np.random.default_rng(seed)
    gt_num, inf_num = rng.integers(0, max_gt + 1, n), rng.integers(0, max_inf + 1, n)
    pos, size = [], []
    for k in range(n):
        wh = rng.integers(4, csize // 2, (gt_num[k], 2))
        xy = (rng.random((gt_num[k], 2)) * (csize - wh)).astype(np.int64)
        pos.append(xy.reshape(-1).tolist()); size.append(wh.reshape(-1).tolist())
    shifts = np.tanh(rng.normal(0, 0.5, (n, max_inf, 2))).astype(np.float32)
    scales = (1 / (1 + np.exp(-rng.normal(-0.5, 0.7, (n, max_inf, 1))))).astype(np.float32)
    return pos, size, shifts, scales, inf_num


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_reference_golden(name):
    z, pos, size = load(name)
    got = detection_ref.evaluation(pos, size, z["inf_shifts"], z["inf_scales"], z["inf_num"], csize=int(z["csize"]))
    for g, key in zip(got, ("precision", "recall", "gt_max_iou", "detected_max_iou", "global_iou_mean")):
        assert np.array_equal(np.asarray(g), z[key]), key      # same float64 operations in the same order: bit-exact


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_matches_reference_golden(name, cuda_device):
    from mog_asr_b200 import detection
    z, pos, size = load(name)
    got = detection.evaluation(pos, size, z["inf_shifts"], z["inf_scales"], z["inf_num"], csize=int(z["csize"]), device=cuda_device)
    for g, key in zip(got[:4], ("precision", "recall", "gt_max_iou", "detected_max_iou")):
        assert np.array_equal(np.asarray(g), z[key]), key
    assert abs(got[4] - z["global_iou_mean"]) <= 1e-14        # matching found by DP, not scipy: equal optimum, ulp-level sum


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,csize,max_gt,max_inf", [(0, 512, 50, 3, 3), (1, 256, 64, 8, 8), (2, 300, 50, 8, 2), (3, 300, 50, 2, 8),
                                                           (4, 1, 50, 1, 1), (5, 64, 256, 6, 6)])
def test_cuda_per_instance_vs_oracle(seed, n, csize, max_gt, max_inf, cuda_device):
    import torch
    from mog_asr_b200 import detection
    pos, size, shifts, scales, inf_num = random_case(seed, n, csize, max_gt, max_inf)
    want = detection_ref.evaluation_per_instance(pos, size, shifts, scales, inf_num, csize=csize)
    P, S, num = detection.pack_ground_truth(pos, size)
    dev = torch.device(cuda_device)
    got = detection.detection_metrics(torch.tensor(P).to(dev), torch.tensor(S).to(dev), torch.tensor(num).to(dev), torch.tensor(shifts).to(dev),
                                      torch.tensor(scales).to(dev), torch.tensor(inf_num).to(dev), csize)
    got = [g.cpu().numpy() for g in got]
    for i in range(4):
        assert np.array_equal(got[i], want[i]), i
    np.testing.assert_allclose(got[4], want[4], rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_cuda_rejects_too_many_boxes(cuda_device):
    import torch
    from mog_asr_b200 import detection
    dev = torch.device(cuda_device)
    z = lambda *s: torch.zeros(*s, device=dev)
    with pytest.raises(ValueError):
        detection.detection_metrics(z(2, 9, 2), z(2, 9, 2), z(2), z(2, 3, 2), z(2, 3), z(2), 50)


@pytest.mark.gpu
def test_non_integral_ground_truth_is_refused(cuda_device):
    """float ground-truth boxes with fractional coordinates would be truncated by the int32 kernel interface: refused."""
    import torch
    from mog_asr_b200 import detection
    pos = torch.tensor([[[3.5, 4.0]]], device=cuda_device)
    size = torch.tensor([[[10.0, 10.0]]], device=cuda_device)
    num = torch.tensor([1], device=cuda_device)
    sh = torch.zeros((1, 1, 2), device=cuda_device, dtype=torch.float64)
    sc = torch.full((1, 1), 0.3, device=cuda_device, dtype=torch.float64)
    with pytest.raises(ValueError):
        detection.detection_metrics(pos, size, num, sh, sc, num, 50)
    detection.detection_metrics(torch.floor(pos), size, num, sh, sc, num, 50)    # integral floats are fine
