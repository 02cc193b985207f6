"""Synthetic-dataset feeder: placement kernel and paste against oracle/synth_ref.py (same
counter-based draws, scalar Python), the reference's placement rules (multi_mnist.py:77-87,:110-221) as invariants."""
import numpy as np
import pytest

from oracle import synth_ref as S

CASES = [  # seed, first, B, canvas, counts, size range, gap, margin, mode, share_size, sprites
    (1, 0, 600, 50, [1, 3], (11, 15), 0, 0, "bbox", False, 64),       # Multi-MNIST "bbox" datasets (train_air_pr.py:83-97)
    (2, 5000, 400, 50, [1, 2, 3], (17, 23), 0, 0, "disjoint", False, 64),
    (3, 17, 300, 64, [3], (12, 15), 2, 1, "bbox", True, 120),         # Multi-dSprites, -ds bbox20k, shared size
    (4, 0, 200, 50, [0, 2], (20, 30), 1, 2, "disjoint", False, 3),    # empty canvases, crowded ones (restarts)
]
MODE = {"bbox": 0, "disjoint": 1}


def oracle_place(c):
    seed, first, B, cs, counts, (smin, smax), gap, margin, mode, share, nsp = c
    return S.place(seed, first, B, cs, counts, smin, smax, gap, margin, MODE[mode], share, nsp)


@pytest.mark.parametrize("c", CASES)
def test_oracle_placement_obeys_the_reference_rules(c):
    seed, first, B, cs, counts, (smin, smax), gap, margin, mode, share, nsp = c
    num, pos, size, sprite = oracle_place(c)
    assert set(np.unique(num)) <= set(range(max(counts) + 1))
    for b in range(B):
        boxes = [(pos[b, g, 0], pos[b, g, 1], size[b, g, 0]) for g in range(num[b])]
        for (x, y, w) in boxes:
            assert smin <= w <= smax and margin <= x <= cs - w - margin and margin <= y <= cs - w - margin   # :171-172
        if share and boxes:
            assert len({w for _, _, w in boxes}) == 1                                                       # :136-142
        for i in range(len(boxes)):
            for k in range(i):
                (x, y, w), (qx, qy, qw) = boxes[i], boxes[k]
                xhit = x - gap <= qx + qw - 1 and qx <= x + w + gap - 1
                yhit = y - gap <= qy + qw - 1 and qy <= y + w + gap - 1
                assert not xhit if mode == "bbox" else not (xhit and yhit)                                   # :77-87
        assert np.all(sprite[b, :num[b]] < nsp) and np.all(pos[b, num[b]:] == 0)
    # a canvas is a function of (seed, canvas index) only
    n2, p2, s2, sp2 = S.place(seed, first + 7, 20, cs, counts, smin, smax, gap, margin, MODE[mode], share, nsp)
    assert np.array_equal(n2, num[7:27]) and np.array_equal(p2, pos[7:27]) and np.array_equal(sp2, sprite[7:27])


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES)
def test_cuda_placement_and_paste_match_oracle(c, cuda_device):
    import torch
    from mog_asr_b200.dataset import DeviceMultiObjectDataset, default_sprites
    seed, first, B, cs, counts, srange, gap, margin, mode, share, nsp = c
    sprites = default_sprites(nsp, 28, seed=9)
    ds = DeviceMultiObjectDataset(sprites, cs, counts, srange, gap, margin, mode, share, seed, device=cuda_device)
    num, pos, size, sprite = ds.place(first, B)
    rn, rp, rs, rsp = oracle_place(c)
    for got, want in ((num, rn), (pos, rp), (size, rs), (sprite, rsp)):
        assert np.array_equal(got.cpu().numpy(), want)
    m = min(B, 96)
    img = ds.paste(num[:m], pos[:m], size[:m], sprite[:m]).cpu().numpy()
    ref = S.paste(sprites, cs, rn[:m], rp[:m], rs[:m], rsp[:m])
    assert np.array_equal(img, ref)                      # sampler forward is bit-exact; objects never share a pixel
    # every object landed inside its box, and the box is not empty
    im3 = img.reshape(m, cs, cs)
    for b in range(m):
        mask = np.zeros((cs, cs), bool)
        for g in range(rn[b]):
            x, y, w = rp[b, g, 0], rp[b, g, 1], rs[b, g, 0]
            mask[y:y + w, x:x + w] = True
            assert im3[b, y:y + w, x:x + w].sum() > 0
        assert np.all(im3[b][~mask] == 0)


@pytest.mark.gpu
def test_batches_are_a_function_of_the_canvas_index_and_feed_the_detection_metrics(cuda_device):
    import torch
    from mog_asr_b200 import detection
    from mog_asr_b200.dataset import DeviceMultiObjectDataset, default_sprites
    ds = DeviceMultiObjectDataset(default_sprites(32, 28, seed=1), 50, (1, 2, 3), (11, 15), mode="disjoint", seed=11, device=cuda_device)
    whole = ds.batch(0, 64)
    halves = [ds.batch(0, 32, rank=r, world=2) for r in range(2)]       # two ranks of a data-parallel job
    assert torch.equal(torch.cat([h["images"] for h in halves]), whole["images"])
    again = list(ds.stream(64, 2))[0]
    assert torch.equal(again["images"], whole["images"])
    # ground truth in the detection layout: inferring exactly the ground-truth boxes scores 1 everywhere
    cs = 50
    n, T = 64, whole["pos"].shape[1]
    centre = (whole["pos"].double() + whole["size"].double() / 2) / (cs / 2) - 1
    scale = whole["size"][:, :, :1].double() / cs
    p, r, g, d, m = detection.detection_metrics(whole["pos"], whole["size"], whole["num"], centre, scale, whole["num"], cs)
    assert float(p[:, 0].min()) == 1.0 and float(r[:, 0].min()) == 1.0 and float(m.min()) > 0.9


def test_overlap_rule_mode0_is_the_reference_function():
    """Parity PINNED for the placement rule: the golden file holds the outputs of the reference's own
    ``bounding_boxes_overlap`` (tests/golden/make_golden_synth_rule.py)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_overlap_rule.npz"))
    got = np.array([S.boxes_clash(0, *[int(v) for v in a], *[int(v) for v in b], int(g)) for a, b, g in zip(z["new"], z["old"], z["gap"])])
    assert np.array_equal(got, z["overlap"])
    assert 0.2 < got.mean() < 0.9
    multi = np.array([any(S.boxes_clash(0, *[int(v) for v in a], *[int(v) for v in q], 1) for q in o)
                      for a, o in zip(z["multi_new"], z["multi_old"])])
    assert np.array_equal(multi, z["multi_overlap"])
    # it is NOT a box-intersection test: boxes in disjoint rows but overlapping columns clash
    assert S.boxes_clash(0, 10, 0, 5, 5, 10, 40, 5, 5, 0) and not S.boxes_clash(1, 10, 0, 5, 5, 10, 40, 5, 5, 0)
