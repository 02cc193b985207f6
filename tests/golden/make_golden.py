"""Mint the golden fixtures in this directory from the oracle (run from the repo root:
``python tests/golden/make_golden.py``).

The reference has no tests, fixtures or golden vectors and its TF-1.12 runtime is not installable, so
these vectors are produced by the restatement in ``oracle/`` -- *parity unpinned*: they pin the build
against the oracle's stated evaluation order (and the oracle against regressions), not against a live
TensorFlow.  Inputs are seeded; outputs are stored exactly (fp32 bit patterns survive .npz).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import stn_ref_numpy as R  # noqa: E402
from oracle import asr_ref  # noqa: E402
from mog_asr_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
F32 = np.float32


def stn_case(name, U, theta, out_size, seed):
    rng = np.random.default_rng(seed)
    full = R.transformer_full(U, theta, out_size)
    g = rng.normal(size=full["out"].shape).astype(F32)
    dU, dth = R.transformer_backward(U, theta, out_size, g, dtype=np.float64)
    aU, ath = R.backward_term_magnitudes(U, theta, out_size, g)
    wsum = sum(np.abs(full[k].astype(np.float64)) for k in ("wa", "wb", "wc", "wd"))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), U=U.astype(F32), theta=np.asarray(theta, F32).reshape(-1, 6),
        out_size=np.asarray(out_size, np.int32), out=full["out"],
        corners=np.stack([full[k] for k in ("x0", "x1", "y0", "y1")]).astype(np.int32), gout=g,
        dU=dU, dtheta=dth, absdU=aU, absdtheta=ath, wsum=wsum.astype(np.float64),
        footprint=R.footprint_counts(full, U.shape[0]))
    print(name, "out", full["out"].shape, "F mean", R.footprint_counts(full, U.shape[0]).mean())


def adversarial_thetas():
    t = []
    t.append([1, 0, 0, 0, 1, 0])                    # identity: x_s = +-1 hits the first/last pixel exactly
    t.append([0.5, 0, 0.5, 0, 0.5, -0.5])           # corner quadrant
    t.append([1, 0, 1.5, 0, 1, 0])                  # out of range right
    t.append([1, 0, -1.5, 0, 1, 0])                 # out of range left
    t.append([1, 0, 0, 0, 1, 1.7])                  # out of range bottom
    t.append([1, 0, 0, 0, 1, -1.7])                 # out of range top
    t.append([1, 0, 3.0, 0, 1, -3.0])               # out of range on both axes -> exactly 0
    t.append([0.7, 0.3, 0.1, -0.3, 0.7, -0.1])      # rotation + scale (non-separable path)
    t.append([0.0, 1.0, 0.0, 1.0, 0.0, 0.0])        # transpose
    t.append([-1, 0, 0, 0, -1, 0])                  # flip both axes
    t.append([0, 0, 0.25, 0, 0, -0.25])             # degenerate: every pixel samples one point
    t.append([1e-3, 0, 0.2, 0, 1e-3, 0.2])          # tiny scale (extreme magnification)
    t.append([1e3, 0, -200.0, 0, 1e3, 300.0])       # 1/s huge (write direction with s -> 0)
    t.append([2.0 / 49 * 24.5, 0, 0, 0, 1, 0])      # grid lands on/near integer coordinates
    t.append([1.0, 0.0, 2.0 ** -20, 0.0, 1.0, -(2.0 ** -20)])  # sub-ulp shifts around the borders
    t.append([1.3, -0.4, 0.6, 0.9, 0.2, -0.8])      # shear, partly outside
    return np.asarray(t, F32)


def main():
    rng = np.random.default_rng(1234)
    # C1 read: 50x50 -> 28x28
    canv, _ = synth.multi_object_canvases(8, 50, 28, (1, 2, 3), seed=0)
    s, x, y = synth.sxy_prior_like(8, seed=1)
    stn_case("read_50_28", canv[..., None], synth.theta_read(s, x, y), (28, 28), 3)
    # C1 write: 28x28 -> 50x50
    Uw = (1 / (1 + np.exp(-rng.normal(size=(8, 28, 28, 1))))).astype(F32)
    stn_case("write_28_50", Uw, synth.theta_write(s, x, y), (50, 50), 4)
    # adversarial thetas on a non-square multi-channel source
    th = adversarial_thetas()
    Ua = rng.random((th.shape[0], 17, 23, 3), dtype=F32)
    stn_case("adversarial_17x23x3_9x31", Ua, th, (9, 31), 5)
    Ub = rng.random((th.shape[0], 50, 50, 1), dtype=F32)
    stn_case("adversarial_50_28", Ub, th, (28, 28), 6)
    stn_case("adversarial_28_50", rng.random((th.shape[0], 28, 28, 1), dtype=F32), th, (50, 50), 7)
    # degenerate output sizes
    stn_case("out_1x1", rng.random((4, 6, 5, 2), dtype=F32), th[[0, 1, 7, 15]], (1, 1), 8)
    stn_case("out_1x7", rng.random((4, 6, 5, 2), dtype=F32), th[[0, 1, 7, 15]], (1, 7), 9)
    # full-cover regime at 64 -> 28 and 128 -> 64 (small batch)
    s2, x2, y2 = synth.sxy_full_cover(4, seed=2)
    stn_case("fullcover_64_28", rng.random((4, 64, 64, 1), dtype=F32), synth.theta_read(s2, x2, y2), (28, 28), 10)
    stn_case("fullcover_128_64", rng.random((2, 128, 128, 1), dtype=F32), synth.theta_read(s2[:2], x2[:2], y2[:2]), (64, 64), 11)

    # write + composite
    B = 8
    canvas = rng.random((B, 50, 50), dtype=F32)
    z = rng.random(B, dtype=F32)
    stop = np.asarray([0.1, 0.95, 0.5, 0.89999, 0.9, 1.7, 0.0, 0.3], F32)
    thw = synth.theta_write(s, x, y)
    mask = stop < F32(0.9)
    newc = R.write_composite(canvas, Uw[..., 0], thw, z, mask)
    gc = rng.normal(size=(B, 50, 50)).astype(F32)
    dU, dth, dz = R.write_composite_backward(Uw[..., 0], thw, z, mask, gc)
    np.savez_compressed(os.path.join(HERE, "composite_28_50.npz"), canvas=canvas, U=Uw[..., 0], theta=thw, z=z,
                        stop_sum=stop, threshold=F32(0.9), out=newc, gcanvas=gc, dU=dU, dtheta=dth, dz=dz)
    print("composite_28_50 active", mask.sum())

    # ASR regularisers: C2 (-dn 13 -gm 100 -gne 10, MNIST) and C3 (-dn 3 bbox -gb 1 -gs 10 -ga 20, sprites)
    for name, cfg, Bq in (
        ("asr_c2", dict(canvas_size=50, counts=[1, 3], max_steps=6, gamma_margin=100.0, gamma_elem=10.0,
                        area_minmax=(17.0, 23.0)), 16),
        ("asr_c3", dict(canvas_size=64, counts=[3], max_steps=6, gamma_bbox=1.0, gamma_size=10.0, gamma_area=20.0,
                        area_minmax=(12.0, 15.0)), 16),
        ("asr_all", dict(canvas_size=50, counts=[2, 4], max_steps=6, gamma_num=0.5, gamma_margin=3.0, gamma_elem=2.0,
                         gamma_bbox=1.5, gamma_size=0.7, gamma_area=0.3, area_minmax=(11.0, 15.0)), 12),
    ):
        for T in (6, 3):
            lo = rng.normal(0, 2, (Bq, T)).astype(F32)
            sh = np.tanh(rng.normal(0, 1, (Bq, T, 2))).astype(F32)
            sc = (1 / (1 + np.exp(-rng.normal(-1, 0.5, (Bq, T))))).astype(F32)
            gp = rng.random(Bq).astype(F32)
            r = asr_ref.asr_numpy(lo, sh, sc, g_per_image=gp, g_margin=0.7, dtype=np.float64, **cfg)
            np.savez_compressed(os.path.join(HERE, f"{name}_T{T}.npz"), log_odds=lo, shifts=sh, scales=sc,
                                g_per_image=gp, g_margin=np.float64(0.7),
                                cfg_keys=np.asarray(list(cfg.keys())),
                                cfg_vals=np.asarray([str(v) for v in cfg.values()]),
                                **{k: v for k, v in r.items()})
            print(f"{name}_T{T}", "per_image mean", r["per_image"].mean(), "margin", r["margin"])


if __name__ == "__main__":
    main()
