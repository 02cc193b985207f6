"""Shared test helpers: golden loading and the tolerances stated in DESIGN.md."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STN_CASES = ["read_50_28", "write_28_50", "adversarial_17x23x3_9x31", "adversarial_50_28", "adversarial_28_50",
             "out_1x1", "out_1x7", "fullcover_64_28", "fullcover_128_64"]
ASR_CASES = [f"{n}_T{t}" for n in ("asr_c2", "asr_c3", "asr_all") for t in (6, 3)]

EPS32 = 2.0 ** -23
# Forward values: the kernels use the oracle's operation order, so equality is expected; the stated
# tolerance is the fallback bound  |out - ref| <= 4 eps * sum|w_k| * max|U|.
FWD_EPS_FACTOR = 4.0
# Gradients: fp32 products and an fp32 sum in a different order than the fp64 oracle; bound relative to
# the sum of |terms| entering each entry.
GRAD_RTOL = 2e-5


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def asr_cfg(d):
    cfg = {}
    for k, v in zip(d["cfg_keys"], d["cfg_vals"]):
        cfg[str(k)] = eval(str(v), {"__builtins__": {}})  # literals written by make_golden.py
    return cfg


def same_bits_or_nan(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def grad_excess(got, ref, absterms, rtol=GRAD_RTOL):
    """max over entries of |got-ref| / (rtol*absterms + tiny); <= 1 passes.  Non-finite refs are skipped."""
    got, ref, absterms = (np.asarray(a, np.float64) for a in (got, ref, absterms))
    ok = np.isfinite(ref) & np.isfinite(absterms)
    if not ok.any():
        return 0.0
    scale = rtol * absterms[ok] + 1e-30 + rtol * 1e-3 * np.max(absterms[ok])
    return float(np.max(np.abs(got[ok] - ref[ok]) / scale))
