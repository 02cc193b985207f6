"""Execute the reference's own ``air/concrete.py`` on the numpy TF shim in float64 and store inputs and outputs
(``tests/golden/graph_concrete.npz``).  The uniform draws are injected (``tf_shim.UNIFORM_QUEUE``).  Run from the repo
root in the authoring container: ``python tests/golden/make_golden_concrete.py``."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

sys.modules["tensorflow"] = tf_shim
spec = importlib.util.spec_from_file_location("ref_concrete", "/root/reference/air/concrete.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(3)
n = 512
log_odds = rng.normal(0, 3, n)
prior_lo = np.where(rng.random(n) < 0.3, rng.choice([100.0, -100.0], n), rng.normal(0, 3, n))   # fixed +-100 priors (:604-608) included
u = np.clip(rng.random(n), 1e-6, 1 - 1e-6)
out = {}
for temp in (0.1, 1.0):
    tf_shim.UNIFORM_QUEUE.append(u)
    y = np.asarray(ref.concrete_binary_pre_sigmoid_sample(tf_shim._t(log_odds), temp))                      # concrete.py:20-27
    kl = np.asarray(ref.concrete_binary_kl_mc_sample(tf_shim._t(y), tf_shim._t(prior_lo), temp, tf_shim._t(log_odds), temp))   # :30-64
    tf_shim.UNIFORM_QUEUE.append(u)
    y2, sig = ref.concrete_binary_sample(tf_shim._t(log_odds), temp)                                        # :4-17
    out[f"y_T{temp}"], out[f"kl_T{temp}"], out[f"sig_T{temp}"] = y, kl, np.asarray(sig)
    print("temperature", temp, "y", y[:3], "kl", kl[:3], "finite", np.isfinite(kl).all())
np.savez_compressed(os.path.join(HERE, "graph_concrete.npz"), log_odds=log_odds, prior_lo=prior_lo, u=u, **out)
