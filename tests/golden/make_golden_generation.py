"""Execute the reference's GENERATION graph -- ``AIRModel._create_generation`` (``air/air_number_bbox_location.py:1124-1361``) with
its own ``vae_generation`` (``air/vae.py:51-86``), ``concrete.py`` and ``transformer.py`` -- on the torch-based TF shim in float64.
The weights are those of the training graph built first under the same variable scope (the reference shares them by scope name,
``train_air_pr.py:201``); noise is injected: shift / scale / latent normals, the Bernoulli uniforms of ``sample_from_mean=True``
(``vae.py:83-84``) and the Concrete uniforms.  Writes ``tests/golden/graph_generation_<config>.npz``: every weight, the noise, the
generated canvases, object counts, per-step write thetas and the loop's trip count.  Run from the repo root in the authoring
container (``/root/reference`` is read at run time, nothing is copied)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim_model as tf  # noqa: E402

tf.DEFAULT["dtype"] = torch.float64
tf.install()
sys.path.insert(0, "/root/reference")
import air.air_number_bbox_location as ref  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mog_asr_b200 import synth  # noqa: E402

CONFIGS = {"learned_prior": (20, [1, 3], [7.0, 9.0], 0.1), "fix_steps": (24, [3], [5.0, 6.0], 0.1)}
MAX_STEPS, B = 6, 7
WS, RNN, LAT, REC, GEN, HID = 12, 32, 8, (40, 24), (24, 40), 16


def uniform_dispatch(shape, minval=0, maxval=1, **k):
    dims = tf._ints(shape)
    kind = "bernoulli" if len(dims) == 2 else "concrete"
    return tf.NOISE["fn"](kind, tf.STEP["t"], dims).to(tf.DEFAULT["dtype"])


tf.random_uniform = uniform_dispatch
tf.sign = torch.sign              # vae.py:84 (the only op of the generation graph the training graph does not use)
import tensorflow as _tfmod  # noqa: E402  (the shim module the reference files imported)
_tfmod.random_uniform = uniform_dispatch
_tfmod.sign = torch.sign

for name, (cs, counts, minmax, zt) in CONFIGS.items():
    tf.VARIABLES.clear(); tf._LAYER_COUNTS.clear(); tf._SCOPE.clear()
    tf._RNG = np.random.default_rng(17)
    tf.NOISE["latent"] = LAT
    rng = np.random.default_rng(23)
    bank = {"shift": rng.standard_normal((MAX_STEPS, B, 2)), "scale": rng.standard_normal((MAX_STEPS, B, 1)),
            "vae": rng.standard_normal((MAX_STEPS, B, LAT)), "concrete": np.clip(rng.random((MAX_STEPS, B)), 1e-4, 1 - 1e-4),
            "bernoulli": rng.random((MAX_STEPS, B, WS * WS))}
    tf.NOISE["fn"] = lambda kind, t, dims: torch.tensor(bank[kind][t]).reshape(dims)
    canv, _ = synth.multi_object_canvases(B, cs, WS, tuple(counts), seed=3)
    images = torch.tensor(np.clip(canv, 0, 1).reshape(B, -1), dtype=torch.float64)
    self = types.SimpleNamespace(
        input_images=images, target_num_digits=torch.zeros(B, dtype=torch.int32), batch_size=B, generation_batch_size=B, max_steps=MAX_STEPS,
        max_digits=MAX_STEPS, rnn_units=RNN, canvas_size=cs, windows_size=WS, vae_latent_dimensions=LAT, vae_recognition_units=REC,
        vae_generative_units=GEN, scale_prior_mean=-1.0, scale_prior_variance=0.05, fix_scale_distribution=True,
        vae_prior_mean=0.0, vae_prior_variance=1.0, vae_likelihood_std=0.0, scale_hidden_units=HID, shift_hidden_units=HID,
        z_pres_hidden_units=HID, reuse_shift_scale_network=True, z_pres_prior_log_odds=-0.01,
        fix_steps=counts[0] if len(counts) == 1 else None, z_pres_temperature=zt, stopping_threshold=0.9, learning_rate=1e-4,
        gradient_clipping_norm=1.0, num_summary_images=4, cnn=False, cnn_filters=8, train=True, constrains_x_y=None,
        constrains_num_list=counts, constrains_num=torch.tensor(counts), constrains_num_gamma=0.0, constrains_bbox_gamma=0.0,
        constrains_margin_gamma=0.0, constrains_num_element_gamma=0.0, constrains_sharesize_gamma=0.0, constrains_area_gamma=0.0,
        constrains_area_minmax=minmax, log_variables={}, global_step=torch.tensor(0),
        vae_prior_log_variance=torch.log(torch.tensor(1.0, dtype=torch.float64)))
    self._sample_from_mvn = ref.AIRModel._sample_from_mvn
    self._visualize_reconstructions = lambda *a, **k: torch.zeros(B, 1, 1, 3)
    captured = {}

    def _viz(reconstruction, st_back, steps, zoom):     # the PNG overlay is not part of the samples; keep what it is given
        captured.update(st_back=st_back.detach().numpy().copy(), steps=steps.numpy().copy())
        return torch.zeros(B, 1, 1, 3)
    self._visualize_generations = _viz
    with torch.no_grad(), tf.variable_scope("air"):
        ref.AIRModel._create_model(self)                 # creates the (seeded) variables the generation graph shares
    nvars = len(tf.VARIABLES)
    tf._LAYER_COUNTS.clear(); tf._SCOPE.clear()
    with torch.no_grad(), tf.variable_scope("air"):
        samples, _ = ref.AIRModel._create_generation(self)
    assert len(tf.VARIABLES) == nvars, "the generation graph must reuse the training graph's variables"
    out = dict(samples=samples.numpy(), steps=tf.STEP["t"], num=captured["steps"], thetas=captured["st_back"],
               cfg=np.array(repr(dict(canvas=cs, counts=counts, minmax=minmax, zt=zt, ws=WS, rnn=RNN, lat=LAT, rec=REC, gen=GEN, hid=HID,
                                      max_steps=MAX_STEPS))))
    out.update({"noise_" + k: v for k, v in bank.items()})
    for k, v in tf.VARIABLES.items():
        out["w:" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, f"graph_generation_{name}.npz"), **out)
    print(name, "steps", out["steps"], "objects", out["num"], "canvas sum", float(samples.sum()), "variables", nvars)
