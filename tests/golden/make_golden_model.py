"""Execute the reference's WHOLE training graph -- ``AIRModel._create_model`` of ``air/air_number_bbox_location.py`` with its
own ``vae.py``, ``concrete.py`` and ``transformer.py`` -- on the torch-based TF shim (``tf_shim_model.py``) in float64, with the
hyper-parameters ``train_air_pr.py:174-212`` passes, seeded weights and injected noise; differentiate ``self.loss`` by autograd.
Writes ``tests/golden/graph_model_<config>.npz``: images, every weight (TF variable names), the noise draws, the loss, the
per-image ELBO / reconstruction loss, inferred counts, scales, shifts, z_pres probabilities, the clipped canvas and the
gradient of the loss w.r.t. every weight.  Run from the repo root in the authoring container."""
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
import air.air_number_bbox_location as ref  # noqa: E402  (imports the reference's vae / concrete / transformer on the shim)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mog_asr_b200 import synth  # noqa: E402  (synthetic canvases only)

CONFIGS = {
    # name: (canvas, counts, gn, gm, gne, gb, gs, ga, minmax, zt) -- the regulariser flags of BASELINE configs 2 and 3; layer
    # widths are shrunk (the wiring does not depend on them) so that weights + gradients stay ~1 MB per file
    "c2": (20, [1, 3], 0.0, 100.0, 10.0, 0.0, 0.0, 0.0, [7.0, 9.0], 0.1),        # -dn 13 -gm 100 -gne 10
    "c3": (24, [3], 0.0, 0.0, 0.0, 1.0, 10.0, 20.0, [5.0, 6.0], 0.1),            # -dn 3 -ds bbox20k -gb 1 -gs 10 -ga 20 (fix_steps = 3)
    "all": (22, [1, 2, 4], 0.7, 3.0, 2.0, 1.5, 0.5, 0.25, [6.0, 8.0], 1.0),      # every regulariser on, temperature 1
}
MAX_STEPS, B = 6, 5
WS, RNN, LAT, REC, GEN, HID = 12, 32, 8, (40, 24), (24, 40), 16


def noise_bank(seed, cs):
    rng = np.random.default_rng(seed)
    bank = {"shift": rng.standard_normal((MAX_STEPS, B, 2)), "scale": rng.standard_normal((MAX_STEPS, B, 1)),
            "vae": rng.standard_normal((MAX_STEPS, B, LAT)), "concrete": np.clip(rng.random((MAX_STEPS, B)), 1e-4, 1 - 1e-4)}
    return bank


for name, (cs, counts, gn, gm, gne, gb, gs, ga, minmax, zt) in CONFIGS.items():
    tf.VARIABLES.clear(); tf._LAYER_COUNTS.clear(); tf._SCOPE.clear()
    tf._RNG = np.random.default_rng(7)
    tf.NOISE["latent"] = LAT
    canv, _ = synth.multi_object_canvases(B, cs, WS, tuple(counts), seed=3)
    images = torch.tensor(np.clip(canv, 0, 1).reshape(B, -1), dtype=torch.float64)
    bank = noise_bank(11, cs)
    tf.NOISE["fn"] = lambda kind, t, dims: torch.tensor(bank[kind][t]).reshape(dims)
    self = types.SimpleNamespace(                                                            # what __init__ sets (:56-113)
        input_images=images, target_num_digits=torch.zeros(B, dtype=torch.int32), batch_size=B, max_steps=MAX_STEPS, max_digits=MAX_STEPS,
        rnn_units=RNN, canvas_size=cs, windows_size=WS, vae_latent_dimensions=LAT, vae_recognition_units=REC,
        vae_generative_units=GEN, scale_prior_mean=-1.0, scale_prior_variance=0.05, fix_scale_distribution=True,
        vae_prior_mean=0.0, vae_prior_variance=1.0, vae_likelihood_std=0.0, scale_hidden_units=HID, shift_hidden_units=HID,
        z_pres_hidden_units=HID, reuse_shift_scale_network=True, z_pres_prior_log_odds=-0.01,
        fix_steps=counts[0] if len(counts) == 1 else None, z_pres_temperature=zt, stopping_threshold=0.9, learning_rate=1e-4,
        gradient_clipping_norm=1.0, num_summary_images=4, cnn=False, cnn_filters=8, train=True, constrains_x_y=None,
        constrains_num_list=counts, constrains_num=torch.tensor(counts), constrains_num_gamma=gn, constrains_bbox_gamma=gb,
        constrains_margin_gamma=gm, constrains_num_element_gamma=gne, constrains_sharesize_gamma=gs, constrains_area_gamma=ga,
        constrains_area_minmax=minmax, log_variables={}, global_step=torch.tensor(0),
        vae_prior_log_variance=torch.log(torch.tensor(1.0, dtype=torch.float64)))
    self._sample_from_mvn = ref.AIRModel._sample_from_mvn
    self._visualize_reconstructions = lambda *a, **k: torch.zeros(B, 1, 1, 3)               # the PNG overlay is not part of the loss
    with tf.variable_scope("air"):
        ref.AIRModel._create_model(self)
    loss = self.loss
    loss.backward()
    out = dict(images=images.numpy(), loss=float(loss.detach()), steps=len(bank["shift"]) and tf.STEP["t"],
               rec_num_digits=self.rec_num_digits.numpy(), rec_scales=self.rec_scales.detach().numpy(), rec_shifts=self.rec_shifts.detach().numpy(),
               z_pres_probs=self.z_pres_probs.detach().numpy(), reconstruction=self.reconstruction.detach().numpy(),
               recon_loss=self.reconstruction_loss.detach().numpy(), margin=float(torch.as_tensor(self.constrains_num_marginal_loss).detach()),
               cfg=np.array(repr(dict(canvas=cs, counts=counts, gn=gn, gm=gm, gne=gne, gb=gb, gs=gs, ga=ga, minmax=minmax, zt=zt, ws=WS, rnn=RNN,
                                      lat=LAT, rec=REC, gen=GEN, hid=HID, max_steps=MAX_STEPS))))
    out.update({"noise_" + k: v for k, v in bank.items()})
    for k, v in tf.VARIABLES.items():
        out["w:" + k] = v.detach().numpy()
        out["g:" + k] = v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape))
    # the reference's TEST model (train=False, :634-635: z_pres rounded) on the same weights and noise
    for v in tf.VARIABLES.values():
        v.grad = None
    tf._LAYER_COUNTS.clear(); tf._SCOPE.clear()
    self.train, self.log_variables = False, {}
    with torch.no_grad(), tf.variable_scope("air"):
        ref.AIRModel._create_model(self)
    out.update(test_steps=tf.STEP["t"], test_rec_num_digits=self.rec_num_digits.numpy(), test_rec_scales=self.rec_scales.numpy(),
               test_rec_shifts=self.rec_shifts.numpy(), test_z_pres_probs=self.z_pres_probs.numpy(), test_reconstruction=self.reconstruction.numpy())
    np.savez_compressed(os.path.join(HERE, f"graph_model_{name}.npz"), **out)
    print(name, "loss", out["loss"], "steps", out["steps"], "digits", out["rec_num_digits"], "variables", len(tf.VARIABLES),
          "| test model: steps", out["test_steps"], "digits", out["test_rec_num_digits"])
    for k, v in tf.VARIABLES.items():
        print("   ", k, tuple(v.shape), "grad" if v.grad is not None else "NO GRAD")
