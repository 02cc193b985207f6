"""CPU tests of the oracle itself: the numpy restatement against the committed golden vectors, against
the independently written C restatement (bit for bit), and its closed-form backward against torch
autograd of the restated graph.  (The reference ships no golden vectors: parity unpinned.)"""
import numpy as np
import pytest

from oracle import stn_ref_numpy as R
from oracle import stn_ref_c as RC
from oracle import stn_ref_torch as RT
from oracle import asr_ref
from tests import helpers as H


@pytest.mark.parametrize("name", H.STN_CASES)
def test_numpy_oracle_matches_golden(name):
    d = H.load(name)
    full = R.transformer_full(d["U"], d["theta"], d["out_size"])
    assert H.same_bits_or_nan(full["out"], d["out"])
    for i, k in enumerate(("x0", "x1", "y0", "y1")):
        assert np.array_equal(full[k], d["corners"][i])


@pytest.mark.parametrize("name", H.STN_CASES)
def test_c_restatement_bit_exact_with_numpy(name):
    d = H.load(name)
    out, corners = RC.forward(d["U"], d["theta"], d["out_size"], want_corners=True)
    assert np.array_equal(corners, d["corners"])
    assert H.same_bits_or_nan(out, d["out"])
    dU, dth = RC.backward(d["U"], d["theta"], d["out_size"], d["gout"])
    assert H.grad_excess(dU, d["dU"], d["absdU"]) <= 1.0
    assert H.grad_excess(dth, d["dtheta"], d["absdtheta"]) <= 1.0


def test_linspace_is_tf_recurrence_not_numpy():
    # -1 + step*i in fp32; differs from np.linspace for some i at n = 256 (SURVEY 7 "hard parts")
    a = R.tf_linspace(-1.0, 1.0, 256)
    b = np.linspace(-1, 1, 256, dtype=np.float32)
    assert a[0] == np.float32(-1.0) and a.dtype == np.float32
    assert np.any(a != b)
    assert np.array_equal(R.tf_linspace(-1, 1, 1), np.asarray([-1.0], np.float32))
    step = np.float32(2.0) / np.float32(255)
    assert a[37] == np.float32(np.float32(-1.0) + np.float32(step * np.float32(37)))


def test_scale_constant_is_fp32_of_w_minus_1_001():
    x_s = np.asarray([[1.0]], np.float32)
    x, _, x0, x1, _, _ = R.pixel_coords_and_corners(x_s, x_s, 50, 50)
    assert x[0, 0] == np.float32(np.float32(2.0) * np.float32(np.float32(50) - np.float32(1.001))) / np.float32(2)
    assert (x0[0, 0], x1[0, 0]) == (48, 49)    # x_s = +1 maps below Ws-1: no clipping needed


def test_row_out_of_range_is_exact_zero_and_x_only_is_residue():
    """The shortcut the CUDA forward takes: y out of range => exactly +0 (add_n pairs a,b then c,d)."""
    rng = np.random.default_rng(5)
    U = rng.random((3, 28, 28, 1), dtype=np.float32) + 0.5
    th = np.asarray([[1, 0, 0, 0, 1, 2.5], [1, 0, 0, 0, 1, -2.5], [1, 0, 2.5, 0, 1, 0]], np.float32)
    full = R.transformer_full(U, th, (50, 50))
    assert np.all(full["y0"][:2] == full["y1"][:2])
    assert np.all(full["out"][:2] == 0.0) and not np.signbit(full["out"][:2]).any()
    assert np.all(full["x0"][2] == full["x1"][2])
    assert np.max(np.abs(full["out"][2])) < 1e-4          # cancellation residue, not exactly zero
    # and both axes out of range is exactly zero as well
    th2 = np.asarray([[1, 0, 3, 0, 1, 3]], np.float32)
    assert np.all(R.transformer(U[:1], th2, (50, 50)) == 0.0)


def test_identity_theta_reproduces_interior():
    rng = np.random.default_rng(6)
    U = rng.random((2, 9, 9, 2), dtype=np.float32)
    out = R.transformer(U, np.tile(np.asarray([[1, 0, 0, 0, 1, 0]], np.float32), (2, 1)), (9, 9))
    # (W-1.001)/2 scaling: sample j sits at j*(8-0.001)/8 -- within 1e-3 of the pixel grid
    assert np.max(np.abs(out - U)) < 2e-3


def test_nan_theta_propagates_without_indexing_out_of_bounds():
    U = np.ones((1, 5, 5, 1), np.float32)
    th = np.asarray([[np.nan, 0, 0, 0, 1, 0]], np.float32)
    with np.errstate(invalid="ignore"):
        full = R.transformer_full(U, th, (3, 3))
    assert np.all(full["x0"] == 0) and np.all(full["x1"] == 0)
    assert np.isnan(full["out"]).all()


@pytest.mark.parametrize("shape", [(50, 50, 28, 28, 1, 6), (28, 28, 50, 50, 1, 6), (11, 13, 7, 5, 3, 4)])
def test_closed_form_backward_equals_autograd_fp32(shape):
    """Same fp32 coordinates in both => only the summation order differs."""
    H_, W_, Ho, Wo, C, B = shape
    rng = np.random.default_rng(7)
    U = rng.random((B, H_, W_, C), dtype=np.float32)
    s = 1 / (1 + np.exp(-rng.normal(-0.5, 0.5, B)))
    th = R.theta_read(s, np.tanh(rng.normal(0, .7, B)), np.tanh(rng.normal(0, .7, B))).reshape(B, 6)
    th = (th + rng.normal(0, 0.05, (B, 6))).astype(np.float32)   # general affine
    g = rng.normal(size=(B, Ho, Wo, C)).astype(np.float32)
    import torch
    dUt, dtt = RT.gradients(U, th, (Ho, Wo), g, dtype=torch.float32)
    dU, dth = R.transformer_backward(U, th, (Ho, Wo), g, dtype=np.float64)
    aU, ath = R.backward_term_magnitudes(U, th, (Ho, Wo), g)
    assert H.grad_excess(dUt, dU, aU) <= 1.0
    assert H.grad_excess(dtt, dth, ath) <= 1.0


def test_batch_transformer_repeats_images():
    rng = np.random.default_rng(8)
    U = rng.random((3, 12, 12, 1), dtype=np.float32)
    th = rng.normal(0, 0.5, (3, 4, 6)).astype(np.float32)
    out = R.batch_transformer(U, th, (5, 5))
    assert out.shape == (12, 5, 5, 1)
    assert np.array_equal(out[5], R.transformer(U[1:2], th[1, 1:2], (5, 5))[0])


def test_composite_golden_and_mask_semantics():
    d = H.load("composite_28_50")
    mask = d["stop_sum"] < d["threshold"]
    out = R.write_composite(d["canvas"], d["U"], d["theta"], d["z"], mask)
    assert H.same_bits_or_nan(out, d["out"])
    assert np.array_equal(out[~mask], d["canvas"][~mask])        # inactive images: canvas unchanged
    dU, dth, dz = R.write_composite_backward(d["U"], d["theta"], d["z"], mask, d["gcanvas"])
    assert np.all(dU[~mask] == 0) and np.all(dth[~mask] == 0) and np.all(dz[~mask] == 0)
    np.testing.assert_allclose(dz, d["dz"], rtol=1e-12)


@pytest.mark.parametrize("name", H.ASR_CASES)
def test_asr_oracle_matches_golden(name):
    d = H.load(name)
    r = asr_ref.asr_numpy(d["log_odds"], d["shifts"], d["scales"], g_per_image=d["g_per_image"],
                          g_margin=float(d["g_margin"]), dtype=np.float64, **H.asr_cfg(d))
    for k in ("per_image", "margin", "area", "out", "size", "overlap", "num_min", "pr_num", "d_log_odds", "d_shifts", "d_scales"):
        np.testing.assert_allclose(r[k], d[k], rtol=1e-12, atol=1e-12, err_msg=k)


def test_asr_gating_and_tie_rules():
    lo = np.zeros((2, 3), np.float32)
    sh = np.zeros((2, 3, 2), np.float32)
    sc = np.full((2, 3), 0.3, np.float32)
    # -gne alone does nothing: both count penalties are gated by gamma_margin (:973)
    r = asr_ref.asr_numpy(lo, sh, sc, canvas_size=50, counts=[1, 3], max_steps=6, gamma_elem=10.0)
    assert np.all(r["num_min"] == 0) and r["margin"] == 0
    # all boxes coincide: |px_t - px_u| - 3 < 0 -> no size loss; overlap is full for every ordered pair
    r = asr_ref.asr_numpy(lo, sh, sc, canvas_size=50, counts=[1], max_steps=6, gamma_bbox=1.0, gamma_size=1.0)
    assert np.all(r["size"] == 0)
    np.testing.assert_allclose(r["overlap"], 6 * 15.0, rtol=1e-6)
    # x_diff == y_diff == 0 -> tf.maximum routes to x_diff and abs'(0) = 0: no shift gradient from overlap
    assert np.all(r["d_shifts"] == 0)


GRAPH_CASES = ["read_50_28", "write_28_50", "adversarial_17x23x3_9x31", "adversarial_50_28", "adversarial_28_50", "out_1x1", "out_1x7",
               "fullcover_64_28"]


@pytest.mark.parametrize("name", GRAPH_CASES)
def test_oracle_equals_the_reference_source_run_on_the_tf_shim(name):
    """``tests/golden/graph_*.npz`` were produced by executing the reference's own ``air/transformer.py`` on a numpy
    stand-in for the TF ops it uses (``tests/golden/tf_shim.py``, ``make_golden_graph.py``): the op graph -- which corner
    pairs with which weight, where ``-1.001`` sits, how indices flatten -- is the reference's, only the per-kernel
    numerics (linspace recurrence, K = 3 accumulation order, add_n order) are the shim's stated assumptions."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, g = np.load(os.path.join(here, name + ".npz")), np.load(os.path.join(here, "graph_" + name + ".npz"))
    rows = g["rows"]
    out = R.transformer(z["U"][rows], z["theta"][rows], tuple(int(v) for v in z["out_size"]))
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))
    assert np.array_equal(z["out"][rows].view(np.uint32), g["out"].view(np.uint32))     # and so is the committed golden output


def test_oracle_batch_transformer_equals_the_reference_source_run_on_the_tf_shim():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_batch_transformer.npz"))
    U, th = g["U"], g["thetas"]
    nb, nt = th.shape[:2]
    out = R.transformer(np.repeat(U, nt, axis=0), th.reshape(nb * nt, 6), g["out"].shape[1:3])      # transformer.py:190-194
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))


GRAD_GRAPH_CASES = ["read_50_28", "write_28_50", "adversarial_17x23x3_9x31", "fullcover_64_28", "out_1x7"]


@pytest.mark.parametrize("name", GRAD_GRAPH_CASES)
def test_closed_form_backward_equals_autodiff_of_the_reference_source(name):
    """``tests/golden/graph_grad_*.npz``: the reference's own ``air/transformer.py`` imported on a torch-based stand-in for
    its TF ops (``tests/golden/tf_shim_torch.py``) and differentiated by autograd in float32 -- the role TF autodiff plays in
    the reference (:1098).  The oracle's closed-form backward must agree within the gradient tolerance
    the CUDA kernels are held to."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, g = np.load(os.path.join(here, name + ".npz")), np.load(os.path.join(here, "graph_grad_" + name + ".npz"))
    rows = g["rows"]
    assert H.grad_excess(g["dU"], z["dU"][rows], z["absdU"][rows]) <= 1.0
    assert H.grad_excess(g["dtheta"], z["dtheta"][rows], z["absdtheta"][rows]) <= 1.0
    assert float(np.abs(g["dU"]).max()) > 0 and float(np.abs(g["dtheta"]).max()) > 0


@pytest.mark.parametrize("name", [f"graph_asr_{c}_T{t}" for c in ("c2", "c3", "all") for t in (6, 3)])
def test_asr_oracle_equals_the_reference_source_run_on_the_tf_shim(name):
    """``tests/golden/graph_asr_*.npz``: the reference's own regulariser lines (air_number_bbox_location.py:645-678 and
    :970-1069) exec'd on the torch-based TF shim in float64 and differentiated by autograd
    (``tests/golden/make_golden_asr_graph.py``).  ``oracle/asr_ref.py`` must reproduce values, logged components and all
    three gradients of ``reduce_mean(pr_loss + num_element_min) + num_marginal_loss`` (:1078-1079)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    cfg = {str(k): eval(str(v), {"__builtins__": {}}) for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    r = asr_ref.asr_numpy(g["log_odds"], g["shifts"], g["scales"][..., 0], canvas_size=cfg["canvas_size"], counts=cfg["counts"],
                          max_steps=cfg["max_steps"], gamma_num=cfg["gn"], gamma_margin=cfg["gm"], gamma_elem=cfg["gne"],
                          gamma_bbox=cfg["gb"], gamma_size=cfg["gs"], gamma_area=cfg["ga"], area_minmax=cfg["minmax"])
    tol = dict(rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(r["per_image"], g["per_image"], **tol)
    np.testing.assert_allclose(r["margin"], g["margin"], **tol)
    for mine, theirs in (("area", "area"), ("out", "out_loss"), ("size", "size"), ("overlap", "overlap")):
        np.testing.assert_allclose(r[mine], g[theirs], **tol)
    np.testing.assert_allclose(r["per_image"].mean() + r["margin"], g["loss"], **tol)
    np.testing.assert_allclose(r["d_log_odds"], g["d_log_odds"], **tol)
    np.testing.assert_allclose(r["d_shifts"], g["d_shifts"], **tol)
    np.testing.assert_allclose(r["d_scales"], g["d_scales"][..., 0], **tol)


def test_recon_loss_oracle_equals_the_reference_source_run_on_the_tf_shim():
    """``tests/golden/graph_recon.npz``: the reference's own reconstruction-loss lines (:944-967) exec'd on the torch-based
    TF shim in float64 (clip bounds hit exactly, sums above 1, zero pixels under objects included), gradient by autograd."""
    import os
    from oracle import bce_ref
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_recon.npz"))
    loss, mse = bce_ref.reconstruction_loss(g["canvas"], g["images"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-12)
    np.testing.assert_allclose(mse, g["mse"], rtol=1e-12)
    d = bce_ref.reconstruction_loss_backward(g["canvas"], g["images"], g["w"])
    np.testing.assert_allclose(d, g["dcanvas"], rtol=1e-12, atol=1e-12)
    assert np.abs(g["dcanvas"]).max() > 1e9          # the 1e10 slopes at canvas == 0 are part of the reference


def test_theta_construction_equals_the_reference_source_run_on_the_tf_shim():
    """``tests/golden/graph_thetas.npz``: the ``st_forward`` / ``st_backward`` lines (:511-531, :563-584) exec'd on the shim in
    float32: ``OracleOps.thetas`` reproduces both matrices bit for bit and their gradients to rounding."""
    import os
    import torch
    from oracle.air_ops import OracleOps
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_thetas.npz"))
    s = torch.tensor(g["s"], requires_grad=True)
    xy = torch.tensor(np.stack([g["x"], g["y"]], 1), requires_grad=True)
    th_r, th_w = OracleOps().thetas(xy, s[:, None])
    assert np.array_equal(th_r.detach().numpy().reshape(-1, 2, 3).view(np.uint32), g["theta"].view(np.uint32))
    assert np.array_equal(th_w.detach().numpy().reshape(-1, 2, 3).view(np.uint32), g["theta_recon"].view(np.uint32))
    ((th_r.reshape(-1, 2, 3) * torch.tensor(g["wr"])).sum() + (th_w.reshape(-1, 2, 3) * torch.tensor(g["ww"])).sum()).backward()
    np.testing.assert_allclose(s.grad.numpy(), g["ds"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(xy.grad.numpy()[:, 0], g["dx"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(xy.grad.numpy()[:, 1], g["dy"], rtol=2e-5, atol=1e-5)
