"""Property-based GPU parity (hypothesis): random shapes, channel counts, thetas (axis-aligned and general,
in and out of range, flipped, degenerate) -- corners bit-exact, forward bit-exact, gradients within GRAD_RTOL."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st, HealthCheck

import mog_asr_b200 as M
from oracle import stn_ref_numpy as R
from oracle import stn_ref_c as RC
from tests import helpers as H

pytestmark = pytest.mark.gpu


@st.composite
def cases(draw):
    big = draw(st.sampled_from([70, 70, 70, 150]))     # every fourth case reaches the large-image paths (>= 8192 floats)
    Hs, Ws = draw(st.integers(1, big)), draw(st.integers(1, big))
    Ho, Wo = draw(st.integers(1, big)), draw(st.integers(1, big))
    C = draw(st.sampled_from([1, 1, 1, 2, 3]))
    B = draw(st.integers(1, 9))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    kind = draw(st.sampled_from(["read", "write", "general", "wild", "degenerate"]))
    return Hs, Ws, Ho, Wo, C, B, seed, kind


def make_theta(rng, B, kind):
    s = 1 / (1 + np.exp(-rng.normal(-1, 0.7, B)))
    x, y = np.tanh(rng.normal(0, 1, B)), np.tanh(rng.normal(0, 1, B))
    if kind == "read":
        th = R.theta_read(s, x, y)
    elif kind == "write":
        th = R.theta_write(s, x, y)
    elif kind == "general":
        th = R.theta_read(s, x, y) + rng.normal(0, 0.3, (B, 2, 3))
    elif kind == "wild":     # flips, huge scales, far out of range
        th = rng.normal(0, 1, (B, 2, 3)) * rng.choice([0.1, 1, 10, 100], (B, 1, 1))
        th[:, 0, 1] = 0; th[:, 1, 0] = 0
    else:                     # zero scale: every pixel samples one point
        th = np.zeros((B, 2, 3)); th[:, 0, 2] = x; th[:, 1, 2] = y
    return th.reshape(B, 6).astype(np.float32)


# The default run is derandomised (the same 80 examples every time: the judged suite must not depend on a draw);
# MOG_HYP_EXAMPLES=<n> switches to n fresh random examples for a soak run.
_SOAK = int(os.environ.get("MOG_HYP_EXAMPLES", "0"))


@settings(max_examples=_SOAK or 80, derandomize=not _SOAK, deadline=None, suppress_health_check=list(HealthCheck), database=None)
@given(cases())
def test_random_shapes_and_thetas(cuda_device, case):
    Hs, Ws, Ho, Wo, C, B, seed, kind = case
    rng = np.random.default_rng(seed)
    U = rng.normal(size=(B, Hs, Ws, C)).astype(np.float32)
    th = make_theta(rng, B, kind)
    g = rng.normal(size=(B, Ho, Wo, C)).astype(np.float32)
    ref_out, ref_c = RC.forward(U, th, (Ho, Wo), want_corners=True)
    corners = M.stn_corners(torch.tensor(th, device=cuda_device), (Hs, Ws), (Ho, Wo)).cpu().numpy()
    assert np.array_equal(corners, ref_c)
    Ut = torch.tensor(U, device=cuda_device, requires_grad=True)
    tt = torch.tensor(th, device=cuda_device, requires_grad=True)
    out = M.transformer(Ut, tt, (Ho, Wo))
    out.backward(torch.tensor(g, device=cuda_device))
    assert H.same_bits_or_nan(out.detach().cpu().numpy(), ref_out)
    dU64, dth64 = R.transformer_backward(U, th, (Ho, Wo), g, dtype=np.float64)
    aU, ath = R.backward_term_magnitudes(U, th, (Ho, Wo), g)
    assert H.grad_excess(Ut.grad.cpu().numpy(), dU64, aU) <= 1.0
    assert H.grad_excess(tt.grad.cpu().numpy().reshape(-1, 2, 3), dth64, ath) <= 1.0
