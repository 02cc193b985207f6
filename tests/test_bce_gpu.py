"""GPU parity of the fused reconstruction-loss kernels vs the oracle (oracle/bce_ref.py).
Tolerance: fp32 logf and an fp32 warp-tree sum vs the fp64 oracle: 2e-6 * sum|terms| per image; the gradient is
elementwise (one division each side): rtol 2e-6."""
import numpy as np
import pytest
import torch

from mog_asr_b200.recon import reconstruction_loss
from oracle import bce_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,P", [(64, 2500), (256, 4096), (7, 37), (1, 1), (33, 2501)])
def test_bce_forward_backward(cuda_device, B, P):
    rng = np.random.default_rng(B * 10007 + P)
    # canvases like running_recon: many exact zeros, values above 1 where windows overlap, a few negatives (residue)
    c = rng.random((B, P), dtype=np.float32) * 1.4 - 0.1
    c[rng.random((B, P)) < 0.5] = 0.0
    c[rng.random((B, P)) < 0.01] = 1.0
    x = np.clip(rng.random((B, P), dtype=np.float32) * 1.2, 0, 1).astype(np.float32)
    x[rng.random((B, P)) < 0.6] = 0.0
    g = rng.normal(size=B).astype(np.float32)
    ct = torch.tensor(c, device=cuda_device, requires_grad=True)
    loss, mse = reconstruction_loss(ct, torch.tensor(x, device=cuda_device))
    loss.backward(torch.tensor(g, device=cuda_device))
    ref_loss, ref_mse = bce_ref.reconstruction_loss(c, x)
    mag = bce_ref.term_magnitudes(c, x)
    assert np.all(np.abs(loss.detach().cpu().numpy() - ref_loss) <= 2e-6 * mag + 1e-12)
    np.testing.assert_allclose(mse.cpu().numpy(), ref_mse, rtol=2e-6, atol=1e-12)
    ref_d = bce_ref.reconstruction_loss_backward(c, x, g)
    got_d = ct.grad.cpu().numpy()
    r = np.clip(c.astype(np.float64), 0, 1)
    dmag = (np.abs(x / (r + 1e-10)) + np.abs((1 - x) / (1 - r + 1e-10))) * np.abs(g)[:, None]   # the two terms may cancel
    passes = (c >= 0) & (c <= 1)
    assert np.all(got_d[~passes] == 0)                     # the clip mask (inclusive bounds) is exact
    assert np.all(np.abs(got_d - ref_d) <= 2e-6 * dmag + 1e-30)


def test_bce_matches_reference_ops_on_device(cuda_device):
    """Same formula with torch ops on the device (what the re-hosted model did before the fusion)."""
    g = torch.Generator(device=cuda_device).manual_seed(5)
    c = (torch.rand((128, 2500), device=cuda_device, generator=g) * 1.3).requires_grad_(True)
    x = torch.rand((128, 2500), device=cuda_device, generator=g)
    loss, _ = reconstruction_loss(c, x)
    r = torch.clamp(c.detach(), 0, 1)
    ref = -(x * torch.log(r + 1e-10) + (1 - x) * torch.log(1 - r + 1e-10)).sum(1)
    assert torch.allclose(loss, ref, rtol=1e-5)


def test_bce_kernel_vs_the_reference_source_run_on_the_tf_shim(cuda_device):
    """Directly against ``tests/golden/graph_recon.npz`` (the reference's own lines :944-967 on the torch TF shim, float64)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_recon.npz"))
    c32, x32 = g["canvas"].astype(np.float32), g["images"].astype(np.float32)
    ct = torch.tensor(c32, device=cuda_device, requires_grad=True)
    loss, mse = reconstruction_loss(ct, torch.tensor(x32, device=cuda_device))
    loss.backward(torch.tensor(g["w"].astype(np.float32), device=cuda_device))
    # the golden inputs are float64; evaluate the oracle (itself pinned to the golden file on the CPU) on the fp32-rounded inputs
    ref_loss, ref_mse = bce_ref.reconstruction_loss(c32, x32)
    mag = bce_ref.term_magnitudes(c32, x32)
    assert np.all(np.abs(loss.detach().cpu().numpy() - ref_loss) <= 2e-6 * mag + 1e-12)
    assert np.all(np.abs(ref_loss - g["loss"]) <= 1e-5 * mag)                 # fp32 rounding of the inputs only
    np.testing.assert_allclose(mse.cpu().numpy(), ref_mse, rtol=2e-6, atol=1e-12)
    d = ct.grad.cpu().numpy()
    passes = (c32 >= 0) & (c32 <= 1)
    assert np.all(d[~passes] == 0)
    ref_d = bce_ref.reconstruction_loss_backward(c32, x32, g["w"].astype(np.float32))
    r = np.clip(c32.astype(np.float64), 0, 1)
    dmag = (np.abs(x32 / (r + 1e-10)) + np.abs((1 - x32) / (1 - r + 1e-10))) * np.abs(g["w"].astype(np.float32))[:, None]
    assert np.all(np.abs(d - ref_d) <= 2e-6 * dmag + 1e-30)
    assert np.abs(d).max() > 1e9
