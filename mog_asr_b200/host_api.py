"""End-to-end (host-buffer) form of the sampler: ``transformer`` + its gradients for arrays that live in
host memory, the way one ``sess.run([window, grads], feed_dict={...})`` of the reference works
(train_air_pr.py:294-295).  Copies, kernels and copies back are chunked over CUDA streams inside
``mog_stn_fwd_bwd_host`` (include/mogstn.h)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class HostSampler:
    """Reusable scratch + streams for ``transformer_fwd_bwd_host`` on one device.  Chunks of 512 images keep the pipeline's
    fill/drain bubbles (one chunk of H2D at the start, one of D2H at the end of every call) at 3 % of a 16 384-image call:
    46.0 GB/s each way against a measured simultaneous-bidirectional ceiling of 46.4 GB/s on this host (2048-image chunks:
    41-43 GB/s)."""

    def __init__(self, device, in_size, out_size, channels=1, chunk=512, nstreams=3, transforms=1):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostSampler needs a CUDA device: there is no CPU fallback")
        self.Hs, self.Ws = int(in_size[0]), int(in_size[1])
        self.Ho, self.Wo = int(out_size[0]), int(out_size[1])
        self.C, self.chunk, self.nstreams, self.T = int(channels), int(chunk), int(nstreams), int(transforms)
        L = _lib.load()
        nbytes = L.mog_stn_batch_host_workspace_bytes(self.chunk, self.T, self.Hs, self.Ws, self.C, self.Ho, self.Wo, self.nstreams)
        if nbytes == 0:
            raise ValueError("bad HostSampler dimensions")
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.streams = [torch.cuda.Stream(self.device) for _ in range(self.nstreams)]
        self._stream_arr = (ctypes.c_void_p * self.nstreams)(*[s.cuda_stream for s in self.streams])

    def fwd_bwd(self, U, theta, gout, out=None, dU=None, dtheta=None, need_dU=True, need_dtheta=True):
        """U ``[B,Hs,Ws,C]``, theta ``[B,6]``, gout ``[B,Ho,Wo,C]``: float32 CPU tensors (pinned for full
        PCIe rate).  Returns ``(out, dU, dtheta)`` CPU tensors (``out``/``dU``/``dtheta`` reuse the given
        buffers when passed)."""
        for t, n in ((U, "U"), (theta, "theta"), (gout, "gout")):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{n} must be a contiguous float32 CPU tensor")
        if self.T != 1:
            raise ValueError("this sampler was built for batch_fwd_bwd (transforms > 1)")
        B = U.shape[0]
        if out is None:
            out = torch.empty((B, self.Ho, self.Wo, self.C), dtype=torch.float32).pin_memory()
        if need_dU and dU is None:
            dU = torch.empty_like(U).pin_memory()
        if need_dtheta and dtheta is None:
            dtheta = torch.empty((B, 6), dtype=torch.float32).pin_memory()
        L = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(L.mog_stn_fwd_bwd_host(
                U.data_ptr(), theta.data_ptr(), gout.data_ptr(), out.data_ptr(),
                dU.data_ptr() if need_dU else None, dtheta.data_ptr() if need_dtheta else None,
                B, self.Hs, self.Ws, self.C, self.Ho, self.Wo, self.chunk,
                self.workspace.data_ptr(), self.workspace.numel(), self._stream_arr, self.nstreams),
                "mog_stn_fwd_bwd_host")
        return out, (dU if need_dU else None), (dtheta if need_dtheta else None)

    def batch_fwd_bwd(self, U, thetas, gout=None, out=None, dU=None, dtheta=None, need_dU=True, need_dtheta=True):
        """``batch_transformer`` form (air/transformer.py:178-195): U ``[B,Hs,Ws,C]``, thetas ``[B,T,6]``, gout ``[B,T,Ho,Wo,C]``
        (``T`` = the ``transforms`` given to the constructor).  Every source image is uploaded once for its T transforms; dU
        (summed over them) comes back once, or not at all with ``need_dU=False`` -- the AIR read call site.
        Returns ``(out [B*T,Ho,Wo,C], dU, dtheta [B,T,6])``."""
        T = self.T
        if thetas.shape[1] != T:
            raise ValueError(f"thetas has {thetas.shape[1]} transforms per image, this sampler was built for {T}")
        need_grad = (need_dU or need_dtheta) and gout is not None
        for t, n in ((U, "U"), (thetas, "thetas")) + (((gout, "gout"),) if need_grad else ()):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{n} must be a contiguous float32 CPU tensor")
        B = U.shape[0]
        if out is None:
            out = torch.empty((B * T, self.Ho, self.Wo, self.C), dtype=torch.float32).pin_memory()
        need_dU, need_dtheta = need_dU and need_grad, need_dtheta and need_grad
        if need_dU and dU is None:
            dU = torch.empty_like(U).pin_memory()
        if need_dtheta and dtheta is None:
            dtheta = torch.empty((B, T, 6), dtype=torch.float32).pin_memory()
        L = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(L.mog_stn_batch_fwd_bwd_host(
                U.data_ptr(), thetas.data_ptr(), gout.data_ptr() if need_grad else None, out.data_ptr(),
                dU.data_ptr() if need_dU else None, dtheta.data_ptr() if need_dtheta else None,
                B, T, self.Hs, self.Ws, self.C, self.Ho, self.Wo, self.chunk,
                self.workspace.data_ptr(), self.workspace.numel(), self._stream_arr, self.nstreams),
                "mog_stn_batch_fwd_bwd_host")
        return out, (dU if need_dU else None), (dtheta if need_dtheta else None)
