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


class HostCompositeWriter:
    """The write call site of the AIR loop on host buffers (reference ``air_number_bbox_location.py:592-600`` and
    ``:718-727``): ``T`` windows per image are written onto one canvas per image and the canvas gradient is differentiated
    back to every step's window, theta and z_pres -- ``mog_stn_write_composite_host``.  Arrays are step-major like the
    model's stacks (``windows [T,B,Hw,Ww]``, ``thetas [T,B,6]``, ``z_pres [T,B]``); the canvas crosses the bus once per
    image instead of once per glimpse."""

    def __init__(self, device, window_size, canvas_size, steps, chunk=512, nstreams=3):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostCompositeWriter needs a CUDA device: there is no CPU fallback")
        self.Hw, self.Ww = int(window_size[0]), int(window_size[1])
        self.Hc, self.Wc = int(canvas_size[0]), int(canvas_size[1])
        self.T, self.chunk, self.nstreams = int(steps), int(chunk), int(nstreams)
        L = _lib.load()
        nbytes = L.mog_stn_write_composite_host_workspace_bytes(self.chunk, self.T, self.Hw, self.Ww, self.Hc, self.Wc, self.nstreams)
        if nbytes == 0:
            raise ValueError("bad HostCompositeWriter dimensions")
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.streams = [torch.cuda.Stream(self.device) for _ in range(self.nstreams)]
        self._stream_arr = (ctypes.c_void_p * self.nstreams)(*[s.cuda_stream for s in self.streams])

    def fwd_bwd(self, windows, thetas, z_pres, gcanvas=None, canvas=None, dW=None, dtheta=None, dz=None,
                need_dW=True, need_dtheta=True, need_dz=True):
        """Returns ``(canvas [B,Hc,Wc], dW, dtheta, dz)`` CPU tensors (pinned when allocated here; the gradients are None
        without ``gcanvas`` or when not needed)."""
        T = self.T
        if windows.shape[0] != T or thetas.shape[0] != T or z_pres.shape[0] != T:
            raise ValueError(f"step-major arrays with {T} steps expected")
        B = windows.shape[1]
        grads = gcanvas is not None
        for t, n in ((windows, "windows"), (thetas, "thetas"), (z_pres, "z_pres")) + (((gcanvas, "gcanvas"),) if grads else ()):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{n} must be a contiguous float32 CPU tensor")
        need_dW, need_dtheta, need_dz = need_dW and grads, need_dtheta and grads, need_dz and grads
        if canvas is None:
            canvas = torch.empty((B, self.Hc, self.Wc), dtype=torch.float32).pin_memory()
        if need_dW and dW is None:
            dW = torch.empty_like(windows).pin_memory()
        if need_dtheta and dtheta is None:
            dtheta = torch.empty((T, B, 6), dtype=torch.float32).pin_memory()
        if need_dz and dz is None:
            dz = torch.empty((T, B), dtype=torch.float32).pin_memory()
        L = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(L.mog_stn_write_composite_host(
                windows.data_ptr(), thetas.data_ptr(), z_pres.data_ptr(), gcanvas.data_ptr() if grads else None,
                canvas.data_ptr(), dW.data_ptr() if need_dW else None, dtheta.data_ptr() if need_dtheta else None,
                dz.data_ptr() if need_dz else None, B, T, self.Hw, self.Ww, self.Hc, self.Wc, self.chunk,
                self.workspace.data_ptr(), self.workspace.numel(), self._stream_arr, self.nstreams),
                "mog_stn_write_composite_host")
        return canvas, (dW if need_dW else None), (dtheta if need_dtheta else None), (dz if need_dz else None)
