#!/usr/bin/env python
"""Development probe: GB/s each way of the two host entry points of the e2e leg for several chunk sizes / stream counts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mog_asr_b200 import synth
from mog_asr_b200.host_api import HostCompositeWriter, HostSampler

dev = torch.device("cuda", 0)
B, cs, gs, T = 16384, 256, 64, 8
pin = lambda *s: torch.zeros(s, dtype=torch.float32).pin_memory()
U, g_r, out_r, dU_r, dth_r = pin(B, cs, cs, 1), pin(B, T, gs, gs, 1), pin(B * T, gs, gs, 1), pin(B, cs, cs, 1), pin(B, T, 6)
W, g_c, z, canvas, dW, dth_w, dz = pin(T, B, gs, gs), pin(B, cs, cs), pin(T, B), pin(B, cs, cs), pin(T, B, gs, gs), pin(T, B, 6), pin(T, B)
U.uniform_(); W.uniform_(); z.uniform_(); g_r.normal_(); g_c.normal_()
thr, thw = [], []
for t in range(T):
    s, x, y = synth.sxy_prior_like(B, seed=100 + t)
    thr.append(synth.theta_read(s, x, y)); thw.append(synth.theta_write(s, x, y))
thr = torch.from_numpy(np.ascontiguousarray(np.stack(thr, 1))).pin_memory()
thw = torch.from_numpy(np.ascontiguousarray(np.stack(thw, 0))).pin_memory()
rbytes = 4 * (U.numel() + g_r.numel())
wbytes = 4 * (W.numel() + g_c.numel())

def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

for chunk, ns in ((128, 3), (192, 3), (128, 2)):
    rd = HostSampler(dev, (cs, cs), (gs, gs), 1, chunk=chunk, nstreams=ns, transforms=T)
    dt = timeit(lambda: rd.batch_fwd_bwd(U, thr, g_r, out=out_r, dU=dU_r, dtheta=dth_r))
    print(f"read  chunk {chunk:4d} streams {ns}: {dt*1e3:7.1f} ms  {rbytes/dt/1e9:5.1f} GB/s each way", flush=True)
    del rd
for chunk, ns in ((512, 3), (512, 4), (512, 6), (1024, 3), (1024, 4), (384, 6)):
    wr = HostCompositeWriter(dev, (gs, gs), (cs, cs), steps=T, chunk=chunk, nstreams=ns)
    dt = timeit(lambda: wr.fwd_bwd(W, thw, z, g_c, canvas=canvas, dW=dW, dtheta=dth_w, dz=dz))
    print(f"write chunk {chunk:4d} streams {ns}: {dt*1e3:7.1f} ms  {wbytes/dt/1e9:5.1f} GB/s each way", flush=True)
    del wr
