"""debug aid: tiny backward calls in subprocesses (one crash must not hide the next case)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import mog_asr_b200 as M
from mog_asr_b200 import _lib, synth
from oracle import stn_ref_numpy as R
Hs, Ho, B, mode = %d, %d, %d, %r
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
U = rng.random((B, Hs, Hs, 1), dtype=np.float32)
s, x, y = synth.sxy_prior_like(B, seed=1) if mode != "full" else synth.sxy_full_cover(B, seed=1)
th = synth.theta_read(s, x, y) if Hs > Ho else synth.theta_write(s, x, y)
g = rng.normal(size=(B, Ho, Ho, 1)).astype(np.float32)
L = _lib.load()
dU = torch.full((B, Hs, Hs, 1), 7.0, device=dev); dth = torch.full((B, 6), 7.0, device=dev)
Ud, thd, gd = (torch.tensor(a, device=dev) for a in (U, th, g))
rc = L.mog_stn_backward(Ud.data_ptr(), thd.data_ptr(), gd.data_ptr(), None if mode == "dtheta" else dU.data_ptr(), dth.data_ptr(), B, Hs, Hs, 1, Ho, Ho, 1, None)
torch.cuda.synchronize()
rU, rth = R.transformer_backward(U, th, (Ho, Ho), g)
aU, ath = R.backward_term_magnitudes(U, th, (Ho, Ho), g)
eU = 0.0 if mode == "dtheta" else float(np.max(np.abs(dU.cpu().numpy() - rU) / (2e-5 * aU + 2e-8 * aU.max() + 1e-30)))
et = float(np.max(np.abs(dth.cpu().numpy().reshape(-1, 2, 3) - rth) / (2e-5 * ath + 2e-8 * ath.max() + 1e-30)))
print("rc", rc, "excess dU %%.3g dtheta %%.3g (<=1 passes)" %% (eU, et))
'''
for so in sys.argv[1:] or ["in-tree"]:
    for (Hs, Ho, B, mode) in ((50, 28, 4, "prior"), (50, 28, 4, "dtheta"), (28, 50, 4, "prior"), (28, 50, 64, "full"), (64, 28, 64, "full"),
                              (256, 64, 8, "prior"), (64, 256, 8, "full"), (64, 256, 8, "prior"), (28, 128, 300, "full"), (28, 64, 300, "prior"),
                              (64, 128, 300, "prior"), (28, 256, 40, "full"), (64, 256, 600, "full")):
        env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
        if so != "in-tree":
            env["MOG_SO"] = os.path.join(ROOT, so)
        r = subprocess.run([sys.executable, "-c", CASE % (ROOT, Hs, Ho, B, mode)], env=env, capture_output=True, text=True, timeout=300)
        tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or ["?"])[-1][:160]
        print(so, (Hs, Ho, B, mode), "->", tail, flush=True)
