"""Development aid: numpy emulation of the grouped (vertical-first) backward that csrc/mog_stn_bwd.cuh
implements, checked against the oracle's closed form on random separable thetas.  Not product code; it
exists so the algebra (row groups, carry rows, column runs, strips, dtheta from group sums) can be
verified on a CPU box before GPU time is spent.

    python tools/emulate_bwd_grouped.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stn_ref_numpy as R  # noqa: E402

F = np.float32


def axis(t0, t2, n_out, n_src):
    step = F(2.0 / (n_out - 1)) if n_out > 1 else F(0)
    lin = (F(-1) + (step * np.arange(n_out, dtype=F)).astype(F)).astype(F)
    s = ((t0 * lin).astype(F) + F(0)).astype(F) + t2          # (t0*x + +-0) + t2
    p = (((s + F(1)).astype(F) * F(F(n_src) - F(1.001))).astype(F) * F(0.5)).astype(F)
    f = np.fmin(np.fmax(np.floor(p), F(-1)), F(n_src)).astype(np.int64)
    c0 = np.clip(f, 0, n_src - 1)
    c1 = np.clip(f + 1, 0, n_src - 1)
    return lin, c0, c1, (c1.astype(F) - p).astype(F), (p - c0.astype(F)).astype(F)


def bwd_image(U, th, g, strip=64, z=1.0):
    Hs, Ws = U.shape
    Ho, Wo = g.shape
    xt, x0, x1, ax, bx = axis(th[0], th[2], Wo, Ws)
    yt, y0, y1, ay, by = axis(th[4], th[5], Ho, Hs)
    dU = np.zeros((Hs, Ws), F)
    p = np.zeros(7, np.float64)
    inx, iny = np.nonzero(x0 != x1)[0], np.nonzero(y0 != y1)[0]
    if len(inx) == 0 or len(iny) == 0:
        return dU, p[:6].reshape(2, 3), 0.0
    jlo, jhi, ilo, ihi = inx.min(), inx.max(), iny.min(), iny.max()
    asc = not (th[4] < 0)
    rows = list(range(ilo, ihi + 1)) if asc else list(range(ihi, ilo - 1, -1))
    # runs over [jlo, jhi]
    run = {}
    for j in range(jlo, jhi + 1):
        run.setdefault(int(x0[j]), []).append(j)
    for js in range(jlo, jhi + 1, strip):
        je = min(js + strip, jhi + 1)
        J = np.arange(js, je)
        A = np.zeros(len(J), F); Bv = A.copy(); A2 = A.copy(); B2 = A.copy(); car = A.copy()
        ycar = -1

        def hreduce(y, V, first):
            xs = x0[J]
            xlo, xhi = xs.min(), xs.max() + 1
            for x in range(xlo, xhi + 1):
                T = F(0)
                for j in run.get(x, []):
                    if js <= j < je:
                        T += F(ax[j] * z) * V[j - js]
                for j in run.get(x - 1, []):
                    if js <= j < je:
                        T += F(bx[j] * z) * V[j - js]
                if x < Ws:
                    dU[y, x] = T if first else dU[y, x] + T
        first = js == jlo
        for k, i in enumerate(rows):
            gi = g[i, J]
            A += ay[i] * gi; Bv += by[i] * gi; A2 += F(ay[i] * yt[i]) * gi; B2 += F(by[i] * yt[i]) * gi
            last = (k == len(rows) - 1) or (y0[rows[k + 1]] != y0[i])
            if not last:
                continue
            y = int(y0[i])
            Ia, Ic = U[y, x0[J]], U[y, x0[J] + 1]
            Ib, Id = U[y + 1, x0[J]], U[y + 1, x0[J] + 1]
            dxa, dxb, dya, dyc = Ic - Ia, Id - Ib, Ib - Ia, Id - Ic
            sxs = dxa * A + dxb * Bv
            sxy = dxa * A2 + dxb * B2
            E = ax[J] * dya + bx[J] * dyc
            sys_ = E * (A + Bv)
            syy = E * (A2 + B2)
            p[0] += float((xt[J] * sxs).sum()); p[1] += float(sxy.sum()); p[2] += float(sxs.sum())
            p[3] += float((xt[J] * sys_).sum()); p[4] += float(syy.sum()); p[5] += float(sys_.sum())
            p[6] += float(((ax[J] * Ia + bx[J] * Ic) * A + (ax[J] * Ib + bx[J] * Id) * Bv).sum())
            if ycar >= 0 and ycar != y:
                hreduce(ycar, car, first)
                car = np.zeros(len(J), F)
            hreduce(y, A + car, first)
            car = Bv.copy(); ycar = y + 1
            A = np.zeros(len(J), F); Bv = A.copy(); A2 = A.copy(); B2 = A.copy()
        if ycar >= 0:
            hreduce(ycar, car, first)
    hw, hh = F(F(Ws) - F(1.001)) * F(0.5), F(F(Hs) - F(1.001)) * F(0.5)
    dth = np.array([p[0] * hw, p[1] * hw, p[2] * hw, p[3] * hh, p[4] * hh, p[5] * hh]) * z
    return dU, dth.reshape(2, 3), p[6]


def check(Hs, Ws, Ho, Wo, B, seed, strip):
    rng = np.random.default_rng(seed)
    U = rng.random((B, Hs, Ws, 1), dtype=F)
    g = rng.normal(size=(B, Ho, Wo, 1)).astype(F)
    th = np.zeros((B, 6), F)
    th[:, 0] = rng.normal(0, 1.0, B); th[:, 4] = rng.normal(0, 1.0, B)
    th[:, 2] = rng.normal(0, 0.7, B); th[:, 5] = rng.normal(0, 0.7, B)
    th[0] = [0.3, 0, 0.1, 0, 0.3, -0.2]
    th[1] = [3.5, 0, 0.2, 0, 3.7, 0.1]
    dU_ref, dth_ref = R.transformer_backward(U, th, (Ho, Wo), g)
    aU, ath = R.backward_term_magnitudes(U, th, (Ho, Wo), g)
    worst = 0.0
    for b in range(B):
        dU, dth, _ = bwd_image(U[b, :, :, 0], th[b], g[b, :, :, 0], strip=strip)
        eU = np.abs(dU - dU_ref[b, :, :, 0]) / (2e-5 * aU[b, :, :, 0] + 2e-8 * aU[b].max() + 1e-30)
        et = np.abs(dth - dth_ref[b]) / (2e-5 * ath[b] + 2e-8 * ath[b].max() + 1e-30)
        worst = max(worst, float(eU.max()), float(et.max()))
    return worst


if __name__ == "__main__":
    for shape in ((50, 50, 28, 28), (28, 28, 50, 50), (17, 23, 9, 31), (64, 64, 28, 28), (28, 28, 128, 128), (40, 33, 70, 75)):
        for strip in (32, 64):
            w = check(*shape, B=12, seed=1, strip=strip)
            print(shape, "strip", strip, "worst excess (<=1 passes):", round(w, 4))
            assert w <= 1.0
    print("grouped backward emulation agrees with the oracle")
