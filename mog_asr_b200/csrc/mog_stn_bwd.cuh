// Backward of the sampler for separable thetas, one channel: "grouped" gather form (sm_100a).
//
// TF autodiff of air/transformer.py:102-116 is dU = W_y^T (G W_x) plus six sums for dtheta.
// The output rows that share a source row y0 form a *group* (consecutive rows: the coordinate map is
// monotone).  For a fixed output column j every pixel of a group reads the same four taps, so the warp
//   1. streams the in-range rows of g once, lanes along the output columns (coalesced), accumulating per
//      lane  A = sum ay*g,  Bv = sum by*g,  A2 = sum ay*yt*g,  B2 = sum by*yt*g        (4 FMA per pixel);
//   2. at the end of a group (once per source row, not once per output row) loads the four taps and adds
//        sum sx      += (Ic-Ia)*A  + (Id-Ib)*Bv          sum sx*yt += (Ic-Ia)*A2 + (Id-Ib)*B2
//        sum sy      += E*(A+Bv)                         sum sy*yt += E*(A2+B2),  E = ax*(Ib-Ia) + bx*(Id-Ic)
//      (ay + by = 1 inside the range); the x_t-weighted sums are formed per lane at the end of the image;
//   3. forms the gradient row of source row y as V = A (this group) + Bv (previous group) and reduces
//      ax*V and bx*V over the column runs (run(x) = the consecutive output columns whose left tap is x)
//      with warp shuffles; the first lane of every run adds the two sums to a per-warp shared-memory row
//      indexed by source column, which the warp then writes out coalesced: each dU row is written once,
//      no atomics, deterministic.
// Column parameters are recomputed from theta where they are needed (a dozen flops) instead of being
// tabulated; the per-warp shared memory is a 16-byte entry per in-range output row (in stream order) and
// one source row of floats -- a few KB, so occupancy is bounded by registers, not by shared memory.
#pragma once
#include "mog_stn_warp.cuh"

namespace mog {

#ifndef MOG_BWD2_RB
#define MOG_BWD2_RB 4
#endif
#ifndef MOG_BWD2_MINB
#define MOG_BWD2_MINB 12
#endif

constexpr int kPend2 = 2 * MOG_BWD2_RB * 32;   // floats of gradient rows a batch can leave pending (2 per output row x strip width)
// per-warp shared memory (4-byte words): row table 4*Ho ({y0 byte offset | last-of-group, ay, by, yt} per in-range row,
// in stream order) | pending gradient rows by output column (kPend2) and their source-row offsets | one gradient
// row by source column, Ws + 1 floats
__host__ __device__ inline int bwd2_warp_smem_words(const Geo& g) { return (4 * g.Ho + kPend2 + 2 * MOG_BWD2_RB + g.Ws + 1 + 3) & ~3; }

struct RowP {      // one output row of a separable theta
    int yoff;      // byte offset of source row y0 (clipped)
    bool in;       // y0 != y1: the row is inside the source range
    float ay, by, yt;
};
__device__ __forceinline__ RowP row_params(const Theta& th, const Geo& g, int i, int ws4) {
    RowP r;
    r.yt = lin_at(i, g.step_h);
    const Axis Y = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, r.yt), g.hsc, g.Hs);
    r.yoff = Y.c0 * ws4;
    r.in = Y.c0 != Y.c1;
    r.ay = Y.a;
    r.by = Y.b;
    return r;
}

template <bool COMPOSITE, int NJC>
__global__ void __launch_bounds__(kWarpThreads, MOG_BWD2_MINB) stn_bwd_group_kernel(const BwdArgs a) {
    extern __shared__ int4 s_dyn[];
    constexpr int SW = 32 * NJC;      // output columns per strip
    constexpr int kRB2 = NJC == 1 ? MOG_BWD2_RB : MOG_BWD2_RB / 2;   // output rows per load batch (RB * NJC loads of g in flight)
    const Geo& g = a.g;
    const int C = g.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* s_base = reinterpret_cast<int*>(s_dyn) + warp * bwd2_warp_smem_words(g);
    int4* s_row = reinterpret_cast<int4*>(s_base);
    float* s_v = reinterpret_cast<float*>(s_base + 4 * g.Ho);   // pending gradient rows (by output column), kSlots x SW
    int* s_sloty = reinterpret_cast<int*>(s_v + kPend2);         // their source-row byte offsets
    float* s_x = reinterpret_cast<float*>(s_sloty + 2 * MOG_BWD2_RB);   // gradient row under construction, by source column
    for (int x = lane; x <= g.Ws; x += 32) s_x[x] = 0.f;
    __syncwarp();
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const int SC = g.S * C;
    const int ws4 = g.Ws * 4;
    const float half_wsc = g.wsc * 0.5f, half_hsc = g.hsc * 0.5f;
    const bool bulk = a.dU && a.coop_zero == 2;
    unsigned zero_smem = 0;
    if (bulk) {  // zeroed block behind the per-warp areas: source of the bulk zero fills
        int* zb = reinterpret_cast<int*>(s_dyn) + kWarpsPerCta * bwd2_warp_smem_words(g);
        bulk_zero_init(zb);
        zero_smem = (unsigned)__cvta_generic_to_shared(zb);
    }
    const bool need_taps = a.dtheta != nullptr || (COMPOSITE && a.dz != nullptr);

    for (long long g0 = (long long)blockIdx.x * kWarpsPerCta; g0 < a.Bsrc; g0 += nwarps) {
        if (a.dU && a.coop_zero == 1) {   // the CTA zero-fills its group's contiguous dU region in one sweep
            const long long ng = min((long long)kWarpsPerCta, a.Bsrc - g0);
            __syncthreads();
            fill_zero_cta(a.dU + g0 * (long long)SC, ng * (long long)SC);
            __syncthreads();
        }
        const long long bs = g0 + warp;
        if (bs >= a.Bsrc) continue;
        const float* __restrict__ Ub = a.U + bs * (long long)SC;
        float* __restrict__ dUb = a.dU ? a.dU + bs * (long long)SC : nullptr;
        if (dUb && a.coop_zero == 0) {
            fill_zero(dUb, 0, SC, lane);
            __syncwarp();
        }

        for (int t = 0; t < a.u_div; ++t) {
            const long long b = bs * a.u_div + t;
            Theta th;
            th.load(a.theta + 6 * b);
            const bool sep = th.separable() && C == 1;
            float z = 1.0f;
            bool active = true;
            if (COMPOSITE) {
                z = __ldg(a.z_pres + b);
                active = a.stop_sum ? (__ldg(a.stop_sum + b) < a.threshold) : true;
            }
            const float* __restrict__ gb = a.gout + b * (long long)g.N * C;
            float p[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // dtheta (6) + dz

            if (!active) {
                if (bulk) fill_zero_bulk(dUb, 0, SC, lane, zero_smem);   // dU of an inactive image is zero
            } else if (!sep) {
                if (bulk) {
                    fill_zero(dUb, 0, SC, lane);
                    __syncwarp();
                }
                bwd_general_image<COMPOSITE>(Ub, dUb, gb, a.dtheta ? a.dtheta + 6 * b : nullptr,
                                             (COMPOSITE && a.dz) ? a.dz + b : nullptr, th.t[0], th.t[1], th.t[2], th.t[3],
                                             th.t[4], th.t[5], z, false, lane, g.Hs, g.Ws, g.C, g.Ho, g.Wo, g.step_w,
                                             g.step_h, g.wsc, g.hsc);
                continue;
            } else {
                const bool need_dU = dUb != nullptr;
                const bool first_write = (t == 0);
                // ---- in-range intervals of rows and columns (one interval each: the coordinate map is monotone) ----
                int ilo = g.Ho, ihi = -1, jlo = g.Wo, jhi = -1;
                for (int i = lane; i < g.Ho; i += 32)
                    if (row_params(th, g, i, ws4).in) { ilo = min(ilo, i); ihi = max(ihi, i); }
                for (int j = lane; j < g.Wo; j += 32) {
                    const Axis X = col_axis(th, g, j);
                    if (X.c0 != X.c1) { jlo = min(jlo, j); jhi = max(jhi, j); }
                }
                ilo = __reduce_min_sync(0xffffffffu, ilo); ihi = __reduce_max_sync(0xffffffffu, ihi);
                jlo = __reduce_min_sync(0xffffffffu, jlo); jhi = __reduce_max_sync(0xffffffffu, jhi);
                const bool any = ihi >= ilo && jhi >= jlo;
                const bool ascending = !(th.t[4] < 0.0f);   // stream order that makes y0 non-decreasing
                const int nrows = ihi - ilo + 1;
                if (bulk) {
                    // dU rows outside the band the stream touches go to the copy engine; the band is zeroed with
                    // ordinary stores (ordered before the warp's own gradient rows by the __syncwarp below)
                    if (any) {
                        const int ya = row_params(th, g, ilo, ws4).yoff, yb = row_params(th, g, ihi, ws4).yoff;
                        const int ylo = min(ya, yb) / ws4, yend = max(ya, yb) / ws4 + 2;   // band rows [ylo, yend)
                        fill_zero_bulk(dUb, 0, ylo * g.Ws, lane, zero_smem);
                        fill_zero_bulk(dUb, yend * g.Ws, SC, lane, zero_smem);
                        fill_zero(dUb, ylo * g.Ws, yend * g.Ws, lane);
                    } else {
                        fill_zero_bulk(dUb, 0, SC, lane, zero_smem);
                    }
                }
                __syncwarp();
                if (any) {
                    // row table in stream order: entry ii describes output row (ascending ? ilo + ii : ihi - ii)
                    for (int ii = lane; ii < nrows; ii += 32) {
                        const RowP r = row_params(th, g, ascending ? ilo + ii : ihi - ii, ws4);
                        bool last = ii + 1 == nrows;
                        if (!last) last = row_params(th, g, ascending ? ilo + ii + 1 : ihi - ii - 1, ws4).yoff != r.yoff;
                        s_row[ii] = make_int4(r.yoff | (last ? 1 : 0), __float_as_int(r.ay), __float_as_int(r.by), __float_as_int(r.yt));
                    }
                    __syncwarp();
                    const char* Ubc = opaque(reinterpret_cast<const char*>(Ub));
                    char* dUbc = reinterpret_cast<char*>(dUb);
                    const int gstep = ascending ? g.Wo : -g.Wo;
                    const float* gfirst = gb + (long long)(ascending ? ilo : ihi) * g.Wo;

                    // ---- strips of SW output columns; every strip streams the in-range rows once ----
                    for (int js = jlo; js <= jhi; js += SW) {
                        const int je = min(js + SW, jhi + 1);   // strip = columns [js, je)
                        const bool store_plain = first_write && js == jlo;   // later strips / transforms accumulate
                        int xo[NJC], seg[NJC];   // byte offset of x0; (lanes after this one in the same run) | head-of-run << 8
                        float caz[NJC], cbz[NJC], cax[NJC], cbx[NJC];
                        bool val[NJC];
                        const float* gcol[NJC];
                        int rmax = 1;   // longest run inside one 32-column chunk
#pragma unroll
                        for (int c = 0; c < NJC; ++c) {
                            const int j = js + 32 * c + lane;
                            val[c] = j < je;
                            const int jc = val[c] ? j : je - 1;      // masked lanes shadow the strip's last column (their g is 0)
                            const Axis X = col_axis(th, g, jc);
                            xo[c] = X.c0 * 4;
                            cax[c] = X.a; cbx[c] = X.b;
                            caz[c] = COMPOSITE ? X.a * z : X.a;
                            cbz[c] = COMPOSITE ? X.b * z : X.b;
                            gcol[c] = gfirst + jc;
                            // run structure inside the chunk (masked lanes belong to no run and never write)
                            const int xprev = __shfl_up_sync(0xffffffffu, xo[c], 1);
                            const int vprev = __shfl_up_sync(0xffffffffu, (int)val[c], 1);
                            const bool cont = lane > 0 && val[c] && vprev != 0 && xprev == xo[c];   // continues lane - 1's run
                            const unsigned eq_prev = __ballot_sync(0xffffffffu, cont);
                            const unsigned above = lane == 31 ? 0u : (eq_prev >> (lane + 1));
                            const int follow = __ffs(~above) - 1;   // lanes after this one in the same run
                            seg[c] = follow | ((val[c] && !cont) ? 256 : 0);
                            rmax = max(rmax, follow + 1);
                        }
                        rmax = __reduce_max_sync(0xffffffffu, rmax);
                        int xlo = 0, nxs = 0;
                        if (need_dU) {
                            const int xa = col_axis(th, g, js).c0, xb = col_axis(th, g, je - 1).c0;
                            xlo = min(xa, xb);
                            nxs = (max(xa, xb) + 1 - xlo + 32) >> 5;   // source columns [xlo, max + 1]
                        }
                        float A[NJC], Bv[NJC], A2[NJC], B2[NJC], car[NJC], SX[NJC], SXY[NJC], SY[NJC], SYY[NJC], SZ[NJC];
#pragma unroll
                        for (int c = 0; c < NJC; ++c)
                            A[c] = Bv[c] = A2[c] = B2[c] = car[c] = SX[c] = SXY[c] = SY[c] = SYY[c] = SZ[c] = 0.f;
                        int ycar = -1;   // byte offset of the source row the carry belongs to (-1: none pending)

                        // A batch leaves its finished gradient rows (by output column) pending in s_v; they are then turned
                        // into source rows one at a time: run sums by shuffles, the first lane of every run adds them to
                        // s_x (left tap, then right tap), s_x is written out coalesced and cleared.
                        bool done = false;
                        for (int ii0 = 0; !done; ii0 += kRB2) {
                            int nslots = 0;
                            if (ii0 < nrows) {
                                const int nb = min(kRB2, nrows - ii0);
                                // ---- loads of the batch: g of every row, the four taps at the end of every group ----
                                int ey[kRB2];   // y0 byte offset | last-of-group (tail rows shadow the last valid row, never 'last')
                                float gq[NJC][kRB2], I[NJC][kRB2][4];
#pragma unroll
                                for (int r = 0; r < kRB2; ++r) {
                                    const int ii = ii0 + min(r, nb - 1);
                                    ey[r] = s_row[ii].x;
                                    if (r >= nb) ey[r] &= ~1;
#pragma unroll
                                    for (int c = 0; c < NJC; ++c) {
                                        gq[c][r] = __ldg(gcol[c] + ii * gstep);
                                        if ((ey[r] & 1) && need_taps) {
                                            const char* pa = Ubc + (unsigned)((ey[r] & ~3) + xo[c]);
                                            I[c][r][0] = ldg_f32(pa);        I[c][r][2] = ldg_f32(pa + 4);
                                            I[c][r][1] = ldg_f32(pa + ws4);  I[c][r][3] = ldg_f32(pa + ws4 + 4);
                                        }
                                    }
                                }
                                // ---- arithmetic ----
#pragma unroll
                                for (int r = 0; r < kRB2; ++r) {
                                    if (r < nb) {
                                        const int4 er = s_row[ii0 + r];
                                        const float ay = __int_as_float(er.y), by = __int_as_float(er.z);
                                        const float ayt = ay * __int_as_float(er.w), byt = by * __int_as_float(er.w);
#pragma unroll
                                        for (int c = 0; c < NJC; ++c) {
                                            const float gv = val[c] ? gq[c][r] : 0.f;
                                            A[c] = fmaf(ay, gv, A[c]);   Bv[c] = fmaf(by, gv, Bv[c]);
                                            A2[c] = fmaf(ayt, gv, A2[c]); B2[c] = fmaf(byt, gv, B2[c]);
                                        }
                                        if (ey[r] & 1) {
                                            const int yoff = ey[r] & ~3;
                                            if (need_taps) {
#pragma unroll
                                                for (int c = 0; c < NJC; ++c) {
                                                    const float Ia = I[c][r][0], Ib = I[c][r][1], Ic = I[c][r][2], Id = I[c][r][3];
                                                    const float dxa = Ic - Ia, dxb = Id - Ib;
                                                    SX[c] = fmaf(dxa, A[c], fmaf(dxb, Bv[c], SX[c]));
                                                    SXY[c] = fmaf(dxa, A2[c], fmaf(dxb, B2[c], SXY[c]));
                                                    const float E = fmaf(cax[c], Ib - Ia, cbx[c] * (Id - Ic));
                                                    SY[c] = fmaf(E, A[c] + Bv[c], SY[c]);
                                                    SYY[c] = fmaf(E, A2[c] + B2[c], SYY[c]);
                                                    if (COMPOSITE)   // dz = sum g * sample  (:722-727)
                                                        SZ[c] = fmaf(fmaf(cax[c], Ia, cbx[c] * Ic), A[c],
                                                                     fmaf(fmaf(cax[c], Ib, cbx[c] * Id), Bv[c], SZ[c]));
                                                }
                                            }
                                            if (need_dU) {
                                                if (ycar >= 0 && ycar != yoff) {   // the previous group's lower row stands alone
                                                    if (lane == 0) s_sloty[nslots] = ycar;
#pragma unroll
                                                    for (int c = 0; c < NJC; ++c) { s_v[nslots * SW + 32 * c + lane] = car[c]; car[c] = 0.f; }
                                                    ++nslots;
                                                }
                                                if (lane == 0) s_sloty[nslots] = yoff;
#pragma unroll
                                                for (int c = 0; c < NJC; ++c) { s_v[nslots * SW + 32 * c + lane] = A[c] + car[c]; car[c] = Bv[c]; }
                                                ++nslots;
                                                ycar = yoff + ws4;
                                            }
#pragma unroll
                                            for (int c = 0; c < NJC; ++c) A[c] = Bv[c] = A2[c] = B2[c] = 0.f;
                                        }
                                    }
                                }
                            } else {
                                done = true;
                                if (need_dU && ycar >= 0) {   // lower row of the last group
                                    if (lane == 0) s_sloty[0] = ycar;
#pragma unroll
                                    for (int c = 0; c < NJC; ++c) s_v[32 * c + lane] = car[c];
                                    nslots = 1;
                                }
                            }
                            if (nslots > 0) {
                                __syncwarp();
#pragma unroll 1
                                for (int k = 0; k < nslots; ++k) {
                                    const int yoff = s_sloty[k];
#pragma unroll
                                    for (int c = 0; c < NJC; ++c) {
                                        const float v = s_v[k * SW + 32 * c + lane];
                                        const float va0 = caz[c] * v, vb0 = cbz[c] * v;
                                        float va = va0, vb = vb0;
                                        const int follow = seg[c] & 255;
#pragma unroll 1
                                        for (int d = 1; d < rmax; ++d) {   // (the lanes' own values travel, not their partial sums)
                                            const float ua = __shfl_down_sync(0xffffffffu, va0, d), ub = __shfl_down_sync(0xffffffffu, vb0, d);
                                            if (d <= follow) { va += ua; vb += ub; }
                                        }
                                        // (a lane only ever adds lanes of its own run, so the head's sum is the run's)
                                        const bool head = seg[c] >= 256;
                                        float* px = reinterpret_cast<float*>(reinterpret_cast<char*>(s_x) + xo[c]);
                                        if (head) px[0] += va;
                                        __syncwarp();
                                        if (head) px[1] += vb;
                                        __syncwarp();
                                    }
#pragma unroll 1
                                    for (int cx = 0; cx < nxs; ++cx) {
                                        const int x = xlo + 32 * cx + lane;
                                        if (x < g.Ws) {
                                            emit_px(dUbc + yoff + x * 4, s_x[x], true, store_plain);
                                            s_x[x] = 0.f;
                                        }
                                    }
                                    __syncwarp();
                                }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < NJC; ++c) {
                            if (val[c]) {
                                const float xt = lin_at(js + 32 * c + lane, g.step_w);
                                p[0] = fmaf(xt, SX[c], p[0]); p[1] += SXY[c]; p[2] += SX[c];
                                p[3] = fmaf(xt, SY[c], p[3]); p[4] += SYY[c]; p[5] += SY[c];
                                if (COMPOSITE) p[6] += SZ[c];
                            }
                        }
                    }  // strips
                }
                // dx_s = dx*(Ws-1.001)/2, dy_s = dy*(Hs-1.001)/2 (transformer.py:75-76); the composite's z*g (:724-726)
                const float sw_ = COMPOSITE ? half_wsc * z : half_wsc, sh_ = COMPOSITE ? half_hsc * z : half_hsc;
                p[0] *= sw_; p[1] *= sw_; p[2] *= sw_;
                p[3] *= sh_; p[4] *= sh_; p[5] *= sh_;
            }
            // dtheta / dz: warp shuffle reduction (one warp per image: this is also the block reduction)
#pragma unroll
            for (int k = 0; k < 7; ++k) p[k] = warp_sum(p[k]);
            if (lane == 0) {
                if (a.dtheta) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) a.dtheta[6 * b + k] = p[k];
                }
                if (COMPOSITE && a.dz) a.dz[b] = p[6];
            }
        }
    }
    if (bulk) bulk_zero_drain(lane);
}

}  // namespace mog
