// Backward of the sampler, write direction (source window <= 64 x 64, output up to 256 x 256): one CTA per image (sm_100a).
//
// Profile of the warp-per-image kernel on this direction (profiles/r02_*): 62 % of its instructions sit in the per-pixel
// phase, most of them 64-bit address arithmetic for the four taps; occupancy is 12 % because every warp carries its own
// axis tables and gather rows (17 KB at a 256-wide output); the top stall is the taps' long scoreboard.  Here
//   * the four warps of a CTA share ONE image: the source window is staged whole in shared memory -- by one
//     cp.async.bulk.tensor (3-D tensor map over [B][Hs][Ws], box = the image) when the row pitch allows it, else by
//     coalesced loads -- so a tap is an LDS at base + immediate; axis tables, run table and gather rows exist once per
//     CTA (40 KB at 256 <- 64: five CTAs = 20 warps per SM);
//   * a batch of RB (4 or 8) in-range output rows is processed by all four warps: phase 1 splits it by 32-column chunk of the
//     OUTPUT: g (coalesced) x taps -> dtheta partial sums in registers and ax*g, bx*g into the shared gather rows;
//     phase 2 splits it by 32-column chunk of the SOURCE: T[x] = sum over run(x) of ax*g + sum over run(x-1) of bx*g,
//     folded into two running source rows (y0, y0 + 1) held in registers by the chunk's owner warp; a source row is
//     stored once, coalesced, when the stream moves past it.  With two buffers of gather rows there is one barrier per
//     batch and the warps that own no source chunk run phase 1 of the next batch meanwhile.  No atomics; dU is
//     deterministic.
//   * dtheta / dz: warp shuffles, then a block reduction through shared memory.
// The arithmetic is that of stn_bwd_warp_kernel (same products, same order along a row).
#pragma once
#include "mog_stn_bwd_tma.cuh"

namespace mog {

constexpr int kCtaWarps = 4;
constexpr int kCtaThreads = 32 * kCtaWarps;
constexpr int kCtaMaxWs = 64;        // source columns: at most 3 chunks of 32 carry gradient (footprint <= Ws + 2), one per warp
#ifndef MOG_BWD_CTA_MINB
#define MOG_BWD_CTA_MINB 5
#endif

struct CtaLayout {
    int U, row, col, run, ga, gb, red, bar, total;   // byte offsets
};
// RB = output rows per batch; NBUF = buffers of gather rows (2: one barrier per batch, phase 1 of the next batch overlaps
// phase 2 of this one)
__host__ __device__ inline CtaLayout bwd_cta_layout(const Geo& g, int RB, int NBUF) {
    CtaLayout l;
    int o = 0;
    l.U = o;   o += align128(g.S * 4);
    l.row = o; o += g.Ho * 16;
    l.col = o; o += g.Wo * 16;
    l.run = o; o += align128(g.Ws * 4);
    l.ga = o;  o += align128(NBUF * RB * (g.Wo + 1) * 4);
    l.gb = o;  o += align128(NBUF * RB * (g.Wo + 1) * 4);
    l.red = o; o += 128 + kCtaWarps * 32;
    l.bar = o; o += 128;
    l.total = o;
    return l;
}

template <bool COMPOSITE, int kCtaRB, int NBUF>
__global__ void __launch_bounds__(kCtaThreads, MOG_BWD_CTA_MINB)
stn_bwd_cta_kernel(const __grid_constant__ CUtensorMap tmU, const BwdArgs a, const int use_tma) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const Geo& g = a.g;
    const CtaLayout L = bwd_cta_layout(g, kCtaRB, NBUF);
    float* s_U = reinterpret_cast<float*>(s_raw + L.U);
    int4* s_row = reinterpret_cast<int4*>(s_raw + L.row);    // {y0 * Ws * 4, y_t, ay, by} per output row; after the interval
                                                             // scan: entry ii = stream row ii (the order that makes y0 non-decreasing)
    int4* s_col = reinterpret_cast<int4*>(s_raw + L.col);    // {x0 * 4, x_t, ax, bx} per output column
    int* s_run = reinterpret_cast<int*>(s_raw + L.run);      // per source column: start | end << 16 of its run (0 = empty)
    float* s_ga = reinterpret_cast<float*>(s_raw + L.ga);    // [2][kCtaRB][Wo + 1]: ax * g (column Wo is a zero slot)
    float* s_gb = reinterpret_cast<float*>(s_raw + L.gb);
    int* s_redi = reinterpret_cast<int*>(s_raw + L.red);     // 5 ints of block reductions (ilo, ihi, jlo, jhi, rmax) per warp ...
    float* s_redf = reinterpret_cast<float*>(s_raw + L.red + 128);   // ... and 7 floats per warp for dtheta / dz
    uint64_t* bar_U = reinterpret_cast<uint64_t*>(s_raw + L.bar);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = g.Wo + 1;
    const int ws4 = g.Ws * 4;
    const int SC = g.S;
    const float half_wsc = g.wsc * 0.5f, half_hsc = g.hsc * 0.5f;

    if (tid < NBUF * kCtaRB) {
        s_ga[tid * P + g.Wo] = 0.f;
        s_gb[tid * P + g.Wo] = 0.f;
    }
    if (use_tma && tid == 0) {
        mbar_init(bar_U, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    unsigned parity_U = 0;

    for (long long b = blockIdx.x; b < a.Bsrc; b += gridDim.x) {
        const float* __restrict__ Ub = a.U + b * (long long)SC;
        float* __restrict__ dUb = a.dU ? a.dU + b * (long long)SC : nullptr;
        const float* __restrict__ gb = a.gout + b * (long long)g.N;
        Theta th;
        th.load(a.theta + 6 * b);
        float z = 1.0f;
        bool active = true;
        if (COMPOSITE) {
            z = __ldg(a.z_pres + b);
            active = a.stop_sum ? (__ldg(a.stop_sum + b) < a.threshold) : true;
        }
        // the image's dU, zero-filled by the whole CTA (rows the stream reaches overwrite it after the next barrier)
        if (dUb) {
            if ((reinterpret_cast<uintptr_t>(dUb) & 15) == 0 && (SC & 3) == 0) {
                float4* v = reinterpret_cast<float4*>(dUb);
                for (int k = tid; k < SC / 4; k += kCtaThreads) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int k = tid; k < SC; k += kCtaThreads) dUb[k] = 0.f;
            }
        }
        float p[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const bool sep = th.separable();
        if (active && !sep) {   // general affine theta: cold path, one warp on global memory
            __syncthreads();
            if (warp == 0)
                bwd_general_image<COMPOSITE>(Ub, dUb, gb, a.dtheta ? a.dtheta + 6 * b : nullptr, (COMPOSITE && a.dz) ? a.dz + b : nullptr,
                                             th.t[0], th.t[1], th.t[2], th.t[3], th.t[4], th.t[5], z, false, lane, g.Hs, g.Ws, 1, g.Ho,
                                             g.Wo, g.step_w, g.step_h, g.wsc, g.hsc);
            __syncthreads();
            continue;
        }
        if (active) {
            // ---- stage the source window ------------------------------------------------------------------------
            if (use_tma) {
                if (tid == 0) {   // (every thread is past the previous image's taps: barrier at the end of the image)
                    mbar_expect_tx(bar_U, (unsigned)(SC * 4));
                    tma_load_3d(s_U, &tmU, bar_U, 0, 0, (int)b);
                }
            } else {
                for (int k = tid; k < SC; k += kCtaThreads) s_U[k] = __ldg(Ub + k);
            }
            // ---- axis tables and in-range intervals -------------------------------------------------------------
            int ilo = g.Ho, ihi = -1, jlo = g.Wo, jhi = -1;
            const bool ascending = !(th.t[4] < 0.0f);   // stream order that makes y0 non-decreasing
            for (int i = tid; i < g.Ho; i += kCtaThreads) {
                const float yt = lin_at(i, g.step_h);
                const Axis Y = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, yt), g.hsc, g.Hs);
                // stream position of row i: in-range rows form one interval, so position = i - ilo (ascending) or ihi - i;
                // written at i (ascending) or Ho - 1 - i (descending) and addressed relative to the interval's first entry
                s_row[ascending ? i : g.Ho - 1 - i] = make_int4(Y.c0 * ws4, __float_as_int(yt), __float_as_int(Y.a), __float_as_int(Y.b));
                if (Y.c0 != Y.c1) { ilo = min(ilo, i); ihi = max(ihi, i); }
            }
            for (int j = tid; j < g.Wo; j += kCtaThreads) {
                const float xt = lin_at(j, g.step_w);
                const Axis X = axis_tap(affine_row(th.t[0], th.t[1], th.t[2], xt, 0.0f), g.wsc, g.Ws);
                s_col[j] = make_int4(X.c0 * 4, __float_as_int(xt), __float_as_int(X.a), __float_as_int(X.b));
                if (X.c0 != X.c1) { jlo = min(jlo, j); jhi = max(jhi, j); }
            }
            for (int x = tid; x < g.Ws; x += kCtaThreads) s_run[x] = 0;
            ilo = __reduce_min_sync(0xffffffffu, ilo); ihi = __reduce_max_sync(0xffffffffu, ihi);
            jlo = __reduce_min_sync(0xffffffffu, jlo); jhi = __reduce_max_sync(0xffffffffu, jhi);
            if (lane == 0) { s_redi[warp * 8 + 0] = ilo; s_redi[warp * 8 + 1] = ihi; s_redi[warp * 8 + 2] = jlo; s_redi[warp * 8 + 3] = jhi; }
            __syncthreads();   // tables, source copy (non-TMA), zero fill and the warps' intervals are in place
#pragma unroll
            for (int w = 0; w < kCtaWarps; ++w) {
                ilo = min(ilo, s_redi[w * 8 + 0]); ihi = max(ihi, s_redi[w * 8 + 1]);
                jlo = min(jlo, s_redi[w * 8 + 2]); jhi = max(jhi, s_redi[w * 8 + 3]);
            }
            const bool any = ihi >= ilo && jhi >= jlo;
            if (any) {
                // column runs: source column x receives output columns [start, end) = {j : x0[j] == x}
                int rmax = 0;
                for (int j = jlo + tid; j <= jhi; j += kCtaThreads) {
                    const int x0 = s_col[j].x;
                    if (j == jlo || s_col[j - 1].x != x0) {
                        int e = j + 1;
                        while (e <= jhi && s_col[e].x == x0) ++e;
                        s_run[x0 >> 2] = j | (e << 16);
                        rmax = max(rmax, e - j);
                    }
                }
                rmax = __reduce_max_sync(0xffffffffu, rmax);
                if (lane == 0) s_redi[warp * 8 + 4] = rmax;
            }
            if (use_tma) {
                mbar_wait(bar_U, parity_U);
                parity_U ^= 1u;
            }
            __syncthreads();   // runs and rmax
            if (any) {
                int rmax = 0;
#pragma unroll
                for (int w = 0; w < kCtaWarps; ++w) rmax = max(rmax, s_redi[w * 8 + 4]);
                const bool need_dU = dUb != nullptr;
                const bool need_taps = a.dtheta != nullptr || (COMPOSITE && a.dz != nullptr);
                const int nrows = ihi - ilo + 1;
                const int njc = (jhi - jlo + 32) >> 5;
                const int4* rows = s_row + (ascending ? ilo : g.Ho - 1 - ihi);   // rows[ii] = stream row ii
                const float* gfirst = gb + (long long)(ascending ? ilo : ihi) * g.Wo;
                const int gstep = ascending ? g.Wo : -g.Wo;
                const int xa = s_col[jlo].x >> 2, xb = s_col[jhi].x >> 2;
                const int xlo = min(xa, xb), nxc = (max(xa, xb) + 2 - xlo + 31) >> 5;   // source columns [xlo, max + 1]
                // phase-2 ownership: source chunk `warp` (nxc <= 3 for Ws <= 64)
                const int x = xlo + 32 * warp + lane;
                const bool own = warp < nxc;
                const bool xok = own && x < g.Ws;
                int a0 = 0, a1 = 0, b0 = 0, b1 = 0;
                if (xok) { const int ra = s_run[x]; a0 = ra & 0xffff; a1 = ra >> 16; }
                if (xok && x > 0) { const int rb = s_run[x - 1]; b0 = rb & 0xffff; b1 = rb >> 16; }
                float v0 = 0.f, v1 = 0.f;
                int oc = -1;   // byte offset of the source row held in v0 (v1: the next row); -1 = none
                char* colp = reinterpret_cast<char*>(dUb) + x * 4;
                const char* Usc = reinterpret_cast<const char*>(s_U);

                int buf = 0;
                for (int ii0 = 0; ii0 < nrows; ii0 += kCtaRB, buf ^= (NBUF - 1)) {
                    const int nb = min(kCtaRB, nrows - ii0);
                    float* ga_b = s_ga + buf * (kCtaRB * P);
                    float* gb_b = s_gb + buf * (kCtaRB * P);
                    // ---- phase 1: output chunks, handed out from the last warp down (the first warps own the source chunks) ----
                    for (int jc = kCtaWarps - 1 - warp; jc < njc; jc += kCtaWarps) {
                        const int j = jlo + 32 * jc + lane;
                        if (j > jhi) continue;
                        const int4 cj = s_col[j];
                        const float xt = __int_as_float(cj.y), ax = __int_as_float(cj.z), bx = __int_as_float(cj.w);
                        const float* gp = gfirst + (long long)ii0 * gstep + j;
                        float gq[kCtaRB];
#pragma unroll
                        for (int r = 0; r < kCtaRB; ++r) gq[r] = (r < nb) ? __ldg(gp + r * gstep) : 0.f;
                        const int4* rw = rows + ii0;
                        float* ga = ga_b + j;
                        float* gbuf = gb_b + j;
                        const char* Uc = Usc + cj.x;
                        float SX = 0.f, SY = 0.f;
                        float Ia = 0.f, Ib = 0.f, Ic = 0.f, Id = 0.f;
                        int poff = -1;
#pragma unroll
                        for (int r = 0; r < kCtaRB; ++r) {
                            const int4 cy = rw[r < nb ? r : nb - 1];   // (rows past the end: last valid row with g = 0)
                            const float yt = __int_as_float(cy.y), ay = __int_as_float(cy.z), by = __int_as_float(cy.w);
                            const float gv = COMPOSITE ? gq[r] * z : gq[r];
                            if (need_taps) {
                                if (cy.x != poff) {   // rows that share a source row share their taps (uniform branch)
                                    const char* pa = Uc + cy.x;
                                    Ia = *reinterpret_cast<const float*>(pa);        Ic = *reinterpret_cast<const float*>(pa + 4);
                                    Ib = *reinterpret_cast<const float*>(pa + ws4);  Id = *reinterpret_cast<const float*>(pa + ws4 + 4);
                                    poff = cy.x;
                                }
                                const float sx = gv * (ay * (Ic - Ia) + by * (Id - Ib));
                                const float sy = gv * (ax * (Ib - Ia) + bx * (Id - Ic));
                                SX += sx; SY += sy;
                                p[1] += sx * yt; p[4] += sy * yt;
                                if (COMPOSITE) p[6] += gq[r] * ((ax * ay) * Ia + (ax * by) * Ib + (bx * ay) * Ic + (bx * by) * Id);
                            }
                            if (need_dU) { ga[r * P] = ax * gv; gbuf[r * P] = bx * gv; }
                        }
                        p[0] += SX * xt; p[2] += SX; p[3] += SY * xt; p[5] += SY;
                    }
                    if (!need_dU) continue;
                    __syncthreads();   // the batch's gather rows are complete (and the other buffer's readers are done: they
                                       // run phase 2 of the previous batch before they get here)
                    // ---- phase 2: source chunk `warp`: T[x] over the runs, folded into the two running source rows ----
                    if (own) {
                        float T[kCtaRB];
#pragma unroll
                        for (int r = 0; r < kCtaRB; ++r) T[r] = 0.f;
                        for (int q = 0; q < rmax; ++q) {
                            const int ia = (a0 + q < a1) ? a0 + q : g.Wo;
                            const int ib = (b0 + q < b1) ? b0 + q : g.Wo;
#pragma unroll
                            for (int r = 0; r < kCtaRB; ++r) T[r] += ga_b[r * P + ia] + gb_b[r * P + ib];
                        }
#pragma unroll
                        for (int r = 0; r < kCtaRB; ++r) {
                            if (r < nb) {
                                const int4 cy = rows[ii0 + r];
                                const int off = cy.x;
                                if (off != oc) {
                                    if (oc >= 0) {
                                        emit_px(colp + oc, v0, xok, true);
                                        if (off == oc + ws4) {
                                            v0 = v1; v1 = 0.f;
                                        } else {
                                            emit_px(colp + oc + ws4, v1, xok, true);
                                            v0 = 0.f; v1 = 0.f;
                                        }
                                    }
                                    oc = off;
                                }
                                v0 += __int_as_float(cy.z) * T[r];
                                v1 += __int_as_float(cy.w) * T[r];
                            }
                        }
                    }
                    if (NBUF == 1) __syncthreads();   // single buffer: the gather rows are free again
                }
                if (need_dU && own && oc >= 0) {
                    emit_px(colp + oc, v0, xok, true);
                    emit_px(colp + oc + ws4, v1, xok, true);
                }
                p[0] *= half_wsc; p[1] *= half_wsc; p[2] *= half_wsc;
                p[3] *= half_hsc; p[4] *= half_hsc; p[5] *= half_hsc;
            }
        } else {
            __syncthreads();   // (keeps the barrier count per image uniform for the zero fill)
        }
        // ---- dtheta / dz: warp shuffles, then the block reduction ----
#pragma unroll
        for (int k = 0; k < 7; ++k) p[k] = warp_sum(p[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 7; ++k) s_redf[warp * 8 + k] = p[k];
        }
        __syncthreads();
        if (tid < 7) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kCtaWarps; ++w) s += s_redf[w * 8 + tid];
            if (tid < 6) {
                if (a.dtheta) a.dtheta[6 * b + tid] = s;
            } else if (COMPOSITE && a.dz) {
                a.dz[b] = s;
            }
        }
        __syncthreads();   // s_redf / s_redi / tables / s_U are free for the next image
    }
}

}  // namespace mog
