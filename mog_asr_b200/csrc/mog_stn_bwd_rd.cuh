// Backward of the sampler, read direction (large source canvas, glimpse <= 64 columns wide), dU + dtheta: one CTA per
// image with warp specialisation (sm_100a).
//
// dU of a read is canvas-sized and 80-95 % zeros; the warp-per-image kernel writes those zeros and does the glimpse's
// arithmetic in the same warp, one after the other (measured: fill alone 662 us, arithmetic alone 489 us, together
// 1072 us on 16 384 canvases of 256 x 256 -- no overlap).  Here the two jobs belong to different warps of one CTA:
//   * warps 2-3 (fill): zero the band of source rows the glimpse touches, meet the compute warps at one barrier, then
//     stream 16-byte zero stores over the rest of the image's dU; they never wait on a load;
//   * warps 0-1 (compute): axis tables, column runs, then batches of 4 in-range glimpse rows: phase 1 (warp = 32-column
//     chunk of the glimpse) g x taps -> dtheta partial sums and ax*g, bx*g into double-buffered gather rows; phase 2
//     (source chunks of 32 columns dealt out to the two warps) T[x] over the runs, folded into two running source rows
//     held in registers and stored once, coalesced, over the zeroed band.  They synchronise among themselves on a named
//     barrier (64 threads) that the fill warps never touch.
// Same arithmetic and order along a row as stn_bwd_warp_kernel; no atomics, dU deterministic.
#pragma once
#include "mog_stn_bwd_cta.cuh"

namespace mog {

constexpr int kRdRB = 4;          // glimpse rows per batch
constexpr int kRdNXCW = 5;        // source chunks (of 32 columns) per compute warp: footprints up to 320 columns
#ifndef MOG_BWD_RD_MINB
#define MOG_BWD_RD_MINB 4
#endif

struct RdLayout {
    int row, col, run, ga, gb, red, total;   // byte offsets
};
__host__ __device__ inline RdLayout bwd_rd_layout(const Geo& g) {
    RdLayout l;
    int o = 0;
    l.row = o; o += g.Ho * 16;
    l.col = o; o += g.Wo * 16;
    l.run = o; o += align128(g.Ws * 4);
    l.ga = o;  o += align128(2 * kRdRB * (g.Wo + 1) * 4);
    l.gb = o;  o += align128(2 * kRdRB * (g.Wo + 1) * 4);
    l.red = o; o += 256;
    l.total = o;
    return l;
}

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 64;" ::: "memory"); }    // the two compute warps
__device__ __forceinline__ void bar_band() { asm volatile("bar.sync 2, 128;" ::: "memory"); }      // S1: fill and compute warps, once per image

// zero p[0, n) with `nt` threads (t = this thread's index among them): 16-byte stores between the 16-byte boundaries
__device__ __forceinline__ void fill_zero_group(float* __restrict__ p, long long n, int t, int nt) {
    if (n <= 0) return;
    const int mis = (int)((reinterpret_cast<uintptr_t>(p) >> 2) & 3);
    const long long head = min(n, (long long)((4 - mis) & 3));
    const long long nv = (n - head) >> 2;
    const long long tail = head + 4 * nv;
    if ((long long)t < head) p[t] = 0.0f;
    float4* v = reinterpret_cast<float4*>(p + head);
    for (long long k = t; k < nv; k += nt) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tail + t < n) p[tail + t] = 0.0f;
}

__global__ void __launch_bounds__(kCtaThreads, MOG_BWD_RD_MINB) stn_bwd_rd_kernel(const BwdArgs a) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const Geo& g = a.g;
    const RdLayout L = bwd_rd_layout(g);
    int4* s_row = reinterpret_cast<int4*>(s_raw + L.row);    // stream order: {y0 * Ws * 4, y_t, ay, by}
    int4* s_col = reinterpret_cast<int4*>(s_raw + L.col);    // {x0 * 4, x_t, ax, bx}
    int* s_run = reinterpret_cast<int*>(s_raw + L.run);
    float* s_ga = reinterpret_cast<float*>(s_raw + L.ga);    // [2][kRdRB][Wo + 1]
    float* s_gb = reinterpret_cast<float*>(s_raw + L.gb);
    int* s_redi = reinterpret_cast<int*>(s_raw + L.red);
    float* s_redf = reinterpret_cast<float*>(s_raw + L.red + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool filler = warp >= 2;
    const int P = g.Wo + 1;
    const int ws4 = g.Ws * 4;
    const int SC = g.S;
    const float half_wsc = g.wsc * 0.5f, half_hsc = g.hsc * 0.5f;
    if (tid < 2 * kRdRB) {
        s_ga[tid * P + g.Wo] = 0.f;
        s_gb[tid * P + g.Wo] = 0.f;
    }
    __syncthreads();

    for (long long b = blockIdx.x; b < a.Bsrc; b += gridDim.x) {
        const float* __restrict__ Ub = a.U + b * (long long)SC;
        float* __restrict__ dUb = a.dU + b * (long long)SC;
        const float* __restrict__ gb = a.gout + b * (long long)g.N;
        Theta th;
        th.load(a.theta + 6 * b);
        if (!th.separable()) {   // general affine theta: cold path (whole-CTA zero fill, then one warp on global memory)
            fill_zero_group(dUb, SC, tid, kCtaThreads);
            __syncthreads();
            if (warp == 0)
                bwd_general_image<false>(Ub, dUb, gb, a.dtheta ? a.dtheta + 6 * b : nullptr, nullptr, th.t[0], th.t[1], th.t[2], th.t[3],
                                         th.t[4], th.t[5], 1.0f, false, lane, g.Hs, g.Ws, 1, g.Ho, g.Wo, g.step_w, g.step_h, g.wsc, g.hsc);
            __syncthreads();
            continue;
        }
        // ---- in-range intervals (every warp for itself: a dozen flops per row / column); the compute warps keep the tables ----
        const bool ascending = !(th.t[4] < 0.0f);
        int ilo = g.Ho, ihi = -1, jlo = g.Wo, jhi = -1;
        for (int i = lane; i < g.Ho; i += 32) {
            const float yt = lin_at(i, g.step_h);
            const Axis Y = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, yt), g.hsc, g.Hs);
            if (warp == 0) s_row[ascending ? i : g.Ho - 1 - i] = make_int4(Y.c0 * ws4, __float_as_int(yt), __float_as_int(Y.a), __float_as_int(Y.b));
            if (Y.c0 != Y.c1) { ilo = min(ilo, i); ihi = max(ihi, i); }
        }
        for (int j = lane; j < g.Wo; j += 32) {
            const float xt = lin_at(j, g.step_w);
            const Axis X = axis_tap(affine_row(th.t[0], th.t[1], th.t[2], xt, 0.0f), g.wsc, g.Ws);
            if (warp == 1) s_col[j] = make_int4(X.c0 * 4, __float_as_int(xt), __float_as_int(X.a), __float_as_int(X.b));
            if (X.c0 != X.c1) { jlo = min(jlo, j); jhi = max(jhi, j); }
        }
        ilo = __reduce_min_sync(0xffffffffu, ilo); ihi = __reduce_max_sync(0xffffffffu, ihi);
        jlo = __reduce_min_sync(0xffffffffu, jlo); jhi = __reduce_max_sync(0xffffffffu, jhi);
        const bool any = ihi >= ilo && jhi >= jlo;
        int ylo = 0, yend = 0;   // band of source rows the stream reaches: [ylo, yend)
        if (any) {
            const int ya = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, lin_at(ilo, g.step_h)), g.hsc, g.Hs).c0;
            const int yb = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, lin_at(ihi, g.step_h)), g.hsc, g.Hs).c0;
            ylo = min(ya, yb); yend = max(ya, yb) + 2;
        }

        if (filler) {
            // ================= fill warps =================
            const int ft = tid - 64;
            fill_zero_group(dUb + (long long)ylo * g.Ws, (long long)(yend - ylo) * g.Ws, ft, 64);   // the band first
            bar_band();                                                                          // S1: the compute warps may store into the band
            fill_zero_group(dUb, (long long)ylo * g.Ws, ft, 64);
            fill_zero_group(dUb + (long long)yend * g.Ws, (long long)(SC - yend * g.Ws), ft, 64);
            continue;
        }
        // ================= compute warps =================
        float p[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int x = tid; x < g.Ws; x += 64) s_run[x] = 0;
        bar_compute();   // tables and the cleared run table
        int rmax = 0;
        if (any) {
            for (int j = jlo + tid; j <= jhi; j += 64) {
                const int x0 = s_col[j].x;
                if (j == jlo || s_col[j - 1].x != x0) {
                    int e = j + 1;
                    while (e <= jhi && s_col[e].x == x0) ++e;
                    s_run[x0 >> 2] = j | (e << 16);
                    rmax = max(rmax, e - j);
                }
            }
            rmax = __reduce_max_sync(0xffffffffu, rmax);
            if (lane == 0) s_redi[warp] = rmax;
        }
        bar_compute();   // runs
        bar_band();      // S1: the band is zeroed
        if (any) {
            rmax = max(s_redi[0], s_redi[1]);
            const bool need_taps = a.dtheta != nullptr;
            const int nrows = ihi - ilo + 1;
            const int njc = (jhi - jlo + 32) >> 5;
            const int4* rows = s_row + (ascending ? ilo : g.Ho - 1 - ihi);
            const float* gfirst = gb + (long long)(ascending ? ilo : ihi) * g.Wo;
            const int gstep = ascending ? g.Wo : -g.Wo;
            const int xa = s_col[jlo].x >> 2, xb = s_col[jhi].x >> 2;
            const int xlo = min(xa, xb), nxc = (max(xa, xb) + 2 - xlo + 31) >> 5;   // source columns [xlo, max + 1]
            // phase-2 ownership: source chunks warp, warp + 2, ...
            int ra_[kRdNXCW], rb_[kRdNXCW], oc[kRdNXCW];
            float v0[kRdNXCW], v1[kRdNXCW];
#pragma unroll
            for (int c = 0; c < kRdNXCW; ++c) {
                const int x = xlo + 32 * (warp + 2 * c) + lane;
                const bool ok = (warp + 2 * c) < nxc && x < g.Ws;
                ra_[c] = ok ? s_run[x] : 0;
                rb_[c] = (ok && x > 0) ? s_run[x - 1] : 0;
                v0[c] = v1[c] = 0.f;
                oc[c] = -1;
            }
            const char* Ubc = opaque(reinterpret_cast<const char*>(Ub));
            char* dUbc = reinterpret_cast<char*>(dUb) + (xlo + 32 * warp + lane) * 4;

            int buf = 0;
            for (int ii0 = 0; ii0 < nrows; ii0 += kRdRB, buf ^= 1) {
                const int nb = min(kRdRB, nrows - ii0);
                float* ga_b = s_ga + buf * (kRdRB * P);
                float* gb_b = s_gb + buf * (kRdRB * P);
                // ---- phase 1: glimpse chunk = warp (wider glimpses: chunks warp, warp + 2, ...) ----
                for (int jc = warp; jc < njc; jc += 2) {
                    const int j = jlo + 32 * jc + lane;
                    if (j > jhi) continue;
                    const int4 cj = s_col[j];
                    const float xt = __int_as_float(cj.y), ax = __int_as_float(cj.z), bx = __int_as_float(cj.w);
                    const float* gp = gfirst + (long long)ii0 * gstep + j;
                    float gq[kRdRB], I[kRdRB][4];
                    int4 cy[kRdRB];
#pragma unroll
                    for (int r = 0; r < kRdRB; ++r) {
                        cy[r] = rows[ii0 + (r < nb ? r : nb - 1)];
                        gq[r] = (r < nb) ? __ldg(gp + r * gstep) : 0.f;
                        if (need_taps) {
                            const char* pa = Ubc + (unsigned)(cy[r].x + cj.x);
                            I[r][0] = ldg_f32(pa);        I[r][2] = ldg_f32(pa + 4);
                            I[r][1] = ldg_f32(pa + ws4);  I[r][3] = ldg_f32(pa + ws4 + 4);
                        }
                    }
                    float SX = 0.f, SY = 0.f;
#pragma unroll
                    for (int r = 0; r < kRdRB; ++r) {
                        const float yt = __int_as_float(cy[r].y), ay = __int_as_float(cy[r].z), by = __int_as_float(cy[r].w);
                        const float gv = gq[r];
                        if (need_taps) {
                            const float sx = gv * (ay * (I[r][2] - I[r][0]) + by * (I[r][3] - I[r][1]));
                            const float sy = gv * (ax * (I[r][1] - I[r][0]) + bx * (I[r][3] - I[r][2]));
                            SX += sx; SY += sy;
                            p[1] += sx * yt; p[4] += sy * yt;
                        }
                        ga_b[r * P + j] = ax * gv;
                        gb_b[r * P + j] = bx * gv;
                    }
                    p[0] += SX * xt; p[2] += SX; p[3] += SY * xt; p[5] += SY;
                }
                bar_compute();   // the batch's gather rows are complete; the other buffer's readers are done
                // ---- phase 2: this warp's source chunks ----
#pragma unroll
                for (int c = 0; c < kRdNXCW; ++c) {
                    if (warp + 2 * c < nxc) {
                        const int a0 = ra_[c] & 0xffff, a1 = ra_[c] >> 16, b0 = rb_[c] & 0xffff, b1 = rb_[c] >> 16;
                        float T[kRdRB];
#pragma unroll
                        for (int r = 0; r < kRdRB; ++r) T[r] = 0.f;
                        for (int q = 0; q < rmax; ++q) {
                            const int ia = (a0 + q < a1) ? a0 + q : g.Wo;
                            const int ib = (b0 + q < b1) ? b0 + q : g.Wo;
#pragma unroll
                            for (int r = 0; r < kRdRB; ++r) T[r] += ga_b[r * P + ia] + gb_b[r * P + ib];
                        }
                        const bool xok = xlo + 32 * (warp + 2 * c) + lane < g.Ws;
                        char* colp = dUbc + c * 256;
                        int o = oc[c];
                        float w0 = v0[c], w1 = v1[c];
#pragma unroll
                        for (int r = 0; r < kRdRB; ++r) {
                            if (r < nb) {
                                const int4 cyr = rows[ii0 + r];
                                const int off = cyr.x;
                                if (off != o) {
                                    if (o >= 0) {
                                        emit_px(colp + o, w0, xok, true);
                                        if (off == o + ws4) {
                                            w0 = w1; w1 = 0.f;
                                        } else {
                                            emit_px(colp + o + ws4, w1, xok, true);
                                            w0 = 0.f; w1 = 0.f;
                                        }
                                    }
                                    o = off;
                                }
                                w0 += __int_as_float(cyr.z) * T[r];
                                w1 += __int_as_float(cyr.w) * T[r];
                            }
                        }
                        oc[c] = o; v0[c] = w0; v1[c] = w1;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kRdNXCW; ++c) {
                if (warp + 2 * c < nxc && oc[c] >= 0) {
                    const bool xok = xlo + 32 * (warp + 2 * c) + lane < g.Ws;
                    emit_px(dUbc + c * 256 + oc[c], v0[c], xok, true);
                    emit_px(dUbc + c * 256 + oc[c] + ws4, v1[c], xok, true);
                }
            }
            p[0] *= half_wsc; p[1] *= half_wsc; p[2] *= half_wsc;
            p[3] *= half_hsc; p[4] *= half_hsc; p[5] *= half_hsc;
        }
        // ---- dtheta: warp shuffles, then the reduction over the two compute warps ----
#pragma unroll
        for (int k = 0; k < 6; ++k) p[k] = warp_sum(p[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) s_redf[warp * 8 + k] = p[k];
        }
        bar_compute();
        if (tid < 6 && a.dtheta) a.dtheta[6 * b + tid] = s_redf[tid] + s_redf[8 + tid];
        bar_compute();   // tables, run table, gather rows and the reduction scratch are free for the next image
    }
}

}  // namespace mog
