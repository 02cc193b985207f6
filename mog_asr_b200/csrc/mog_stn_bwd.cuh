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
//   3. forms the gradient row of source row y as V = A (this group) + Bv (previous group) and turns it into
//      source columns: ax*V and bx*V are summed over the column runs (run(x) = the consecutive output
//      columns whose left tap is x) with warp shuffles; the first lane of every run (its head) fetches the
//      right-tap sum of the neighbouring run with one more shuffle and stores T[x] with a plain store into
//      a shared-memory row indexed by source column -- no read-modify-write, no atomics; the rows are then
//      written out coalesced, each dU row once.
// Wide outputs are processed in strips of 64 output columns; the one or two source columns two adjacent
// strips share are handed over through a small shared-memory edge array instead of being accumulated in
// global memory.  Everything is deterministic.
#pragma once
#include "mog_stn_warp.cuh"

namespace mog {

#ifndef MOG_BWD2_RB
#define MOG_BWD2_RB 4
#endif
#ifndef MOG_BWD2_MINB
#define MOG_BWD2_MINB 10
#endif

constexpr int kBwdSlots = 2 * MOG_BWD2_RB;     // gradient rows a batch of MOG_BWD2_RB output rows can finish (a row may also flush a carry row)
constexpr int kBwdSW = 64;                     // strip width (two 32-lane chunks)

// per-warp shared memory of the emission machinery in 4-byte words: pending rows by output column (kBwdSlots x 64) |
// their source-row byte offsets (kBwdSlots) | rows by source column (kBwdSlots x (Ws + 1)) | two edge arrays (Hs x 2 each)
__host__ __device__ inline int bwd_emit_smem_words(const Geo& g) { return kBwdSlots * kBwdSW + kBwdSlots + kBwdSlots * (g.Ws + 1) + 4 * g.Hs; }
// grouped kernel with register loads: row table 4*Ho ({y0 << 1 | last-of-group, ay, by, yt} per in-range row, in
// stream order) in front of the emission area
__host__ __device__ inline int bwd2_warp_smem_words(const Geo& g) { return (4 * g.Ho + bwd_emit_smem_words(g) + 3) & ~3; }

struct RowP {      // one output row of a separable theta
    int y;         // source row y0 (clipped)
    int yoff;      // its byte offset
    bool in;       // y0 != y1: the row is inside the source range
    float ay, by, yt;
};
__device__ __forceinline__ RowP row_params(const Theta& th, const Geo& g, int i, int ws4) {
    RowP r;
    r.yt = lin_at(i, g.step_h);
    const Axis Y = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, r.yt), g.hsc, g.Hs);
    r.y = Y.c0;
    r.yoff = Y.c0 * ws4;
    r.in = Y.c0 != Y.c1;
    r.ay = Y.a;
    r.by = Y.b;
    return r;
}

// ---- per-strip state ------------------------------------------------------------------------------------------
struct EmitSmem {
    float* v;      // [kBwdSlots][64] pending rows by output column
    int* sloty;    // [kBwdSlots] source row of every pending row
    float* x;      // [kBwdSlots][Ws + 1] rows by source column (kept all-zero between emissions)
    float* edge;   // [2][Hs][2] partial sums of the source columns shared by two adjacent strips
};
__device__ __forceinline__ EmitSmem emit_smem(int* base, const Geo& g) {
    EmitSmem e;
    e.v = reinterpret_cast<float*>(base);
    e.sloty = base + kBwdSlots * kBwdSW;
    e.x = reinterpret_cast<float*>(e.sloty + kBwdSlots);
    e.edge = e.x + kBwdSlots * (g.Ws + 1);
    return e;
}

// seg[c] packs, per lane and chunk: bits 0-5 number of lanes after this one in the same run; bit 8 head of a run;
// bits 9-13 lane of the head whose source column is x - 1 (bit 14: there is one); bit 15 no head has column x + 1
// (this head stores its right-tap sum itself); bit 16 one of the first two heads of a chunk other than the first (their
// columns may hold what the previous chunk stored -- the other part of a run split by the chunk boundary, or its right-tap
// sum: they accumulate instead of storing; the rows are all-zero otherwise)
template <int NJC>
struct Strip {
    int xo[NJC], seg[NJC];
    float caz[NJC], cbz[NJC], cax[NJC], cbx[NJC];
    bool val[NJC];
    int rmax;                  // longest run inside a chunk
    int xlo, nxs;              // source columns [xlo, xlo + 32 nxs) cover the strip's taps
    int in_lo, in_n;           // columns shared with the previous strip (their partial sums arrive through the edge array)
    int ov_lo, ov_n;           // columns shared with the next strip (handed over, not stored)
    int parity;                // edge array written by this strip
};

// columns [jfirst, je) of the strip are valid; lanes outside shadow a valid column with g = 0
template <int NJC, bool COMPOSITE>
__device__ __forceinline__ void strip_setup(Strip<NJC>& s, const Theta& th, const Geo& g, int js, int jfirst, int je, int jlo, int jhi,
                                            float z, int lane, bool need_dU, int strip_index) {
    s.rmax = 1;
#pragma unroll
    for (int c = 0; c < NJC; ++c) {
        const int j = js + 32 * c + lane;
        s.val[c] = j >= jfirst && j < je;
        const Axis X = col_axis(th, g, min(max(j, jfirst), je - 1));
        s.xo[c] = X.c0 * 4;
        s.cax[c] = X.a; s.cbx[c] = X.b;
        s.caz[c] = COMPOSITE ? X.a * z : X.a;
        s.cbz[c] = COMPOSITE ? X.b * z : X.b;
        // run structure inside the chunk (masked lanes belong to no run)
        const int xprev = __shfl_up_sync(0xffffffffu, s.xo[c], 1);
        const int vprev = __shfl_up_sync(0xffffffffu, (int)s.val[c], 1);
        const bool cont = lane > 0 && s.val[c] && vprev != 0 && xprev == s.xo[c];   // continues lane - 1's run
        const bool head = s.val[c] && !cont;
        const unsigned cmask = __ballot_sync(0xffffffffu, cont);
        const unsigned hmask = __ballot_sync(0xffffffffu, head);
        const unsigned above = lane == 31 ? 0u : (cmask >> (lane + 1));
        const int follow = __ffs(~above) - 1;
        // neighbouring heads in lane order
        const unsigned below_h = hmask & ((1u << lane) - 1u);
        const unsigned above_h = lane == 31 ? 0u : (hmask >> (lane + 1));
        const int prev_h = below_h ? 31 - __clz(below_h) : -1;
        const int next_h = above_h ? lane + __ffs(above_h) : -1;
        const int xp = __shfl_sync(0xffffffffu, s.xo[c], prev_h < 0 ? lane : prev_h);
        const int xn = __shfl_sync(0xffffffffu, s.xo[c], next_h < 0 ? lane : next_h);
        int donor = -1;          // head whose source column is x - 1
        bool taker = false;      // some head has column x + 1 and will fetch this head's right-tap sum
        if (prev_h >= 0 && xp == s.xo[c] - 4) donor = prev_h;
        if (next_h >= 0 && xn == s.xo[c] - 4) donor = next_h;
        if (prev_h >= 0 && xp == s.xo[c] + 4) taker = true;
        if (next_h >= 0 && xn == s.xo[c] + 4) taker = true;
        int pk = follow;
        if (head) {
            pk |= 256;
            if (donor >= 0) pk |= (donor << 9) | (1 << 14);
            if (!taker) pk |= 1 << 15;
            if (c > 0 && __popc(below_h) <= 1) pk |= 1 << 16;
        }
        s.seg[c] = pk;
        s.rmax = max(s.rmax, follow + 1);
    }
    s.rmax = __reduce_max_sync(0xffffffffu, s.rmax);
    s.xlo = 0; s.nxs = 0; s.in_lo = s.in_n = s.ov_lo = s.ov_n = 0;
    s.parity = strip_index & 1;
    if (need_dU) {
        auto range = [&](int ja, int jb, int& lo, int& hi) {   // source columns touched by output columns [ja, jb]
            const int xa = col_axis(th, g, ja).c0, xb = col_axis(th, g, jb).c0;
            lo = min(xa, xb);
            hi = max(xa, xb) + 1;
        };
        int lo, hi;
        range(jfirst, je - 1, lo, hi);
        s.xlo = lo;
        s.nxs = (hi - lo + 32) >> 5;
        if (jfirst > jlo) {   // previous strip: columns [max(js - 32 NJC, jlo), jfirst - 1]
            int plo, phi;
            range(max(js - 32 * NJC, jlo), jfirst - 1, plo, phi);
            s.in_lo = max(lo, plo);
            s.in_n = max(0, min(hi, phi) - s.in_lo + 1);
        }
        if (je <= jhi) {      // next strip: columns [je, min(je + 32 NJC - 1, jhi)]
            int nlo, nhi;
            range(je, min(je + 32 * NJC - 1, jhi), nlo, nhi);
            s.ov_lo = max(lo, nlo);
            s.ov_n = max(0, min(hi, nhi) - s.ov_lo + 1);
        }
    }
}

// Turn the pending rows (slots [0, nslots) of e.v, by output column) into dU rows.
template <int NJC>
__device__ __forceinline__ void emit_slots(const Strip<NJC>& s, const EmitSmem& e, const Geo& g, char* dUbc, int nslots, int lane,
                                           bool first_write) {
    const int XP = g.Ws + 1;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NJC; ++c) {
        const int pk = s.seg[c];
        const int follow = pk & 63;
        const bool head = (pk & 256) != 0;
        const int donor = (pk >> 9) & 31;
        const int xi = s.xo[c] >> 2;
#pragma unroll 2
        for (int k = 0; k < nslots; ++k) {
            const float v = e.v[k * kBwdSW + 32 * c + lane];
            const float va0 = s.caz[c] * v, vb0 = s.cbz[c] * v;
            float va = va0, vb = vb0;
#pragma unroll 1
            for (int d = 1; d < s.rmax; ++d) {   // run sums (the lanes' own values travel, not their partial sums)
                const float ua = __shfl_down_sync(0xffffffffu, va0, d), ub = __shfl_down_sync(0xffffffffu, vb0, d);
                if (d <= follow) { va += ua; vb += ub; }
            }
            const float sbd = __shfl_sync(0xffffffffu, vb, (pk & (1 << 14)) ? donor : lane);
            if (head) {
                float* px = e.x + k * XP + xi;
                const float T = (pk & (1 << 14)) ? va + sbd : va;
                if (pk & (1 << 16)) {          // first head of a later chunk: the previous chunk may have stored here
                    px[0] += T;
                    if (pk & (1 << 15)) px[1] += vb;
                } else {
                    px[0] = T;
                    if (pk & (1 << 15)) px[1] = vb;
                }
            }
        }
        __syncwarp();
    }
    float* edge_out = e.edge + s.parity * 2 * g.Hs;
    const float* edge_in = e.edge + (s.parity ^ 1) * 2 * g.Hs;
#pragma unroll 1
    for (int cx = 0; cx < s.nxs; ++cx) {
        const int x = s.xlo + 32 * cx + lane;
        if (x < g.Ws) {
            const bool incoming = x >= s.in_lo && x < s.in_lo + s.in_n;
            const bool outgoing = x >= s.ov_lo && x < s.ov_lo + s.ov_n;
#pragma unroll 2
            for (int k = 0; k < nslots; ++k) {
                const int yrow = e.sloty[k];
                float T = e.x[k * XP + x];
                e.x[k * XP + x] = 0.f;
                if (incoming) T += edge_in[2 * yrow + (x - s.in_lo)];
                if (outgoing) edge_out[2 * yrow + (x - s.ov_lo)] = T;
                else emit_px(dUbc + (yrow * g.Ws + x) * 4, T, true, first_write);
            }
        }
    }
    __syncwarp();
}

// One batch of RB output rows whose g values (gq) and taps (I, for rows that end a group) are in registers:
// accumulate, finish groups (dtheta sums, pending dU rows).  Returns the number of pending rows.
template <int NJC, int RB, bool COMPOSITE>
__device__ __forceinline__ int batch_arith(const Strip<NJC>& s, const EmitSmem& e, const int4* rows, int nb, const int (&ey)[RB],
                                           const float (&gq)[NJC][RB], const float (&I)[NJC][RB][4], float (&A)[NJC], float (&Bv)[NJC],
                                           float (&A2)[NJC], float (&B2)[NJC], float (&car)[NJC], float (&SX)[NJC], float (&SXY)[NJC],
                                           float (&SY)[NJC], float (&SYY)[NJC], float (&SZ)[NJC], int& ycar, bool need_taps,
                                           bool need_dU, int lane) {
    int nslots = 0;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        if (r < nb) {
            const int4 er = rows[r];
            const float ay = __int_as_float(er.y), by = __int_as_float(er.z);
            const float ayt = ay * __int_as_float(er.w), byt = by * __int_as_float(er.w);
#pragma unroll
            for (int c = 0; c < NJC; ++c) {
                const float gv = s.val[c] ? gq[c][r] : 0.f;
                A[c] = fmaf(ay, gv, A[c]);   Bv[c] = fmaf(by, gv, Bv[c]);
                A2[c] = fmaf(ayt, gv, A2[c]); B2[c] = fmaf(byt, gv, B2[c]);
            }
            if (ey[r] & 1) {
                const int yrow = ey[r] >> 1;
                if (need_taps) {
#pragma unroll
                    for (int c = 0; c < NJC; ++c) {
                        const float Ia = I[c][r][0], Ib = I[c][r][1], Ic = I[c][r][2], Id = I[c][r][3];
                        const float dxa = Ic - Ia, dxb = Id - Ib;
                        SX[c] = fmaf(dxa, A[c], fmaf(dxb, Bv[c], SX[c]));
                        SXY[c] = fmaf(dxa, A2[c], fmaf(dxb, B2[c], SXY[c]));
                        const float E = fmaf(s.cax[c], Ib - Ia, s.cbx[c] * (Id - Ic));
                        SY[c] = fmaf(E, A[c] + Bv[c], SY[c]);
                        SYY[c] = fmaf(E, A2[c] + B2[c], SYY[c]);
                        if (COMPOSITE)   // dz = sum g * sample  (:722-727)
                            SZ[c] = fmaf(fmaf(s.cax[c], Ia, s.cbx[c] * Ic), A[c], fmaf(fmaf(s.cax[c], Ib, s.cbx[c] * Id), Bv[c], SZ[c]));
                    }
                }
                if (need_dU) {
                    if (ycar >= 0 && ycar != yrow) {   // the previous group's lower row stands alone
                        if (lane == 0) e.sloty[nslots] = ycar;
#pragma unroll
                        for (int c = 0; c < NJC; ++c) { e.v[nslots * kBwdSW + 32 * c + lane] = car[c]; car[c] = 0.f; }
                        ++nslots;
                    }
                    if (lane == 0) e.sloty[nslots] = yrow;
#pragma unroll
                    for (int c = 0; c < NJC; ++c) { e.v[nslots * kBwdSW + 32 * c + lane] = A[c] + car[c]; car[c] = Bv[c]; }
                    ++nslots;
                    ycar = yrow + 1;
                }
#pragma unroll
                for (int c = 0; c < NJC; ++c) A[c] = Bv[c] = A2[c] = B2[c] = 0.f;
            }
        }
    }
    return nslots;
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
    const EmitSmem em = emit_smem(s_base + 4 * g.Ho, g);
    for (int k = lane; k < kBwdSlots * (g.Ws + 1); k += 32) em.x[k] = 0.f;
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
                        if (!last) last = row_params(th, g, ascending ? ilo + ii + 1 : ihi - ii - 1, ws4).y != r.y;
                        s_row[ii] = make_int4((r.y << 1) | (last ? 1 : 0), __float_as_int(r.ay), __float_as_int(r.by), __float_as_int(r.yt));
                    }
                    __syncwarp();
                    const char* Ubc = opaque(reinterpret_cast<const char*>(Ub));
                    char* dUbc = reinterpret_cast<char*>(dUb);
                    const int gstep = ascending ? g.Wo : -g.Wo;
                    const float* gfirst = gb + (long long)(ascending ? ilo : ihi) * g.Wo;

                    // ---- strips of SW output columns; every strip streams the in-range rows once ----
                    int strip_index = 0;
                    for (int js = jlo; js <= jhi; js += SW, ++strip_index) {
                        const int je = min(js + SW, jhi + 1);   // strip = columns [js, je)
                        Strip<NJC> sp;
                        strip_setup<NJC, COMPOSITE>(sp, th, g, js, js, je, jlo, jhi, z, lane, need_dU, strip_index);
                        const float* gcol[NJC];
#pragma unroll
                        for (int c = 0; c < NJC; ++c) gcol[c] = gfirst + min(js + 32 * c + lane, je - 1);
                        float A[NJC], Bv[NJC], A2[NJC], B2[NJC], car[NJC], SX[NJC], SXY[NJC], SY[NJC], SYY[NJC], SZ[NJC];
#pragma unroll
                        for (int c = 0; c < NJC; ++c)
                            A[c] = Bv[c] = A2[c] = B2[c] = car[c] = SX[c] = SXY[c] = SY[c] = SYY[c] = SZ[c] = 0.f;
                        int ycar = -1;   // source row the carry belongs to (-1: none pending)

                        bool done = false;
                        for (int ii0 = 0; !done; ii0 += kRB2) {
                            int nslots = 0;
                            if (ii0 < nrows) {
                                const int nb = min(kRB2, nrows - ii0);
                                // ---- loads of the batch: g of every row, the four taps at the end of every group ----
                                int ey[kRB2];   // source row << 1 | last-of-group (tail rows shadow the last valid row, never 'last')
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
                                            const char* pa = Ubc + (unsigned)((ey[r] >> 1) * ws4 + sp.xo[c]);
                                            I[c][r][0] = ldg_f32(pa);        I[c][r][2] = ldg_f32(pa + 4);
                                            I[c][r][1] = ldg_f32(pa + ws4);  I[c][r][3] = ldg_f32(pa + ws4 + 4);
                                        }
                                    }
                                }
                                nslots = batch_arith<NJC, kRB2, COMPOSITE>(sp, em, s_row + ii0, nb, ey, gq, I, A, Bv, A2, B2, car, SX, SXY, SY, SYY,
                                                                           SZ, ycar, need_taps, need_dU, lane);
                            } else {
                                done = true;
                                if (need_dU && ycar >= 0) {   // lower row of the last group
                                    if (lane == 0) em.sloty[0] = ycar;
#pragma unroll
                                    for (int c = 0; c < NJC; ++c) em.v[32 * c + lane] = car[c];
                                    nslots = 1;
                                }
                            }
                            if (nslots > 0) emit_slots<NJC>(sp, em, g, dUbc, nslots, lane, first_write);
                        }
#pragma unroll
                        for (int c = 0; c < NJC; ++c) {
                            if (sp.val[c]) {
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
