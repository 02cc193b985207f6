// Backward of the sampler, write direction (small source, large output), staged through shared memory by the
// tensor memory accelerator (sm_100a).
//
// The write call site (air_number_bbox_location.py:592-600) samples a small window (28x28 ... 64x64) onto a large
// canvas; its backward reads the in-range box of the canvas gradient (up to the whole canvas) and the window.  With
// register loads that stream is a chain of dependent load rounds (a few rows per round), i.e. latency-bound.  Here
//   * the source window U[b] is fetched whole by ONE cp.async.bulk.tensor (3-D tensor map over [B][Hs][Ws], box = the
//     image) into the warp's shared memory and every tap is an LDS;
//   * the in-range box of g is cut into tiles of 8 rows x 64 columns at coordinates ((jlo & ~3) + 64 s, ilo + 8 t, b) of a
//     3-D tensor map over [B][Ho][Wo] (the innermost coordinate must be a multiple of 16 bytes -- measured: anything else
//     faults; rows and out-of-bounds parts are unrestricted); the tiles stream through a 4-stage ring, one
//     mbarrier per stage (expect_tx = tile bytes; out-of-bounds parts arrive as zeros), three tiles in flight
//     while the fourth is consumed.  No registers are held by loads, no address arithmetic per element.
// The arithmetic is the grouped gather form of mog_stn_bwd.cuh (same sums, same order): rows that share a source
// row accumulate first, taps / dtheta / the dU row are formed once per source row.
// One warp owns one image; all waits are on the warp's own barriers (bounded spin, then trap: a lost
// transaction must fail the launch, not hang the GPU).
#pragma once
#include <cuda.h>

#include "mog_stn_bwd.cuh"

namespace mog {

constexpr int kTmaTR = 8;                   // rows of a g tile
constexpr int kTmaSW = 64;                  // columns of a g tile (one strip: two 32-lane chunks)
constexpr int kTmaStages = 4;
constexpr int kTmaRB = MOG_BWD2_RB;         // rows per arithmetic sub-batch
constexpr int kTmaTileBytes = kTmaTR * kTmaSW * 4;
constexpr int kTmaMaxSrc = 4096;            // source pixels staged whole (64 x 64)
#ifndef MOG_BWD_TMA_MINB
#define MOG_BWD_TMA_MINB 4
#endif

__host__ __device__ inline int align128(int v) { return (v + 127) & ~127; }
// per-warp shared memory in bytes: source image | tile ring | tile row table | mbarriers | emission area (mog_stn_bwd.cuh)
__host__ __device__ inline int bwd_tma_warp_smem_bytes(const Geo& g) {
    return align128(g.S * 4) + kTmaStages * kTmaTileBytes + 128 + 128 + align128(bwd_emit_smem_words(g) * 4);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded: a transaction that never completes (a bug) traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    for (int it = 0; it < (1 << 24); ++it)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

template <bool COMPOSITE>
__global__ void __launch_bounds__(kWarpThreads, MOG_BWD_TMA_MINB)
stn_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmG, const BwdArgs a) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int NJC = 2, SW = kTmaSW, TR = kTmaTR, RB = kTmaRB;
    static_assert(SW == kBwdSW && TR % RB == 0, "tile shape");
    const Geo& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* s_base = s_raw + (size_t)warp * bwd_tma_warp_smem_bytes(g);
    float* s_U = reinterpret_cast<float*>(s_base);
    float* s_ring = reinterpret_cast<float*>(s_base + align128(g.S * 4));
    int4* s_rt = reinterpret_cast<int4*>(reinterpret_cast<unsigned char*>(s_ring) + kTmaStages * kTmaTileBytes);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_rt) + 128);
    uint64_t* bar_U = s_bar + kTmaStages;
    const EmitSmem em = emit_smem(reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s_bar) + 128), g);

    for (int k = lane; k < kBwdSlots * (g.Ws + 1); k += 32) em.x[k] = 0.f;
    if (lane == 0) {
        for (int s = 0; s <= kTmaStages; ++s) mbar_init(s_bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;   // bit s: parity the next wait on stage s expects; bit kTmaStages: the source barrier

    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const int SC = g.S;
    const int ws4 = g.Ws * 4;
    const float half_wsc = g.wsc * 0.5f, half_hsc = g.hsc * 0.5f;
    const bool need_taps = a.dtheta != nullptr || (COMPOSITE && a.dz != nullptr);

    for (long long b = (long long)blockIdx.x * kWarpsPerCta + warp; b < a.Bsrc; b += nwarps) {
        const float* __restrict__ Ub = a.U + b * (long long)SC;
        float* __restrict__ dUb = a.dU ? a.dU + b * (long long)SC : nullptr;
        if (dUb) {
            fill_zero(dUb, 0, SC, lane);
            __syncwarp();
        }
        Theta th;
        th.load(a.theta + 6 * b);
        float z = 1.0f;
        bool active = true;
        if (COMPOSITE) {
            z = __ldg(a.z_pres + b);
            active = a.stop_sum ? (__ldg(a.stop_sum + b) < a.threshold) : true;
        }
        const float* __restrict__ gb = a.gout + b * (long long)g.N;
        float p[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

        if (active && !th.separable()) {   // general affine theta: cold path on global memory
            bwd_general_image<COMPOSITE>(Ub, dUb, gb, a.dtheta ? a.dtheta + 6 * b : nullptr, (COMPOSITE && a.dz) ? a.dz + b : nullptr,
                                         th.t[0], th.t[1], th.t[2], th.t[3], th.t[4], th.t[5], z, false, lane, g.Hs, g.Ws, 1, g.Ho,
                                         g.Wo, g.step_w, g.step_h, g.wsc, g.hsc);
            continue;
        }
        if (active) {
            const bool need_dU = dUb != nullptr;
            int ilo = g.Ho, ihi = -1, jlo = g.Wo, jhi = -1;
            for (int i = lane; i < g.Ho; i += 32)
                if (row_params(th, g, i, ws4).in) { ilo = min(ilo, i); ihi = max(ihi, i); }
            for (int j = lane; j < g.Wo; j += 32) {
                const Axis X = col_axis(th, g, j);
                if (X.c0 != X.c1) { jlo = min(jlo, j); jhi = max(jhi, j); }
            }
            ilo = __reduce_min_sync(0xffffffffu, ilo); ihi = __reduce_max_sync(0xffffffffu, ihi);
            jlo = __reduce_min_sync(0xffffffffu, jlo); jhi = __reduce_max_sync(0xffffffffu, jhi);
            if (ihi >= ilo && jhi >= jlo) {
                const bool ascending = !(th.t[4] < 0.0f);
                const int nrows = ihi - ilo + 1;
                const int ntile = (nrows + TR - 1) / TR;
                // the source image, whole (every lane is past the previous image's reads of s_U)
                if (lane == 0) {
                    mbar_expect_tx(bar_U, (unsigned)(g.S * 4));
                    tma_load_3d(s_U, &tmU, bar_U, 0, 0, (int)b);
                }
                bool u_ready = false;
                const char* Usc = reinterpret_cast<const char*>(s_U);
                char* dUbc = reinterpret_cast<char*>(dUb);

                const int jbase = jlo & ~3;   // tile columns start on a 16-byte boundary
                int strip_index = 0;
                for (int js = jbase; js <= jhi; js += SW, ++strip_index) {
                    const int jfirst = max(js, jlo), je = min(js + SW, jhi + 1);   // valid columns of the strip: [jfirst, je)
                    // tile t of this strip: stream rows [t TR, t TR + TR) = box rows of the tensor at (js, ybox(t), b)
                    auto issue_tile = [&](int t) {
                        if (lane == 0) {
                            const int st = t % kTmaStages;
                            const int ybox = ascending ? ilo + t * TR : ihi - t * TR - (TR - 1);
                            mbar_expect_tx(s_bar + st, (unsigned)kTmaTileBytes);
                            tma_load_3d(reinterpret_cast<unsigned char*>(s_ring) + st * kTmaTileBytes, &tmG, s_bar + st, js, ybox, (int)b);
                        }
                    };
                    for (int t = 0; t < min(kTmaStages, ntile); ++t) issue_tile(t);

                    Strip<NJC> sp;
                    strip_setup<NJC, COMPOSITE>(sp, th, g, js, jfirst, je, jlo, jhi, z, lane, need_dU, strip_index);
                    float A[NJC], Bv[NJC], A2[NJC], B2[NJC], car[NJC], SX[NJC], SXY[NJC], SY[NJC], SYY[NJC], SZ[NJC];
#pragma unroll
                    for (int c = 0; c < NJC; ++c)
                        A[c] = Bv[c] = A2[c] = B2[c] = car[c] = SX[c] = SXY[c] = SY[c] = SYY[c] = SZ[c] = 0.f;
                    int ycar = -1;

                    for (int t = 0; t < ntile; ++t) {
                        const int st = t % kTmaStages;
                        // the tile's row parameters (lane r: stream row t TR + r; lane TR supplies the row after the tile)
                        {
                            const int ii = t * TR + lane;
                            const bool have = lane <= TR && ii < nrows;
                            RowP r;
                            r.y = -2; r.yoff = 0; r.ay = r.by = r.yt = 0.f; r.in = false;
                            if (have) r = row_params(th, g, ascending ? ilo + ii : ihi - ii, ws4);
                            const int ynext = __shfl_down_sync(0xffffffffu, r.y, 1);
                            const bool last = ii + 1 >= nrows || ynext != r.y;
                            if (have && lane < TR)
                                s_rt[lane] = make_int4((r.y << 1) | (last ? 1 : 0), __float_as_int(r.ay), __float_as_int(r.by), __float_as_int(r.yt));
                        }
                        mbar_wait(s_bar + st, (phase >> st) & 1u);
                        phase ^= 1u << st;
                        if (!u_ready) {
                            mbar_wait(bar_U, (phase >> kTmaStages) & 1u);
                            phase ^= 1u << kTmaStages;
                            u_ready = true;
                        }
                        __syncwarp();
                        const float* tile = s_ring + st * (kTmaTileBytes / 4);
                        const int nr = min(TR, nrows - t * TR);
#pragma unroll 1
                        for (int r0 = 0; r0 < nr; r0 += RB) {
                            const int nb = min(RB, nr - r0);
                            int ey[RB];
                            float gq[NJC][RB], I[NJC][RB][4];
#pragma unroll
                            for (int r = 0; r < RB; ++r) {
                                const int rr = r0 + min(r, nb - 1);
                                ey[r] = s_rt[rr].x;
                                if (r >= nb) ey[r] &= ~1;
                                const float* trow = tile + (ascending ? rr : TR - 1 - rr) * SW + lane;
#pragma unroll
                                for (int c = 0; c < NJC; ++c) {
                                    gq[c][r] = trow[32 * c];
                                    if ((ey[r] & 1) && need_taps) {
                                        const char* pa = Usc + ((ey[r] >> 1) * ws4 + sp.xo[c]);
                                        I[c][r][0] = *reinterpret_cast<const float*>(pa);
                                        I[c][r][2] = *reinterpret_cast<const float*>(pa + 4);
                                        I[c][r][1] = *reinterpret_cast<const float*>(pa + ws4);
                                        I[c][r][3] = *reinterpret_cast<const float*>(pa + ws4 + 4);
                                    }
                                }
                            }
                            const int nslots = batch_arith<NJC, RB, COMPOSITE>(sp, em, s_rt + r0, nb, ey, gq, I, A, Bv, A2, B2, car, SX, SXY, SY, SYY, SZ,
                                                                               ycar, need_taps, need_dU, lane);
                            if (nslots > 0) emit_slots<NJC>(sp, em, g, dUbc, nslots, lane, true);
                        }
                        __syncwarp();   // every lane is done with stage st and with the row table
                        if (t + kTmaStages < ntile) issue_tile(t + kTmaStages);
                    }
                    if (need_dU && ycar >= 0) {   // lower row of the last group
                        if (lane == 0) em.sloty[0] = ycar;
#pragma unroll
                        for (int c = 0; c < NJC; ++c) em.v[32 * c + lane] = car[c];
                        emit_slots<NJC>(sp, em, g, dUbc, 1, lane, true);
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
                __syncwarp();   // reads of s_U are over before the next image's copy may land
            }
            const float sw_ = COMPOSITE ? half_wsc * z : half_wsc, sh_ = COMPOSITE ? half_hsc * z : half_hsc;
            p[0] *= sw_; p[1] *= sw_; p[2] *= sw_;
            p[3] *= sh_; p[4] *= sh_; p[5] *= sh_;
        }
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

}  // namespace mog
