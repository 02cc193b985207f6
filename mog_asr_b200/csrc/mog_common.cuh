// Shared device helpers for libmogstn (sm_100a).  See include/mogstn.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mogstn.h"

namespace mog {

// ---- error plumbing (host) ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();

#define MOG_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            mog::set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

#define MOG_CUDA_LAUNCH_CHECK(what)                                               \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            mog::set_error("%s: %s", (what), cudaGetErrorString(e__));            \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

// ---- sampler geometry ----------------------------------------------------------------------------------
// Everything that depends only on (Hs, Ws, Ho, Wo); computed once on the host in fp32 so that host
// and device agree to the bit (transformer.py:75-76,126-128).
struct Geo {
    int Hs, Ws, C, Ho, Wo;
    int N;          // Ho*Wo
    int S;          // Hs*Ws
    float step_w;   // fp32(2/(Wo-1)), 0 when Wo == 1   (tf.linspace recurrence)
    float step_h;
    float wsc;      // fp32(Ws) - fp32(1.001)
    float hsc;
    unsigned magic_wo;  // ceil(2^32 / Wo): n / Wo == __umulhi(n, magic_wo) for n*Wo < 2^32 (Wo >= 2)
};

inline Geo make_geo(int Hs, int Ws, int C, int Ho, int Wo) {
    Geo g;
    g.Hs = Hs; g.Ws = Ws; g.C = C; g.Ho = Ho; g.Wo = Wo;
    g.N = Ho * Wo;
    g.S = Hs * Ws;
    g.step_w = Wo > 1 ? 2.0f / (float)(Wo - 1) : 0.0f;
    g.step_h = Ho > 1 ? 2.0f / (float)(Ho - 1) : 0.0f;
    g.wsc = (float)Ws - 1.001f;
    g.hsc = (float)Hs - 1.001f;
    g.magic_wo = Wo > 1 ? (unsigned)((0x100000000ull + (unsigned long long)Wo - 1) / (unsigned long long)Wo) : 0u;
    return g;
}

#ifdef __CUDACC__
// n -> (i, j) with n = i*Wo + j
__device__ __forceinline__ void split_n(const Geo& g, int n, int& i, int& j) {
    if (g.Wo == 1) { i = n; j = 0; return; }
    i = (int)__umulhi((unsigned)n, g.magic_wo);
    j = n - i * g.Wo;
}

// tf.linspace(-1, 1, n)[i] = -1 + step*i, two fp32 roundings (TF-1.12 LinSpace kernel)
__device__ __forceinline__ float lin_at(int i, float step) { return __fadd_rn(-1.0f, __fmul_rn(step, (float)i)); }

// One axis of _interpolate (transformer.py:75-87,108-115): source coordinate -> clipped corners and
// the two 1-D weights taken from the *clipped* corners.  Explicit _rn intrinsics: no FMA contraction,
// one rounding per op, identical to the oracle's stated order.
struct Axis {
    int c0, c1;   // clipped corners
    float a, b;   // a = c1f - p,  b = p - c0f
};

__device__ __forceinline__ Axis axis_tap(float s, float scale, int size) {
    const float p = __fmul_rn(__fmul_rn(__fadd_rn(s, 1.0f), scale), 0.5f);
    // int32(floor(p)) made total: clamp to [-1, size] before the cast (NaN -> -1); same clipped corners
    const float f = fminf(fmaxf(floorf(p), -1.0f), (float)size);
    const int i0 = (int)f;
    Axis r;
    r.c0 = min(max(i0, 0), size - 1);
    r.c1 = min(max(i0 + 1, 0), size - 1);
    r.a = __fsub_rn((float)r.c1, p);
    r.b = __fsub_rn(p, (float)r.c0);
    return r;
}

// transformer.py:159 -- (t0*x_t + t1*y_t) + t2*1, left to right, no FMA
__device__ __forceinline__ float affine_row(float t0, float t1, float t2, float xt, float yt) {
    return __fadd_rn(__fadd_rn(__fmul_rn(t0, xt), __fmul_rn(t1, yt)), t2);
}

struct Theta {
    float t[6];
    __device__ __forceinline__ void load(const float* __restrict__ p) {
#pragma unroll
        for (int k = 0; k < 6; ++k) t[k] = __ldg(p + k);
    }
    // theta built in the kernel from the model's (s, x, y) (air_number_bbox_location.py:511-531, :563-584), same fp32
    // expressions as mog_air_thetas_forward: mode 1 (read) [[s,0,x],[0,s,y]], mode 2 (write) [[1/s,0,-x/s],[0,1/s,-y/s]]
    __device__ __forceinline__ void load_sxy(const float* __restrict__ shift, const float* __restrict__ scale, long long b, int mode) {
        const float s = __ldg(scale + b), x = __ldg(shift + 2 * b), y = __ldg(shift + 2 * b + 1);
        t[1] = 0.0f; t[3] = 0.0f;
        if (mode == 1) {
            t[0] = s; t[2] = x; t[4] = s; t[5] = y;
        } else {
            const float inv = 1.0f / s;
            t[0] = inv; t[2] = -x / s; t[4] = inv; t[5] = -y / s;
        }
    }
    // axis-aligned transform: x_s depends on j only, y_s on i only.  With t01 == 0 the middle product
    // is +-0 and (t00*x_t + +-0) + t02 equals t00*x_t + t02 after the "+1" of transformer.py:75, so the
    // per-column / per-row tables are bit-identical to the per-pixel evaluation.
    __device__ __forceinline__ bool separable() const { return t[1] == 0.0f && t[3] == 0.0f; }
};

// per-pixel evaluation for a general affine theta
__device__ __forceinline__ void taps_general(const Theta& th, const Geo& g, int i, int j, Axis& ax, Axis& ay) {
    const float xt = lin_at(j, g.step_w), yt = lin_at(i, g.step_h);
    ax = axis_tap(affine_row(th.t[0], th.t[1], th.t[2], xt, yt), g.wsc, g.Ws);
    ay = axis_tap(affine_row(th.t[3], th.t[4], th.t[5], xt, yt), g.hsc, g.Hs);
}

// Table entry shared by the separable paths: {c0 (or c0*pitch), c1 (or c1*pitch), a, b} in one 16-byte word.
__device__ __forceinline__ int4 pack_axis(const Axis& a, int pitch) {
    return make_int4(a.c0 * pitch, a.c1 * pitch, __float_as_int(a.a), __float_as_int(a.b));
}

// Build column table [Wo] followed by row table [Ho] for a separable theta.  Column entries hold x0,x1;
// row entries hold y0*Ws, y1*Ws.
__device__ __forceinline__ void build_tables(int4* tab, const Theta& th, const Geo& g) {
    for (int k = threadIdx.x; k < g.Wo + g.Ho; k += blockDim.x) {
        if (k < g.Wo) {
            const float xt = lin_at(k, g.step_w);
            // y_t is irrelevant up to the sign of a zero; use y_t = 0 -> t01*0 = +-0
            tab[k] = pack_axis(axis_tap(affine_row(th.t[0], th.t[1], th.t[2], xt, 0.0f), g.wsc, g.Ws), 1);
        } else {
            const float yt = lin_at(k - g.Wo, g.step_h);
            tab[k] = pack_axis(axis_tap(affine_row(th.t[3], th.t[4], th.t[5], 0.0f, yt), g.hsc, g.Hs), g.Ws);
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__

}  // namespace mog
