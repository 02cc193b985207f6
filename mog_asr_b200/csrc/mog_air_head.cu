// libmogstn -- the two-layer (mean, log-variance) heads of the AIR loop body, everything after the first GEMM fused.
//
// /root/reference/air/air_number_bbox_location.py:424-460 (inference shift / scale heads), :472-481 (prior shift head):
//   h_m = relu(x W1m + b1m), mean = h_m W2m + b2m;  h_v = relu(x W1v + b1v), logvar = h_v W2v + b2v;
//   latent = mean + eps * sqrt(exp(logvar));  squashed = tanh / sigmoid(latent)            (:180-184, :433-436, :456-459)
// and for the scale head the previous shift sample is concatenated to BOTH layers' inputs (:439-455).
// The host computes pre1 = x [W1m | W1v] with one library GEMM (the only part with real arithmetic: K = 256); this kernel
// does the rest per row in one warp: bias, the skip columns of the first layer, relu, the second layer (<= 2 outputs per
// branch), its skip columns and bias, the reparameterised sample and the squashing -- ~22 framework launches per head
// per loop iteration become 2 (GEMM + this), and as many again in the backward pass.
// Weights stay in the nn.Linear layout the parameters have: w1*: [h][K+S] (skip columns at K..K+S), w2*: [O][h+S].
#include "mog_common.cuh"

namespace mog {

constexpr int kHeadWarps = 8;
constexpr int MAXS = MOG_HEAD_MAX_SKIP;
constexpr int MAXO = MOG_HEAD_MAX_OUT;

struct HeadArgs {
    const float* pre1;   // [B][2h]
    const float* skip;   // [B][S] or null
    const float* eps;    // [B][O]
    const float *w1m, *w1v, *b1m, *b1v;  // first layer (only the skip columns and the biases are read here)
    const float *w2m, *w2v, *b2m, *b2v;
    long long B;
    int h, K, S, O, act;
    // forward outputs
    float *mean, *logvar, *latent, *squashed;
    // backward inputs (nullable) / outputs
    const float *g_mean, *g_logvar, *g_latent, *g_squashed;
    const float *logvar_in, *squashed_in;
    float* dpre1;  // [B][2h]
    float* dskip;  // [B][S] or null
    float *gw2m, *gw2v, *gb2m, *gb2v;  // parameter gradients, accumulated with atomics
};

__device__ __forceinline__ float squash(float l, int act) { return act == 1 ? tanhf(l) : (act == 2 ? 1.0f / (1.0f + expf(-l)) : l); }

template <int NQ, bool BACKWARD>
__global__ void __launch_bounds__(kHeadWarps * 32) air_head_kernel(const HeadArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = a.h, S = a.S, O = a.O, H2 = 2 * a.h, ld1 = a.K + a.S, ld2 = a.h + a.S;
    // per-lane constants: hidden units k = lane + 32 q
    float b1[NQ], w1s[NQ][MAXS], w2[NQ][MAXO];
    bool isv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int k = lane + 32 * q;
        isv[q] = k >= h;
        const int kk = isv[q] ? k - h : k;
        b1[q] = isv[q] ? a.b1v[kk] : a.b1m[kk];
#pragma unroll
        for (int s = 0; s < MAXS; ++s) w1s[q][s] = (s < S) ? (isv[q] ? a.w1v : a.w1m)[(long long)kk * ld1 + a.K + s] : 0.0f;
#pragma unroll
        for (int o = 0; o < MAXO; ++o) w2[q][o] = (o < O) ? (isv[q] ? a.w2v : a.w2m)[o * ld2 + kk] : 0.0f;
    }
    float gw2[NQ][MAXO];                         // backward: lane-local weight-gradient partials over this warp's rows
    float gskm[MAXS][MAXO], gskv[MAXS][MAXO], gbm[MAXO], gbv[MAXO];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int o = 0; o < MAXO; ++o) gw2[q][o] = 0.0f;
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
        gbm[o] = gbv[o] = 0.0f;
#pragma unroll
        for (int s = 0; s < MAXS; ++s) gskm[s][o] = gskv[s][o] = 0.0f;
    }

    for (long long b = blockIdx.x * (long long)kHeadWarps + warp; b < a.B; b += (long long)gridDim.x * kHeadWarps) {
        float sk[MAXS];
#pragma unroll
        for (int s = 0; s < MAXS; ++s) sk[s] = (s < S) ? a.skip[b * S + s] : 0.0f;
        float v[NQ], accm[MAXO], accv[MAXO];
#pragma unroll
        for (int o = 0; o < MAXO; ++o) accm[o] = accv[o] = 0.0f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float t = a.pre1[b * H2 + lane + 32 * q] + b1[q];
#pragma unroll
            for (int s = 0; s < MAXS; ++s) t += sk[s] * w1s[q][s];
            v[q] = t;
            const float hv = fmaxf(t, 0.0f);
#pragma unroll
            for (int o = 0; o < MAXO; ++o) {
                const float c = hv * w2[q][o];
                if (isv[q]) accv[o] += c; else accm[o] += c;
            }
        }
        if (!BACKWARD) {
#pragma unroll
            for (int o = 0; o < MAXO; ++o) {
                if (o < O) {
                    float m = warp_sum(accm[o]) + a.b2m[o], lv = warp_sum(accv[o]) + a.b2v[o];
#pragma unroll
                    for (int s = 0; s < MAXS; ++s)
                        if (s < S) { m += sk[s] * a.w2m[o * ld2 + h + s]; lv += sk[s] * a.w2v[o * ld2 + h + s]; }
                    if (lane == o) {
                        const float l = m + a.eps[b * O + o] * sqrtf(expf(lv));
                        a.mean[b * O + o] = m; a.logvar[b * O + o] = lv; a.latent[b * O + o] = l;
                        if (a.act != 0) a.squashed[b * O + o] = squash(l, a.act);
                    }
                }
            }
        } else {
            float dm[MAXO], dv[MAXO];
#pragma unroll
            for (int o = 0; o < MAXO; ++o) {
                dm[o] = dv[o] = 0.0f;
                if (o < O) {
                    const long long i = b * O + o;
                    float dl = a.g_latent ? a.g_latent[i] : 0.0f;
                    if (a.act != 0 && a.g_squashed) {
                        const float sq = a.squashed_in[i];
                        dl += a.g_squashed[i] * (a.act == 1 ? (1.0f - sq * sq) : sq * (1.0f - sq));
                    }
                    dm[o] = (a.g_mean ? a.g_mean[i] : 0.0f) + dl;
                    dv[o] = (a.g_logvar ? a.g_logvar[i] : 0.0f) + dl * a.eps[i] * 0.5f * sqrtf(expf(a.logvar_in[i]));
                }
            }
            float dsk[MAXS];
#pragma unroll
            for (int s = 0; s < MAXS; ++s) dsk[s] = 0.0f;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float dh = 0.0f;
#pragma unroll
                for (int o = 0; o < MAXO; ++o) dh += (isv[q] ? dv[o] : dm[o]) * w2[q][o];
                const float dp = v[q] > 0.0f ? dh : 0.0f;        // relu'(0) = 0 like the framework's threshold_backward
                a.dpre1[b * H2 + lane + 32 * q] = dp;
                const float hv = fmaxf(v[q], 0.0f);
#pragma unroll
                for (int o = 0; o < MAXO; ++o) gw2[q][o] += hv * (isv[q] ? dv[o] : dm[o]);
#pragma unroll
                for (int s = 0; s < MAXS; ++s) dsk[s] += dp * w1s[q][s];
            }
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
                if (s < S) {
                    float d = warp_sum(dsk[s]);
#pragma unroll
                    for (int o = 0; o < MAXO; ++o)
                        if (o < O) d += dm[o] * a.w2m[o * ld2 + h + s] + dv[o] * a.w2v[o * ld2 + h + s];
                    if (lane == s && a.dskip) a.dskip[b * S + s] = d;
#pragma unroll
                    for (int o = 0; o < MAXO; ++o) { gskm[s][o] += sk[s] * dm[o]; gskv[s][o] += sk[s] * dv[o]; }
                }
            }
#pragma unroll
            for (int o = 0; o < MAXO; ++o) { gbm[o] += dm[o]; gbv[o] += dv[o]; }
        }
    }

    if (BACKWARD) {
        // parameter gradients: sum the warps of the CTA in shared memory, then one atomic per value per CTA
        __shared__ float red[kHeadWarps][32 * NQ * MAXO + 2 * MAXS * MAXO + 2 * MAXO];
        float* mine = red[warp];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int o = 0; o < MAXO; ++o) mine[(q * 32 + lane) * MAXO + o] = gw2[q][o];
        if (lane == 0) {
            float* t = mine + 32 * NQ * MAXO;
#pragma unroll
            for (int s = 0; s < MAXS; ++s)
#pragma unroll
                for (int o = 0; o < MAXO; ++o) { t[(s * MAXO + o) * 2] = gskm[s][o]; t[(s * MAXO + o) * 2 + 1] = gskv[s][o]; }
            t += 2 * MAXS * MAXO;
#pragma unroll
            for (int o = 0; o < MAXO; ++o) { t[o * 2] = gbm[o]; t[o * 2 + 1] = gbv[o]; }
        }
        __syncthreads();
        const int total = 32 * NQ * MAXO + 2 * MAXS * MAXO + 2 * MAXO;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            float sum = 0.0f;
#pragma unroll
            for (int w = 0; w < kHeadWarps; ++w) sum += red[w][i];
            if (sum == 0.0f) continue;
            if (i < 32 * NQ * MAXO) {
                const int o = i % MAXO, k = i / MAXO;   // k = q*32 + lane = hidden unit
                if (o < O) atomicAdd((k >= h ? a.gw2v : a.gw2m) + o * ld2 + (k >= h ? k - h : k), sum);
            } else if (i < 32 * NQ * MAXO + 2 * MAXS * MAXO) {
                const int j = i - 32 * NQ * MAXO, isV = j & 1, so = j >> 1, s = so / MAXO, o = so % MAXO;
                if (s < S && o < O) atomicAdd((isV ? a.gw2v : a.gw2m) + o * ld2 + h + s, sum);
            } else {
                const int j = i - 32 * NQ * MAXO - 2 * MAXS * MAXO, isV = j & 1, o = j >> 1;
                if (o < O) atomicAdd((isV ? a.gb2v : a.gb2m) + o, sum);
            }
        }
    }
}

template <bool BACKWARD>
static int launch_head(const HeadArgs& a, cudaStream_t st) {
    long long blocks = (a.B + kHeadWarps - 1) / kHeadWarps;
    const long long cap = (long long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    switch (2 * a.h / 32) {
        case 1: air_head_kernel<1, BACKWARD><<<(int)blocks, kHeadWarps * 32, 0, st>>>(a); break;
        case 2: air_head_kernel<2, BACKWARD><<<(int)blocks, kHeadWarps * 32, 0, st>>>(a); break;
        case 4: air_head_kernel<4, BACKWARD><<<(int)blocks, kHeadWarps * 32, 0, st>>>(a); break;
        case 8: air_head_kernel<8, BACKWARD><<<(int)blocks, kHeadWarps * 32, 0, st>>>(a); break;
        default: set_error("air_head: hidden units %d not in {16, 32, 64, 128}", a.h); return MOG_ERR_UNSUPPORTED;
    }
    MOG_CUDA_LAUNCH_CHECK("air_head_kernel");
    return MOG_OK;
}

static int head_check(const HeadArgs& a, const char* what) {
    MOG_REQUIRE(a.B >= 0 && a.h > 0 && a.K > 0 && a.S >= 0 && a.S <= MAXS && a.O > 0 && a.O <= MAXO && a.act >= 0 && a.act <= 2, MOG_ERR_DIM,
                "%s: B=%lld hidden=%d K=%d skip=%d out=%d act=%d", what, (long long)a.B, a.h, a.K, a.S, a.O, a.act);
    MOG_REQUIRE(a.h == 16 || a.h == 32 || a.h == 64 || a.h == 128, MOG_ERR_UNSUPPORTED, "%s: hidden units %d not in {16, 32, 64, 128}", what, a.h);
    return MOG_OK;
}

}  // namespace mog

using namespace mog;

extern "C" int mog_air_head_forward(const float* pre1, const float* skip, const float* eps, const float* w1m, const float* b1m,
                                    const float* w1v, const float* b1v, const float* w2m, const float* b2m, const float* w2v,
                                    const float* b2v, int64_t B, int hidden, int K, int S, int O, int act, float* mean, float* logvar,
                                    float* latent, float* squashed, void* stream) {
    HeadArgs a{};
    a.pre1 = pre1; a.skip = skip; a.eps = eps; a.w1m = w1m; a.w1v = w1v; a.b1m = b1m; a.b1v = b1v; a.w2m = w2m; a.w2v = w2v;
    a.b2m = b2m; a.b2v = b2v; a.B = B; a.h = hidden; a.K = K; a.S = S; a.O = O; a.act = act;
    a.mean = mean; a.logvar = logvar; a.latent = latent; a.squashed = squashed;
    if (int rc = head_check(a, "air_head forward")) return rc;
    if (B == 0) return MOG_OK;
    MOG_REQUIRE(pre1 && eps && w1m && w1v && b1m && b1v && w2m && w2v && b2m && b2v && mean && logvar && latent && (act == 0 || squashed) &&
                    (S == 0 || skip), MOG_ERR_NULL, "air_head forward: NULL pointer");
    return launch_head<false>(a, (cudaStream_t)stream);
}

extern "C" int mog_air_head_backward(const float* pre1, const float* skip, const float* eps, const float* w1m, const float* b1m,
                                     const float* w1v, const float* b1v, const float* w2m, const float* w2v, const float* logvar,
                                     const float* squashed, const float* g_mean, const float* g_logvar, const float* g_latent,
                                     const float* g_squashed, int64_t B, int hidden, int K, int S, int O, int act, float* dpre1,
                                     float* dskip, float* gw2m, float* gb2m, float* gw2v, float* gb2v, void* stream) {
    HeadArgs a{};
    a.pre1 = pre1; a.skip = skip; a.eps = eps; a.w1m = w1m; a.w1v = w1v; a.b1m = b1m; a.b1v = b1v; a.w2m = w2m; a.w2v = w2v;
    a.B = B; a.h = hidden; a.K = K; a.S = S; a.O = O; a.act = act;
    a.logvar_in = logvar; a.squashed_in = squashed; a.g_mean = g_mean; a.g_logvar = g_logvar; a.g_latent = g_latent; a.g_squashed = g_squashed;
    a.dpre1 = dpre1; a.dskip = dskip; a.gw2m = gw2m; a.gw2v = gw2v; a.gb2m = gb2m; a.gb2v = gb2v;
    if (int rc = head_check(a, "air_head backward")) return rc;
    if (B == 0) return MOG_OK;
    MOG_REQUIRE(pre1 && eps && w1m && w1v && b1m && b1v && w2m && w2v && logvar && dpre1 && gw2m && gw2v && gb2m && gb2v &&
                    (act == 0 || !g_squashed || squashed) && (S == 0 || skip), MOG_ERR_NULL, "air_head backward: NULL pointer");
    return launch_head<true>(a, (cudaStream_t)stream);
}
