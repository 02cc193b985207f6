// libmogstn -- fused reconstruction loss of AIR (the step right after the hot path).
//
// Replaces the clip / log / multiply / reduce_sum chain of
// /root/reference/air/air_number_bbox_location.py:945-968 with one forward kernel (per-image cross-entropy and
// squared error, one warp per image, coalesced float4 reads) and one backward kernel.  HBM-bound: 8 B read per
// pixel forward, 8 B read + 4 B written backward.
//   r = max(min(canvas, 1), 0);  loss_b = -sum_p x log(r + 1e-10) + (1 - x) log(1 - r + 1e-10);  mse_b = sum_p (x - r)^2
//   d loss_b / d canvas = [0 <= canvas <= 1] * ( -x / (r + 1e-10) + (1 - x) / (1 - r + 1e-10) )
// (tf.minimum / tf.maximum pass the gradient to their first argument on ties, so both clip bounds are inclusive).
#include "mog_common.cuh"

namespace mog {

constexpr float kBceEps = 1e-10f;

__device__ __forceinline__ void bce_px(float c, float x, float& loss, float& mse) {
    const float r = fmaxf(fminf(c, 1.0f), 0.0f);
    loss -= x * logf(r + kBceEps) + (1.0f - x) * logf(1.0f - r + kBceEps);
    const float d = x - r;
    mse += d * d;
}

__global__ void __launch_bounds__(256) bce_fwd_kernel(const float* __restrict__ canvas, const float* __restrict__ images,
                                                       float* __restrict__ loss, float* __restrict__ mse, long long B, int P) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long b = warp; b < B; b += nwarps) {
        const float* c = canvas + b * P;
        const float* x = images + b * P;
        float l = 0.f, m = 0.f;
        const bool vec = ((reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
        int done = 0;
        if (vec) {
            const int nv = P >> 2;
            const float4* c4 = reinterpret_cast<const float4*>(c);
            const float4* x4 = reinterpret_cast<const float4*>(x);
            for (int k = lane; k < nv; k += 32) {
                const float4 cv = __ldg(c4 + k), xv = __ldg(x4 + k);
                bce_px(cv.x, xv.x, l, m); bce_px(cv.y, xv.y, l, m); bce_px(cv.z, xv.z, l, m); bce_px(cv.w, xv.w, l, m);
            }
            done = nv << 2;
        }
        for (int k = done + lane; k < P; k += 32) bce_px(__ldg(c + k), __ldg(x + k), l, m);
        l = warp_sum(l);
        m = warp_sum(m);
        if (lane == 0) {
            loss[b] = l;
            if (mse) mse[b] = m;
        }
    }
}

__device__ __forceinline__ float bce_dpx(float c, float x, float g) {
    const float r = fmaxf(fminf(c, 1.0f), 0.0f);
    const float d = -x / (r + kBceEps) + (1.0f - x) / (1.0f - r + kBceEps);
    return (c >= 0.0f && c <= 1.0f) ? g * d : 0.0f;
}

// one CTA per image (grid-stride): g_loss[b] is loaded once, pixels stream through as float4
__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ canvas, const float* __restrict__ images,
                                                       const float* __restrict__ gloss, float* __restrict__ dcanvas,
                                                       long long B, int P) {
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const float* c = canvas + b * P;
        const float* x = images + b * P;
        float* d = dcanvas + b * P;
        const float g = __ldg(gloss + b);
        const bool vec = ((reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
        int done = 0;
        if (vec) {
            const int nv = P >> 2;
            const float4* c4 = reinterpret_cast<const float4*>(c);
            const float4* x4 = reinterpret_cast<const float4*>(x);
            float4* d4 = reinterpret_cast<float4*>(d);
            for (int k = threadIdx.x; k < nv; k += blockDim.x) {
                const float4 cv = __ldg(c4 + k), xv = __ldg(x4 + k);
                d4[k] = make_float4(bce_dpx(cv.x, xv.x, g), bce_dpx(cv.y, xv.y, g), bce_dpx(cv.z, xv.z, g), bce_dpx(cv.w, xv.w, g));
            }
            done = nv << 2;
        }
        for (int k = done + threadIdx.x; k < P; k += blockDim.x) d[k] = bce_dpx(__ldg(c + k), __ldg(x + k), g);
    }
}

}  // namespace mog

using namespace mog;

extern "C" int mog_bce_recon_forward(const float* canvas, const float* images, float* loss, float* mse, int64_t B, int P,
                                     void* stream) {
    MOG_REQUIRE(B >= 0 && P > 0, MOG_ERR_DIM, "bce forward: B=%lld P=%d", (long long)B, P);
    if (B == 0) return MOG_OK;
    MOG_REQUIRE(canvas && images && loss, MOG_ERR_NULL, "bce forward: NULL pointer");
    long long blocks = (B + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bce_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(canvas, images, loss, mse, B, P);
    MOG_CUDA_LAUNCH_CHECK("bce_fwd_kernel");
    return MOG_OK;
}

extern "C" int mog_bce_recon_backward(const float* canvas, const float* images, const float* g_loss, float* dcanvas,
                                      int64_t B, int P, void* stream) {
    MOG_REQUIRE(B >= 0 && P > 0, MOG_ERR_DIM, "bce backward: B=%lld P=%d", (long long)B, P);
    if (B == 0) return MOG_OK;
    MOG_REQUIRE(canvas && images && g_loss && dcanvas, MOG_ERR_NULL, "bce backward: NULL pointer");
    long long blocks = B;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bce_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(canvas, images, g_loss, dcanvas, B, P);
    MOG_CUDA_LAUNCH_CHECK("bce_bwd_kernel");
    return MOG_OK;
}
