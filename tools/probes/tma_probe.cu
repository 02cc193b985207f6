// probe: one cp.async.bulk.tensor.3d box load at unaligned coordinates, checked on the host.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include -I mog_asr_b200/csrc -o build_variants/tma_probe tools/probes/tma_probe.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "mog_stn_bwd_tma.cuh"
namespace mog { void set_error(const char*, ...) {} int sm_count() { return 148; } }
using namespace mog;

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int x, int y, int z, int bytes, int* status) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    float* tile = reinterpret_cast<float*>(s_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + 32768);
    const int lane = threadIdx.x;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, (unsigned)bytes);
        tma_load_3d(tile, &tm, bar, x, y, z);
    }
    bool ok = false;
    for (int it = 0; it < (1 << 20) && !ok; ++it) ok = mbar_try_wait(bar, 0);
    if (lane == 0) *status = ok ? 1 : -1;
    __syncwarp();
    if (ok)
        for (int k = lane; k < bytes / 4; k += 32) out[k] = tile[k];
}

int main(int argc, char** argv) {
    // usage: tma_probe W H bw bh x y
    const int W = atoi(argv[1]), H = atoi(argv[2]), B = 4, bw = atoi(argv[3]), bh = atoi(argv[4]), x = atoi(argv[5]), y = atoi(argv[6]), z = 2;
    std::vector<float> h((size_t)B * H * W);
    for (size_t k = 0; k < h.size(); ++k) h[k] = (float)(k + 1);
    float *d, *out; int* st;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 65536); cudaMalloc(&st, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B}; const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = bw * bh * 4;
    cudaMemset(st, 0, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 33024);
    probe<<<1, 32, 33024>>>(tm, out, x, y, z, bytes, st);
    cudaError_t e = cudaDeviceSynchronize();
    int hs = 0; std::vector<float> ho(bytes / 4);
    cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho.data(), out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < bh; ++r2)
        for (int c = 0; c < bw; ++c) {
            const int yy = y + r2, xx = x + c;
            const float want = (yy >= 0 && xx >= 0 && yy < H && xx < W) ? (float)(((size_t)z * H + yy) * W + xx + 1) : 0.f;
            if (ho[r2 * bw + c] != want) ++bad;
        }
    printf("W=%d H=%d box=%dx%d at (%d,%d): encode=%d sync=%s status=%d mismatches=%d\n", W, H, bw, bh, x, y, (int)r, cudaGetErrorString(e), hs, bad);
    return 0;
}
