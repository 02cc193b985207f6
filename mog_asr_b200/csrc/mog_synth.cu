// libmogstn -- object placement of the synthetic Multi-MNIST / Multi-dSprites generator (one thread per canvas).
//
// Restates the placement loop of /root/reference/multi_mnist.py:110-221 (same rules in multi_dsprites.py:92-302):
// per canvas a count is drawn from the configured set, every object gets a sprite and a square size (shared by the
// canvas' objects with share_size, :119,:136-142), and a position is drawn uniformly inside the margins (:171-172) up to
// 100 times (:169) until it passes the overlap rule against the objects already placed; if an object cannot be placed
// the whole canvas starts over (:112-113,:209-210).  Overlap rules:
//   mode 0  the reference's bounding_boxes_overlap (:77-87) LITERALLY: two boxes clash when their x-intervals (the new
//           one widened by `gap`) intersect, whatever their rows; its second test can only fire for degenerate boxes.
//   mode 1  true box intersection (x AND y intervals, widened by `gap`): a conservative stand-in for the pixel-overlap
//           rule (:53-57,:180-182) -- disjoint boxes never share a pixel.
// Randomness is a counter-based hash of (seed, canvas, restart, object, attempt, field), so the oracle
// (oracle/synth_ref.py) reproduces every draw bit for bit and canvases do not depend on batch size or launch shape.
// The pixels are pasted afterwards with the sampler's own forward kernel (mog_asr_b200/dataset.py).
#include "mog_common.cuh"

namespace mog {

__host__ __device__ inline uint32_t mix32(uint32_t x) {  // "lowbias32" finaliser
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__host__ __device__ inline uint32_t draw32(uint32_t seed_lo, uint32_t seed_hi, uint32_t canvas, uint32_t restart, uint32_t object,
                                           uint32_t attempt, uint32_t field) {
    uint32_t h = mix32(seed_lo ^ 0x9e3779b9U);
    h = mix32(h ^ seed_hi);
    h = mix32(h ^ canvas);
    h = mix32(h ^ (restart * 0x85ebca6bU + object));
    h = mix32(h ^ (attempt * 0xc2b2ae35U + field));
    return h;
}
// uniform integer in [lo, hi)  (hi > lo)
__host__ __device__ inline int draw_int(uint32_t h, int lo, int hi) { return lo + (int)(((uint64_t)h * (uint64_t)(hi - lo)) >> 32); }

struct SynthArgs {
    uint32_t seed_lo, seed_hi;
    long long B, first_canvas;
    int canvas, G, ncounts, size_min, size_max, gap, margin, mode, share_size, nsprites, max_restarts;
    int counts[MOG_SYNTH_MAX_COUNTS];
    int* num;
    int* pos;
    int* size;
    int* sprite;
};

__global__ void __launch_bounds__(128) synth_place_kernel(const SynthArgs a) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const uint32_t cid = (uint32_t)(a.first_canvas + b);
    const int n = a.counts[draw_int(draw32(a.seed_lo, a.seed_hi, cid, 0, 0, 0, 7), 0, a.ncounts)];
    int px[MOG_SYNTH_MAX_OBJECTS], py[MOG_SYNTH_MAX_OBJECTS], pw[MOG_SYNTH_MAX_OBJECTS], ps[MOG_SYNTH_MAX_OBJECTS];
    int placed = 0;
    for (int restart = 0; restart < a.max_restarts; ++restart) {
        placed = 0;
        const int shared = draw_int(draw32(a.seed_lo, a.seed_hi, cid, restart, 0, 0, 1), a.size_min, a.size_max + 1);  // :119
        bool ok = true;
        for (int i = 0; i < n && ok; ++i) {
            const int w = a.share_size ? shared : draw_int(draw32(a.seed_lo, a.seed_hi, cid, restart, i, 0, 2), a.size_min, a.size_max + 1);
            const int sp = draw_int(draw32(a.seed_lo, a.seed_hi, cid, restart, i, 0, 3), 0, a.nsprites);
            const int span = a.canvas - w - 2 * a.margin + 1;  // positions margin .. canvas - w - margin (:171-172)
            bool found = false;
            int x = 0, y = 0;
            if (span > 0) {
                for (int att = 0; att < 100 && !found; ++att) {                                                    // :169
                    x = a.margin + draw_int(draw32(a.seed_lo, a.seed_hi, cid, restart, i, att, 4), 0, span);
                    y = a.margin + draw_int(draw32(a.seed_lo, a.seed_hi, cid, restart, i, att, 5), 0, span);
                    found = true;                                                                                     // :177-178
                    for (int k = 0; k < placed && found; ++k) {
                        const int l1x = x - a.gap, l1y = y - a.gap, r1x = x + w + a.gap - 1, r1y = y + w + a.gap - 1;     // :79
                        const int l2x = px[k], l2y = py[k], r2x = px[k] + pw[k] - 1, r2y = py[k] + pw[k] - 1;         // :80
                        const bool xhit = l1x <= r2x && l2x <= r1x;
                        if (a.mode == 0) {
                            if (xhit) found = false;                                                                  // :82-83
                            if (l1y >= r2y && l2y >= r1y) found = false;                                              // :84-85
                        } else {
                            if (xhit && l1y <= r2y && l2y <= r1y) found = false;
                        }
                    }
                }
            }
            if (found) {
                px[placed] = x; py[placed] = y; pw[placed] = w; ps[placed] = sp;                                      // :203-205
                ++placed;
            } else {
                ok = false;                                                                                           // :209-210
            }
        }
        if (ok) break;
    }
    // (when every restart failed, the objects of the last attempt that did fit are kept and num says how many)
    a.num[b] = placed;
    for (int g = 0; g < a.G; ++g) {
        const bool v = g < placed;
        a.pos[(b * a.G + g) * 2] = v ? px[g] : 0;
        a.pos[(b * a.G + g) * 2 + 1] = v ? py[g] : 0;
        a.size[(b * a.G + g) * 2] = v ? pw[g] : 0;
        a.size[(b * a.G + g) * 2 + 1] = v ? pw[g] : 0;
        a.sprite[b * a.G + g] = v ? ps[g] : 0;
    }
}

}  // namespace mog

extern "C" int mog_synth_place(uint64_t seed, int64_t first_canvas, int64_t B, int canvas, int max_objects, const int* counts,
                               int num_counts, int size_min, int size_max, int gap, int margin, int mode, int share_size,
                               int num_sprites, int* num, int* pos, int* size, int* sprite, void* stream) {
    using namespace mog;
    MOG_REQUIRE(B >= 0 && canvas > 0 && max_objects > 0 && max_objects <= MOG_SYNTH_MAX_OBJECTS && num_counts > 0 &&
                    num_counts <= MOG_SYNTH_MAX_COUNTS && size_min > 0 && size_max >= size_min && gap >= 0 && margin >= 0 &&
                    (mode == 0 || mode == 1) && num_sprites > 0 && first_canvas >= 0,
                MOG_ERR_DIM, "synth_place: B=%lld canvas=%d max_objects=%d num_counts=%d size=[%d,%d] gap=%d margin=%d mode=%d sprites=%d",
                (long long)B, canvas, max_objects, num_counts, size_min, size_max, gap, margin, mode, num_sprites);
    if (B == 0) return MOG_OK;
    MOG_REQUIRE(counts && num && pos && size && sprite, MOG_ERR_NULL, "synth_place: NULL pointer");
    SynthArgs a{};
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
    a.B = B; a.first_canvas = first_canvas; a.canvas = canvas; a.G = max_objects; a.ncounts = num_counts;
    a.size_min = size_min; a.size_max = size_max; a.gap = gap; a.margin = margin; a.mode = mode; a.share_size = share_size ? 1 : 0;
    a.nsprites = num_sprites; a.max_restarts = MOG_SYNTH_MAX_RESTARTS;
    for (int k = 0; k < num_counts; ++k) {
        MOG_REQUIRE(counts[k] >= 0 && counts[k] <= max_objects, MOG_ERR_DIM, "synth_place: counts[%d]=%d exceeds max_objects=%d", k,
                    counts[k], max_objects);
        a.counts[k] = counts[k];
    }
    a.num = num; a.pos = pos; a.size = size; a.sprite = sprite;
    synth_place_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    MOG_CUDA_LAUNCH_CHECK("synth_place_kernel");
    return MOG_OK;
}
