// libmogstn -- detection metrics of the evaluation pass (one thread per image).
//
// Replaces the per-image Python loops of /root/reference/air/evaluation_detection.py:29-98 (IoU matrix :44-63,
// precision/recall at 11 thresholds :77-85, max-IoU means :86-87, Hungarian-matched IoU :89-91) with one kernel over
// padded [B,G] ground-truth and [B,T] inferred boxes (G, T <= 8).  float64 with explicit round-to-nearest
// multiplies/adds (no FMA contraction) in the reference's evaluation order, so every output except the matched IoU
// is bit-identical to numpy; the optimal matching is found by a subset DP instead of scipy's shortest-augmenting-path
// solver (same optimum; with ties the pair set may differ, the sum then agrees to an ulp).  Latency-bound, a few
// hundred bytes per image: no roofline claim.
#include "mog_common.cuh"

namespace mog {

constexpr int MAXD = MOG_DET_MAX_BOXES;
constexpr int kDetThreads = 64;

// numpy's add.reduce order for a contiguous run of n <= 8 doubles (pairwise_sum: n < 8 is a plain loop starting from
// the first element, n == 8 uses eight accumulators combined as a tree)
__device__ __forceinline__ double np_sum(const double* v, int n) {
    if (n == 8)
        return __dadd_rn(__dadd_rn(__dadd_rn(v[0], v[1]), __dadd_rn(v[2], v[3])),
                         __dadd_rn(__dadd_rn(v[4], v[5]), __dadd_rn(v[6], v[7])));
    double s = v[0];
    for (int i = 1; i < n; ++i) s = __dadd_rn(s, v[i]);
    return s;
}

struct DetArgs {
    const int* gt_pos;
    const int* gt_size;
    const int* gt_num;
    const double* inf_shifts;
    const double* inf_scales;
    const int* inf_num;
    long long B;
    int G, T;
    double csize_2;
    double* precision;
    double* recall;
    double* gt_max;
    double* det_max;
    double* global_iou;
};

__global__ void __launch_bounds__(kDetThreads) detection_kernel(const DetArgs a) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= a.B) return;
    const int R = max(0, min(a.gt_num[k], a.G)), Cn = max(0, min(a.inf_num[k], a.T));
    double* prec = a.precision + k * 11;
    double* rec = a.recall + k * 11;
    if (R == 0 || Cn == 0) {  // :66-75
        const double both = (R == 0 && Cn == 0) ? 1.0 : 0.0;
        for (int i = 0; i < 11; ++i) {
            prec[i] = both;
            rec[i] = (R == 0) ? 1.0 : 0.0;
        }
        a.gt_max[k] = both;
        a.det_max[k] = both;
        a.global_iou[k] = both;
        return;
    }
    double M[MAXD][MAXD];
    for (int g = 0; g < R; ++g) {
        const int* p = a.gt_pos + (k * a.G + g) * 2;
        const int* s = a.gt_size + (k * a.G + g) * 2;
        const double ax0 = p[0], ay0 = p[1], ax1 = p[0] + s[0], ay1 = p[1] + s[1];   // :45-48 (integer box)
        const double areaA = (double)((long long)(s[0] + 1) * (long long)(s[1] + 1));  // :18 in integers
        for (int t = 0; t < Cn; ++t) {
            const double cx = a.inf_shifts[(k * a.T + t) * 2], cy = a.inf_shifts[(k * a.T + t) * 2 + 1];
            const double half = __dmul_rn(a.inf_scales[k * a.T + t], a.csize_2);
            const double mx = __dmul_rn(__dadd_rn(cx, 1.0), a.csize_2), my = __dmul_rn(__dadd_rn(cy, 1.0), a.csize_2);
            const double bx0 = __dsub_rn(mx, half), by0 = __dsub_rn(my, half);        // :57-61
            const double bx1 = __dadd_rn(mx, half), by1 = __dadd_rn(my, half);
            const double xA = fmax(ax0, bx0), yA = fmax(ay0, by0), xB = fmin(ax1, bx1), yB = fmin(ay1, by1);   // :7-10
            const double iw = fmax(0.0, __dadd_rn(__dsub_rn(xB, xA), 1.0)), ih = fmax(0.0, __dadd_rn(__dsub_rn(yB, yA), 1.0));
            const double inter = __dmul_rn(iw, ih);                                   // :13
            const double areaB = __dmul_rn(__dadd_rn(__dsub_rn(bx1, bx0), 1.0), __dadd_rn(__dsub_rn(by1, by0), 1.0));   // :19
            M[g][t] = __ddiv_rn(inter, __dsub_rn(__dadd_rn(areaA, areaB), inter));    // :24
        }
    }
    // precision / recall (:77-85): a detection is true when ANY ground truth exceeds the threshold
    double colmax[MAXD], rowmax[MAXD];
    for (int t = 0; t < Cn; ++t) {
        double m = M[0][t];
        for (int g = 1; g < R; ++g) m = fmax(m, M[g][t]);
        colmax[t] = m;
    }
    for (int g = 0; g < R; ++g) {
        double m = M[g][0];
        for (int t = 1; t < Cn; ++t) m = fmax(m, M[g][t]);
        rowmax[g] = m;
    }
    for (int i = 0; i < 11; ++i) {
        const double thr = __dadd_rn(__dmul_rn((double)i, 0.05), 0.5);
        int tp = 0;
        for (int t = 0; t < Cn; ++t) tp += (colmax[t] > thr) ? 1 : 0;
        prec[i] = __ddiv_rn((double)tp, (double)Cn);
        rec[i] = __ddiv_rn((double)tp, (double)R);
    }
    a.gt_max[k] = __ddiv_rn(np_sum(rowmax, R), (double)R);      // :86
    a.det_max[k] = __ddiv_rn(np_sum(colmax, Cn), (double)Cn);   // :87

    // maximum-weight matching of min(R,Cn) pairs (:89): items = the smaller side in order, slots = the larger side
    const bool rows_are_items = R <= Cn;
    const int n_items = rows_are_items ? R : Cn, n_slots = rows_are_items ? Cn : R;
    double best[1 << MAXD];
    const int full = 1 << n_slots;
    best[0] = 0.0;
    for (int mask = 1; mask < full; ++mask) {
        const int i = __popc(mask) - 1;  // item placed last
        double b = -1.0;
        if (i < n_items)
            for (int s = 0; s < n_slots; ++s)
                if (mask >> s & 1) {
                    const double prev = best[mask ^ (1 << s)];
                    if (prev >= 0.0) b = fmax(b, prev + (rows_are_items ? M[i][s] : M[s][i]));
                }
        best[mask] = b;
    }
    int arg = 0;
    double top = -1.0;
    for (int mask = 1; mask < full; ++mask)
        if (__popc(mask) == n_items && best[mask] > top) { top = best[mask]; arg = mask; }
    int slot_of[MAXD];
    for (int i = n_items - 1, mask = arg; i >= 0; --i)   // backtrack
        for (int s = 0; s < n_slots; ++s)
            if (mask >> s & 1) {
                const double prev = best[mask ^ (1 << s)];
                if (prev >= 0.0 && prev + (rows_are_items ? M[i][s] : M[s][i]) == best[mask]) {
                    slot_of[i] = s;
                    mask ^= 1 << s;
                    break;
                }
            }
    double pairs[MAXD];  // matched IoUs in ascending ROW order, the order of np.sum(IoU[row_ind, col_ind]) at :90
    if (rows_are_items) {
        for (int i = 0; i < n_items; ++i) pairs[i] = M[i][slot_of[i]];
    } else {
        int n = 0;
        for (int s = 0; s < n_slots; ++s)
            for (int i = 0; i < n_items; ++i)
                if (slot_of[i] == s) pairs[n++] = M[s][i];
    }
    a.global_iou[k] = __ddiv_rn(np_sum(pairs, n_items), (double)(R > Cn ? R : Cn));   // :91
}

}  // namespace mog

extern "C" int mog_detection_eval(const int* gt_pos, const int* gt_size, const int* gt_num, const double* inf_shifts,
                                  const double* inf_scales, const int* inf_num, int64_t B, int G, int T, double csize,
                                  double* precision, double* recall, double* gt_max_iou, double* detected_max_iou,
                                  double* global_iou, void* stream) {
    using namespace mog;
    MOG_REQUIRE(B >= 0 && G >= 0 && T >= 0 && G <= MAXD && T <= MAXD, MOG_ERR_DIM, "detection: B=%lld G=%d T=%d (max %d boxes)",
                (long long)B, G, T, MAXD);
    if (B == 0) return MOG_OK;
    MOG_REQUIRE(gt_num && inf_num && precision && recall && gt_max_iou && detected_max_iou && global_iou &&
                    (G == 0 || (gt_pos && gt_size)) && (T == 0 || (inf_shifts && inf_scales)),
                MOG_ERR_NULL, "detection: NULL pointer");
    DetArgs a{gt_pos, gt_size, gt_num, inf_shifts, inf_scales, inf_num, B, G, T, csize / 2, precision, recall, gt_max_iou,
              detected_max_iou, global_iou};
    detection_kernel<<<(unsigned)((B + kDetThreads - 1) / kDetThreads), kDetThreads, 0, (cudaStream_t)stream>>>(a);
    MOG_CUDA_LAUNCH_CHECK("detection_kernel");
    return MOG_OK;
}
