/* ORACLE (test infrastructure, NOT product code) -- scalar C restatement of the reference sampler.
 *
 * Follows /root/reference/air/transformer.py:48-171 (forward) and the closed-form backward that TF
 * autodiff yields for that graph.  It exists for two reasons: (1) a second,
 * independently written evaluation of the op order defined in oracle/stn_ref_numpy.py -- the tests
 * require the two to agree bit for bit on the forward; (2) a CPU baseline that can use every host
 * core (OpenMP over images), which numpy cannot.
 *
 * PARITY UNPINNED: the reference has no tests/golden vectors and TensorFlow 1.12 is not installable
 * here, so TF's evaluation order is *assumed* (see stn_ref_numpy.py for the list).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * the library built from this file.  Build: see oracle/Makefile (-ffp-contract=off is mandatory: the
 * contract is one IEEE fp32 rounding per operation, no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* tf.linspace(-1, 1, n)[i]  (transformer.py:126-128; TF-1.12 LinSpace: start + step*i in fp32) */
static inline float lin(int i, int n) {
    if (n == 1) return -1.0f;
    const float step = 2.0f / (float)(n - 1);
    const float p = step * (float)i;
    return -1.0f + p;
}

/* int32(floor(v)) made total: clamp floor(v) to [-1, hi] first (NaN -> -1); identical clipped
 * corners for every v the reference defines (transformer.py:79-87). */
static inline int sat_floor(float v, int hi) {
    float f = floorf(v);
    f = fmaxf(f, -1.0f);
    f = fminf(f, (float)hi);
    return (int)f;
}

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct {
    float x, y;
    int x0, x1, y0, y1;
} tap_t;

static inline tap_t taps(const float* t, int i, int j, int Hs, int Ws, int Ho, int Wo) {
    tap_t r;
    const float xt = lin(j, Wo), yt = lin(i, Ho);
    /* transformer.py:159  (t0*x_t + t1*y_t) + t2*1 */
    const float a0 = t[0] * xt, a1 = t[1] * yt, a2 = t[2] * 1.0f;
    const float b0 = t[3] * xt, b1 = t[4] * yt, b2 = t[5] * 1.0f;
    const float s0 = a0 + a1, s1 = b0 + b1;
    const float xs = s0 + a2, ys = s1 + b2;
    /* :75-76 */
    const float wsc = (float)Ws - 1.001f, hsc = (float)Hs - 1.001f;
    const float xp = xs + 1.0f, yp = ys + 1.0f;
    const float xm = xp * wsc, ym = yp * hsc;
    r.x = xm / 2.0f;
    r.y = ym / 2.0f;
    /* :79-87 */
    const int fx = sat_floor(r.x, Ws), fy = sat_floor(r.y, Hs);
    r.x0 = clipi(fx, 0, Ws - 1);
    r.x1 = clipi(fx + 1, 0, Ws - 1);
    r.y0 = clipi(fy, 0, Hs - 1);
    r.y1 = clipi(fy + 1, 0, Hs - 1);
    return r;
}

/* corners (nullable) is int32 [4][B*N] in the order x0, x1, y0, y1 */
int stn_ref_forward(const float* U, const float* theta, float* out, int32_t* corners, int64_t B, int Hs,
                    int Ws, int C, int Ho, int Wo, int nthreads) {
    if (!U || !theta || !out || B < 0 || Hs <= 0 || Ws <= 0 || C <= 0 || Ho <= 0 || Wo <= 0) return -1;
    const int64_t N = (int64_t)Ho * Wo, S = (int64_t)Hs * Ws;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const float* t = theta + 6 * b;
        const float* im = U + b * S * C;
        for (int i = 0; i < Ho; ++i)
            for (int j = 0; j < Wo; ++j) {
                const tap_t r = taps(t, i, j, Hs, Ws, Ho, Wo);
                const int64_t n = b * N + (int64_t)i * Wo + j;
                if (corners) {
                    corners[0 * B * N + n] = r.x0;
                    corners[1 * B * N + n] = r.x1;
                    corners[2 * B * N + n] = r.y0;
                    corners[3 * B * N + n] = r.y1;
                }
                /* :108-115 */
                const float x0f = (float)r.x0, x1f = (float)r.x1, y0f = (float)r.y0, y1f = (float)r.y1;
                const float ax = x1f - r.x, bx = r.x - x0f, ay = y1f - r.y, by = r.y - y0f;
                const float wa = ax * ay, wb = ax * by, wc = bx * ay, wd = bx * by;
                const float* pa = im + ((int64_t)r.y0 * Ws + r.x0) * C;
                const float* pb = im + ((int64_t)r.y1 * Ws + r.x0) * C;
                const float* pc = im + ((int64_t)r.y0 * Ws + r.x1) * C;
                const float* pd = im + ((int64_t)r.y1 * Ws + r.x1) * C;
                for (int c = 0; c < C; ++c) {
                    /* :116 add_n in list order */
                    const float ta = wa * pa[c], tb = wb * pb[c], tc = wc * pc[c], td = wd * pd[c];
                    const float s1 = ta + tb;
                    const float s2 = s1 + tc;
                    out[n * C + c] = s2 + td;
                }
            }
    }
    return 0;
}

/* dU nullable (zero-filled here when given); dtheta nullable.  fp32 scatter like the reference's
 * UnsortedSegmentSum, dtheta reduced in double. */
int stn_ref_backward(const float* U, const float* theta, const float* gout, float* dU, float* dtheta,
                     int64_t B, int Hs, int Ws, int C, int Ho, int Wo, int nthreads) {
    if (!U || !theta || !gout || B < 0 || Hs <= 0 || Ws <= 0 || C <= 0 || Ho <= 0 || Wo <= 0) return -1;
    const int64_t N = (int64_t)Ho * Wo, S = (int64_t)Hs * Ws;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const float* t = theta + 6 * b;
        const float* im = U + b * S * C;
        float* dim = dU ? dU + b * S * C : NULL;
        if (dim) memset(dim, 0, sizeof(float) * S * C);
        double acc[6] = {0, 0, 0, 0, 0, 0};
        const float wsc = (float)Ws - 1.001f, hsc = (float)Hs - 1.001f;
        for (int i = 0; i < Ho; ++i)
            for (int j = 0; j < Wo; ++j) {
                const tap_t r = taps(t, i, j, Hs, Ws, Ho, Wo);
                const int64_t n = b * N + (int64_t)i * Wo + j;
                const float x0f = (float)r.x0, x1f = (float)r.x1, y0f = (float)r.y0, y1f = (float)r.y1;
                const float ax = x1f - r.x, bx = r.x - x0f, ay = y1f - r.y, by = r.y - y0f;
                const int64_t ia = ((int64_t)r.y0 * Ws + r.x0) * C, ib = ((int64_t)r.y1 * Ws + r.x0) * C;
                const int64_t ic = ((int64_t)r.y0 * Ws + r.x1) * C, id = ((int64_t)r.y1 * Ws + r.x1) * C;
                double da = 0, db = 0, dc = 0, dd = 0;
                for (int c = 0; c < C; ++c) {
                    const float g = gout[n * C + c];
                    if (dim) {
                        dim[ia + c] += (ax * ay) * g;
                        dim[ib + c] += (ax * by) * g;
                        dim[ic + c] += (bx * ay) * g;
                        dim[id + c] += (bx * by) * g;
                    }
                    da += (double)g * im[ia + c];
                    db += (double)g * im[ib + c];
                    dc += (double)g * im[ic + c];
                    dd += (double)g * im[id + c];
                }
                const double dx = -(double)ay * da - (double)by * db + (double)ay * dc + (double)by * dd;
                const double dy = -(double)ax * da + (double)ax * db - (double)bx * dc + (double)bx * dd;
                const double dxs = dx * (double)wsc / 2.0, dys = dy * (double)hsc / 2.0;
                const double xt = lin(j, Wo), yt = lin(i, Ho);
                acc[0] += dxs * xt; acc[1] += dxs * yt; acc[2] += dxs;
                acc[3] += dys * xt; acc[4] += dys * yt; acc[5] += dys;
            }
        if (dtheta)
            for (int k = 0; k < 6; ++k) dtheta[6 * b + k] = (float)acc[k];
    }
    return 0;
}

int stn_ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
