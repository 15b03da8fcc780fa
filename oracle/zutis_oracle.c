/*
 * oracle/zutis_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the arithmetic on ZUTIS's dense mask-decode + scoring
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product path (zutis_b200/) never does.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout, NoelShin/zutis):
 *   networks/zutis.py:355-372   semantic decode  (einsum -> bilinear -> argmax)
 *   networks/zutis.py:374-427   instance decode  (low-res statistics, bilinear -> threshold)
 *   utils/running_score.py:10-20 confusion-matrix accumulation
 *   utils/iou.py:6-38           binary-mask IoU
 * The bilinear formula is ATen's upsample_bilinear2d (align_corners=False, size= given),
 * the third-party arithmetic the reference calls at zutis.py:367 / :424 (torch is
 * un-pinned by the reference, README.md:78; torch 2.11 is the de-facto oracle version).
 *
 * Parity pinning: the reference ships no tests or golden vectors for this path.  This
 * restatement is pinned instead against outputs of the reference code itself, generated
 * in the build container by tests/golden/make_golden.py and committed under
 * tests/golden/ (see tests/test_oracle_golden.py).
 *
 * Build: see oracle/build_oracle.py  (gcc -O2 -mfma -ffp-contract=off -fopenmp -shared).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fmaf() so the
 * contraction pattern is the one written here and not the compiler's choice.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ZO_EXPORT __attribute__((visibility("default")))

ZO_EXPORT int zo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ZO_EXPORT void zo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------
 * Source-index table of one axis of ATen's bilinear upsample, align_corners=False, no
 * user scale (F.interpolate(size=...), zutis.py:367).  All fp32:
 *     scale = in / out;  src = scale*(dst+0.5) - 0.5, clamped at 0   (one fma)
 *     i0 = min(floor(src), in-1);  i1 = i0 + (i0 < in-1);  l1 = clamp(src-i0,0,1); l0 = 1-l1
 * in == out is ATen's copy branch: i0 = i1 = dst, l0 = 1, l1 = 0.
 * ---------------------------------------------------------------------------------- */
ZO_EXPORT void zo_axis_table(int in, int out, int32_t* i0, int32_t* i1, float* l0, float* l1) {
    if (in == out) {
        for (int d = 0; d < out; ++d) { i0[d] = d; i1[d] = d; l0[d] = 1.0f; l1[d] = 0.0f; }
        return;
    }
    const float scale = (float)in / (float)out;
    for (int d = 0; d < out; ++d) {
        float src = fmaf(scale, (float)d + 0.5f, -0.5f);
        if (src < 0.0f) src = 0.0f;
        int a = (int)floorf(src);
        if (a > in - 1) a = in - 1;
        const int b = a + ((a < in - 1) ? 1 : 0);
        float t = src - (float)a;
        if (t < 0.0f) t = 0.0f;
        if (t > 1.0f) t = 1.0f;
        i0[d] = a; i1[d] = b; l1[d] = t; l0[d] = 1.0f - t;
    }
}

typedef struct {
    int32_t *y0, *y1, *x0, *x1;
    float *ly0, *ly1, *lx0, *lx1;
} zo_tables;

static int zo_tables_make(zo_tables* t, int h, int w, int H, int W) {
    t->y0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)H);
    t->y1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)H);
    t->x0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
    t->x1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
    t->ly0 = (float*)malloc(sizeof(float) * (size_t)H);
    t->ly1 = (float*)malloc(sizeof(float) * (size_t)H);
    t->lx0 = (float*)malloc(sizeof(float) * (size_t)W);
    t->lx1 = (float*)malloc(sizeof(float) * (size_t)W);
    if (!t->y0 || !t->y1 || !t->x0 || !t->x1 || !t->ly0 || !t->ly1 || !t->lx0 || !t->lx1) return -1;
    zo_axis_table(h, H, t->y0, t->y1, t->ly0, t->ly1);
    zo_axis_table(w, W, t->x0, t->x1, t->lx0, t->lx1);
    return 0;
}

static void zo_tables_free(zo_tables* t) {
    free(t->y0); free(t->y1); free(t->x0); free(t->x1);
    free(t->ly0); free(t->ly1); free(t->lx0); free(t->lx1);
}

/* One interpolated value: width first, then height, in the fma pattern that is bit-exact
 * with ATen's CPU kernel:  top = fma(lx0,a, lx1*b);
 * out = fma(ly0, top, ly1*bot). */
static inline float zo_lerp2(const float* p, int w, int y0, int y1, int x0, int x1,
                             float ly0, float ly1, float lx0, float lx1) {
    const float a = p[(size_t)y0 * w + x0], b = p[(size_t)y0 * w + x1];
    const float c = p[(size_t)y1 * w + x0], d = p[(size_t)y1 * w + x1];
    const float top = fmaf(lx0, a, lx1 * b);
    const float bot = fmaf(lx0, c, lx1 * d);
    return fmaf(ly0, top, ly1 * bot);
}

/* F.interpolate(x, size=(H,W), mode="bilinear") on `planes` contiguous [h,w] planes
 * (zutis.py:367, :424).  dst is [planes,H,W]. */
ZO_EXPORT int zo_bilinear_f32(const float* src, long planes, int h, int w, float* dst, int H, int W) {
    zo_tables t;
    if (zo_tables_make(&t, h, w, H, W)) return -1;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < planes; ++p) {
        const float* s = src + (size_t)p * h * w;
        float* o = dst + (size_t)p * H * W;
        for (int Y = 0; Y < H; ++Y)
            for (int X = 0; X < W; ++X)
                o[(size_t)Y * W + X] = zo_lerp2(s, w, t.y0[Y], t.y1[Y], t.x0[X], t.x1[X],
                                                t.ly0[Y], t.ly1[Y], t.lx0[X], t.lx1[X]);
    }
    zo_tables_free(&t);
    return 0;
}

/* torch.argmax rule (zutis.py:372): the first maximal value wins; NaN counts as the
 * maximum, and the first NaN wins. */
static inline int zo_better(float v, float best) {
    return (v > best) || (v != v && best == best);
}

/* torch.einsum("nc,bchw->bnhw") restated with a double accumulator (zutis.py:361-365).
 * text [Q,D], tokens [P,D] (P = B*h*w pixels, channel-last) -> logits [P,Q] (pixel-major)
 * if pixel_major else [Q,P].  Not bit-exact with MKL's summation order by construction;
 * the parity bar for this step is 1e-5 of max|logit|. */
ZO_EXPORT void zo_logits(const float* text, const float* tokens, long P, int Q, int D,
                         float* out, int pixel_major) {
#pragma omp parallel for schedule(static)
    for (long p = 0; p < P; ++p) {
        const float* t = tokens + (size_t)p * D;
        for (int q = 0; q < Q; ++q) {
            const float* e = text + (size_t)q * D;
            double acc = 0.0;
            for (int c = 0; c < D; ++c) acc += (double)e[c] * (double)t[c];
            if (pixel_major) out[(size_t)p * Q + q] = (float)acc;
            else out[(size_t)q * P + p] = (float)acc;
        }
    }
}

/* Semantic decode of zutis.py:366-372 without the [B,Q,H,W] temporary: per output pixel
 * interpolate every category (same arithmetic as zo_bilinear_f32) and keep the argmax.
 * logits: [B,Q,h,w] contiguous.  labels: int64 [B,H,W].  H==0 means size=None: labels at
 * low resolution, plain argmax.  */
ZO_EXPORT int zo_decode_semantic(const float* logits, int B, int Q, int h, int w,
                                 int H, int W, int64_t* labels) {
    if (H == 0 || W == 0) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)B * h * w; ++i) {
            const long b = i / ((long)h * w), r = i % ((long)h * w);
            const float* base = logits + (size_t)b * Q * h * w + r;
            float best = base[0]; int idx = 0;
            for (int q = 1; q < Q; ++q) {
                const float v = base[(size_t)q * h * w];
                if (zo_better(v, best)) { best = v; idx = q; }
            }
            labels[i] = idx;
        }
        return 0;
    }
    zo_tables t;
    if (zo_tables_make(&t, h, w, H, W)) return -1;
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < B; ++b) {
        for (int Y = 0; Y < H; ++Y) {
            const float* img = logits + (size_t)b * Q * h * w;
            int64_t* row = labels + ((size_t)b * H + Y) * W;
            for (int X = 0; X < W; ++X) {
                float best = 0.0f; int idx = 0;
                for (int q = 0; q < Q; ++q) {
                    const float v = zo_lerp2(img + (size_t)q * h * w, w, t.y0[Y], t.y1[Y], t.x0[X], t.x1[X],
                                             t.ly0[Y], t.ly1[Y], t.lx0[X], t.lx1[X]);
                    if (q == 0 || zo_better(v, best)) { best = v; idx = q; }
                }
                row[X] = idx;
            }
        }
    }
    zo_tables_free(&t);
    return 0;
}

/* Instance masks of zutis.py:422-425: interp(probabilities) > threshold (strict), or the
 * low-res comparison of :390 when H==0.  probs [B,Q,h,w] -> masks uint8 [B,Q,H,W]. */
ZO_EXPORT int zo_decode_threshold(const float* probs, int B, int Q, int h, int w,
                                  int H, int W, float threshold, uint8_t* masks) {
    if (H == 0 || W == 0) {
        const long n = (long)B * Q * h * w;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) masks[i] = probs[i] > threshold;
        return 0;
    }
    zo_tables t;
    if (zo_tables_make(&t, h, w, H, W)) return -1;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < (long)B * Q; ++p) {
        const float* s = probs + (size_t)p * h * w;
        uint8_t* o = masks + (size_t)p * H * W;
        for (int Y = 0; Y < H; ++Y)
            for (int X = 0; X < W; ++X)
                o[(size_t)Y * W + X] = zo_lerp2(s, w, t.y0[Y], t.y1[Y], t.x0[X], t.x1[X],
                                                t.ly0[Y], t.ly1[Y], t.lx0[X], t.lx1[X]) > threshold;
    }
    zo_tables_free(&t);
    return 0;
}

/* RunningScore._fast_hist (running_score.py:10-16): rows = ground truth, columns =
 * prediction, only pixels with 0 <= gt < n counted.  The reference does not range-check
 * predictions; a prediction outside [0,n) would alias into another bin (or grow the
 * bincount) -- here it is reported through the return value (count of such pixels) and
 * skipped. hist is int64 [n,n], accumulated into (not cleared). */
ZO_EXPORT long zo_fast_hist(const int64_t* gt, const int64_t* pred, long npix, int n, int64_t* hist) {
    long bad = 0;
    for (long i = 0; i < npix; ++i) {
        const int64_t g = gt[i];
        if (g < 0 || g >= n) continue;
        const int64_t p = pred[i];
        if (p < 0 || p >= n) { ++bad; continue; }
        hist[(size_t)g * n + p] += 1;
    }
    return bad;
}

/* compute_iou (iou.py:23-33), numpy branch: pixels with gt outside [0,1] are dropped,
 * optional strict threshold on the prediction, iou = |p & g| / (|p | g| + eps).
 * Returns intersection/union counts; the division is done by the caller in float64. */
ZO_EXPORT void zo_mask_iou_counts(const float* pred, const float* gt, long npix, int use_threshold,
                                  float threshold, int64_t* inter, int64_t* uni) {
    int64_t a = 0, o = 0;
    for (long i = 0; i < npix; ++i) {
        const float g = gt[i];
        if (!(0.0f <= g && g <= 1.0f)) continue;
        const int pb = use_threshold ? (pred[i] > threshold) : (pred[i] != 0.0f);
        const int gb = (g != 0.0f);
        a += (pb && gb);
        o += (pb || gb);
    }
    *inter = a; *uni = o;
}

/* Low-resolution instance statistics of zutis.py:390-396: per (image, query)
 * mask size = #(p > thr) and sum of in-mask probabilities, accumulated in the order of a
 * plain loop (the reference's torch.sum order is a tree; the bar is fp32 tolerance). */
ZO_EXPORT void zo_instance_lowres_stats(const float* probs, long BQ, long hw, float threshold,
                                        int64_t* sizes, float* psum) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < BQ; ++i) {
        const float* p = probs + (size_t)i * hw;
        int64_t n = 0; double s = 0.0;
        for (long k = 0; k < hw; ++k) if (p[k] > threshold) { ++n; s += (double)p[k]; }
        sizes[i] = n; psum[i] = (float)s;
    }
}
