// Fused bilinear-upsample -> argmax -> int16 labels -> confusion histogram (sm_100a).
//
// Replaces, without ever writing full-resolution logits to HBM:
//   F.interpolate(lo, size, "bilinear") ; torch.argmax(dim=1)     networks/zutis.py:366-372
//   RunningScore._fast_hist                                        utils/running_score.py:10-16
//
// Two kernels:
//   decode_generic_kernel : one thread per output pixel, taps read straight from global memory.
//                           Any scale (also down-sampling), any strides, identity (size=None),
//                           NaN-exact.  The always-correct path.
//   decode_tiled_kernel   : one warp per (image, low-res cell row, 32 output columns).  The two
//                           low-res rows x few columns x all categories that the warp needs are
//                           staged in a warp-private shared-memory tile; each lane owns one output
//                           column and up to 8 output rows of the cell row, so the horizontal
//                           interpolation (top/bot) is computed once per category and reused by
//                           all rows.  Labels leave as coalesced int16 rows; (gt,pred) pairs go to
//                           a CTA-shared Q x Q histogram with warp-aggregated atomics.
//
// Arithmetic (both kernels, bit-exact with oracle/zutis_oracle.c and with ATen's CPU kernel):
//   top = fma(lx0, a, lx1*b); bot = fma(lx0, c, lx1*d); v = fma(ly0, top, ly1*bot); first max wins.
// This file is compiled with -fmad=false so only the fmas written here exist.
#include "common.cuh"

namespace zutis {

struct DecodeParams {
    const float* logits;
    long sb, sq, sy, sx;
    int B, Q, h, w, H, W;
    float scale_y, scale_x;
    const void* gt;
    int gt_dtype;
    long gt_sb;
    int16_t* labels;
    int* hist;       // global int32 [n*n] partial
    int n;
    int identity;    // H == h && W == w: plain argmax
    // tiled kernel only
    int XB;          // 32-column blocks per row
    int XR;          // staged low-res columns per block
    int QC;          // categories per staged chunk
    int QS;          // shared-memory stride between staged pixels (floats)
    int hist_in_smem;
    long n_items;
};

// ------------------------------------------------------------------------------ generic kernel
__global__ void __launch_bounds__(256) decode_generic_kernel(const DecodeParams p) {
    const long total = (long)p.B * p.H * p.W;
    const long nchunks = (total + 31) >> 5;
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long chunk = warp0; chunk < nchunks; chunk += nwarps) {
        const long i = chunk * 32 + lane;
        int key = -1;
        if (i < total) {
            const int X = (int)(i % p.W);
            const long t = i / p.W;
            const int Y = (int)(t % p.H);
            const int b = (int)(t / p.H);
            const float* img = p.logits + (long)b * p.sb;
            float best = -INFINITY;
            int idx = 0;
            if (p.identity) {
                const float* px = img + (long)Y * p.sy + (long)X * p.sx;
                for (int q = 0; q < p.Q; ++q) {
                    const float v = __ldg(px + (long)q * p.sq);
                    if (q == 0 || better_nan_aware(v, best)) { best = v; idx = q; }
                }
            } else {
                const AxisTap ty = axis_tap(Y, p.h, p.H, p.scale_y);
                const AxisTap tx = axis_tap(X, p.w, p.W, p.scale_x);
                const float* pa = img + (long)ty.i0 * p.sy + (long)tx.i0 * p.sx;
                const float* pb = img + (long)ty.i0 * p.sy + (long)tx.i1 * p.sx;
                const float* pc = img + (long)ty.i1 * p.sy + (long)tx.i0 * p.sx;
                const float* pd = img + (long)ty.i1 * p.sy + (long)tx.i1 * p.sx;
                for (int q = 0; q < p.Q; ++q) {
                    const long o = (long)q * p.sq;
                    const float top = lerp_w(tx.l0, __ldg(pa + o), tx.l1, __ldg(pb + o));
                    const float bot = lerp_w(tx.l0, __ldg(pc + o), tx.l1, __ldg(pd + o));
                    const float v = __fmaf_rn(ty.l0, top, __fmul_rn(ty.l1, bot));
                    if (q == 0 || better_nan_aware(v, best)) { best = v; idx = q; }
                }
            }
            if (p.labels) p.labels[i] = (int16_t)idx;
            if (p.hist) {
                const long long g = load_label(p.gt, p.gt_dtype, (size_t)b * p.gt_sb + (size_t)Y * p.W + X);
                if (g >= 0 && g < p.n) key = (int)g * p.n + idx;
            }
        }
        if (p.hist) warp_hist_add(p.hist, key);
    }
}

// -------------------------------------------------------------------------------- tiled kernel
// smallest d in [0,out] whose first tap index is >= target
__device__ __forceinline__ int first_dst_with_tap_ge(int target, int in, int out, float scale) {
    int lo = 0, hi = out;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (axis_tap(mid, in, out, scale).i0 >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int NR, bool NAN_AWARE>
__device__ __forceinline__ void tile_rows_argmax(const float* __restrict__ ra, const float* __restrict__ rb,
                                                 int row_stride, int qc, int q0, float lx0, float lx1,
                                                 const float (&ly0)[8], const float (&ly1)[8],
                                                 float (&best)[8], int (&idx)[8]) {
    const float* rc = ra + row_stride;
    const float* rd = rb + row_stride;
#pragma unroll 4
    for (int j = 0; j < qc; ++j) {
        const float top = lerp_w(lx0, ra[j], lx1, rb[j]);
        const float bot = lerp_w(lx0, rc[j], lx1, rd[j]);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const float v = __fmaf_rn(ly0[r], top, __fmul_rn(ly1[r], bot));
            const bool take = NAN_AWARE ? better_nan_aware(v, best[r]) : (v > best[r]);
            if (take) { best[r] = v; idx[r] = q0 + j; }
        }
    }
}

constexpr int kTiledWarps = 8;

__global__ void __launch_bounds__(kTiledWarps * 32) decode_tiled_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nn = p.n * p.n;
    int* s_hist = reinterpret_cast<int*>(smem);
    const int hist_words = p.hist_in_smem ? ((nn + 3) & ~3) : 0;
    const int row_stride = p.XR * p.QS;                       // floats between the two staged rows
    float* tile = smem + hist_words + warp * (2 * row_stride);
    if (p.hist_in_smem) {
        for (int i = threadIdx.x; i < nn; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    int* hist = p.hist ? (p.hist_in_smem ? s_hist : p.hist) : nullptr;

    for (long item = (long)blockIdx.x * kTiledWarps + warp; item < p.n_items; item += (long)gridDim.x * kTiledWarps) {
        const int xb = (int)(item % p.XB);
        const long t = item / p.XB;
        const int cy = (int)(t % p.h);
        const int b = (int)(t / p.h);
        const int Ya = first_dst_with_tap_ge(cy, p.h, p.H, p.scale_y);
        const int Yb = first_dst_with_tap_ge(cy + 1, p.h, p.H, p.scale_y);
        if (Ya >= Yb) continue;
        const int X = xb * 32 + lane;
        const bool xvalid = X < p.W;
        const AxisTap tx = axis_tap(xvalid ? X : p.W - 1, p.w, p.W, p.scale_x);
        const int rx_lo = axis_tap(xb * 32, p.w, p.W, p.scale_x).i0;
        const int cy1 = cy + (cy < p.h - 1 ? 1 : 0);
        const float* img = p.logits + (long)b * p.sb;
        const float* ra = tile + (tx.i0 - rx_lo) * p.QS;
        const float* rb = tile + (tx.i1 - rx_lo) * p.QS;
        const int nchunk = (p.Q + p.QC - 1) / p.QC;
        int staged = -1;
        bool finite = true;
        if (nchunk > 1) {
            // Several chunks: learn up front whether any tap of this item is non-finite, so every chunk
            // uses the same ordering rule (the taps are re-read from L1/L2 when they are staged).
            bool ok = true;
            for (int pix = 0; pix < 2 * p.XR; ++pix) {
                const int ry = pix >= p.XR;
                const int gx = min(rx_lo + (pix - ry * p.XR), p.w - 1);
                const float* src = img + (long)(ry ? cy1 : cy) * p.sy + (long)gx * p.sx;
                for (int j = lane; j < p.Q; j += 32) ok = ok && (fabsf(__ldg(src + (long)j * p.sq)) <= 3.402823466e38f);
            }
            finite = __all_sync(0xffffffffu, ok);
        }

        for (int Ys = Ya; Ys < Yb;) {
            const int rem = Yb - Ys;
            const int nr = rem >= 8 ? 8 : (rem >= 4 ? 4 : (rem >= 2 ? 2 : 1));
            float ly0[8], ly1[8], best[8];
            int idx[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const AxisTap ty = axis_tap(min(Ys + r, p.H - 1), p.h, p.H, p.scale_y);
                ly0[r] = ty.l0; ly1[r] = ty.l1; best[r] = -INFINITY; idx[r] = 0;
            }
            for (int ch = 0; ch < nchunk; ++ch) {
                const int q0 = ch * p.QC;
                const int qc = min(p.QC, p.Q - q0);
                if (staged != ch) {
                    __syncwarp();
                    bool ok = true;
                    for (int pix = 0; pix < 2 * p.XR; ++pix) {
                        const int ry = pix >= p.XR;
                        const int gx = min(rx_lo + (pix - ry * p.XR), p.w - 1);
                        const float* src = img + (long)(ry ? cy1 : cy) * p.sy + (long)gx * p.sx + (long)q0 * p.sq;
                        float* dst = tile + pix * p.QS;
                        for (int j = lane; j < qc; j += 32) {
                            const float v = __ldg(src + (long)j * p.sq);
                            ok = ok && (fabsf(v) <= 3.402823466e38f);
                            dst[j] = v;
                        }
                    }
                    if (nchunk == 1) finite = __all_sync(0xffffffffu, ok);
                    staged = ch;
                    __syncwarp();
                }
                // A non-finite tap anywhere in this item switches the whole item to torch's NaN ordering.
                if (finite) {
                    switch (nr) {
                        case 8: tile_rows_argmax<8, false>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                        case 4: tile_rows_argmax<4, false>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                        case 2: tile_rows_argmax<2, false>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                        default: tile_rows_argmax<1, false>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                    }
                } else {
                    switch (nr) {
                        case 8: tile_rows_argmax<8, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                        case 4: tile_rows_argmax<4, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                        case 2: tile_rows_argmax<2, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                        default: tile_rows_argmax<1, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, idx); break;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r < nr) {
                    const int Y = Ys + r;
                    if (p.labels && xvalid) p.labels[((size_t)b * p.H + Y) * p.W + X] = (int16_t)idx[r];
                    if (hist) {
                        int key = -1;
                        if (xvalid) {
                            const long long g = load_label(p.gt, p.gt_dtype, (size_t)b * p.gt_sb + (size_t)Y * p.W + X);
                            if (g >= 0 && g < p.n) key = (int)g * p.n + idx[r];
                        }
                        warp_hist_add(hist, key);
                    }
                }
            }
            Ys += nr;
        }
    }
    if (p.hist && p.hist_in_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            const int v = s_hist[i];
            if (v) atomicAdd(p.hist + i, v);
        }
    }
}

}  // namespace zutis

using namespace zutis;

extern "C" int zutis_decode_score(const float* logits, long sb, long sq, long sy, long sx,
                                  int B, int Q, int h, int w, int H, int W,
                                  const void* gt, int gt_dtype, long gt_sb,
                                  int16_t* labels, int32_t* hist_partial, int n_classes,
                                  int mode, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ZUTIS_REQUIRE(logits != nullptr, "zutis_decode_score: logits is NULL");
    ZUTIS_REQUIRE(B > 0 && Q > 0 && h > 0 && w > 0 && H > 0 && W > 0,
                  "zutis_decode_score: non-positive shape B=%d Q=%d h=%d w=%d H=%d W=%d", B, Q, h, w, H, W);
    ZUTIS_REQUIRE(Q <= 32767, "zutis_decode_score: Q=%d does not fit int16 labels", Q);
    ZUTIS_REQUIRE((long)B * H * W < 2147483647L, "zutis_decode_score: B*H*W=%ld overflows int32 partial counts", (long)B * H * W);
    if (hist_partial) {
        ZUTIS_REQUIRE(gt != nullptr, "zutis_decode_score: hist_partial given without gt");
        ZUTIS_REQUIRE(gt_dtype_bytes(gt_dtype) > 0, "zutis_decode_score: bad gt_dtype %d", gt_dtype);
        ZUTIS_REQUIRE(n_classes >= Q && n_classes <= 46340, "zutis_decode_score: need Q <= n_classes <= 46340 (Q=%d n_classes=%d)", Q, n_classes);
        ZUTIS_REQUIRE(gt_sb >= (long)H * W || B == 1, "zutis_decode_score: gt_sb=%ld smaller than H*W", gt_sb);
    }
    ZUTIS_REQUIRE(labels != nullptr || hist_partial != nullptr, "zutis_decode_score: nothing to produce (labels and hist_partial both NULL)");
    ZUTIS_REQUIRE(mode >= ZUTIS_DECODE_AUTO && mode <= ZUTIS_DECODE_PRUNED, "zutis_decode_score: bad mode %d", mode);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;

    DecodeParams p;
    p.logits = logits; p.sb = sb; p.sq = sq; p.sy = sy; p.sx = sx;
    p.B = B; p.Q = Q; p.h = h; p.w = w; p.H = H; p.W = W;
    p.scale_y = axis_scale(h, H); p.scale_x = axis_scale(w, W);
    p.gt = gt; p.gt_dtype = gt_dtype; p.gt_sb = gt_sb;
    p.labels = labels; p.hist = hist_partial; p.n = hist_partial ? n_classes : 1;
    p.identity = (H == h && W == w);
    p.XB = (W + 31) / 32; p.XR = 0; p.QC = 0; p.QS = 0; p.hist_in_smem = 0; p.n_items = 0;

    const int sms = sm_count();

    // ---- can the tiled kernel take this shape?  (up-sampling, few low-res columns per 32 outputs)
    bool tiled_ok = !p.identity && H >= h && W >= w;
    int XR = 0;
    if (tiled_ok) {
        for (int xb = 0; xb < p.XB; ++xb) {
            const int lo = axis_tap(xb * 32, w, W, p.scale_x).i0;
            const int last = (xb * 32 + 31 < W) ? xb * 32 + 31 : W - 1;
            const int hi = axis_tap(last, w, W, p.scale_x).i1;
            if (hi - lo + 1 > XR) XR = hi - lo + 1;
        }
        if (XR > 8) tiled_ok = false;
    }
    if (mode == ZUTIS_DECODE_TILED && !tiled_ok)
        return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: tiled kernel needs up-sampling with <= 8 low-res columns per 32 outputs");
    if (mode == ZUTIS_DECODE_PRUNED)
        return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: pruned kernel not built yet");
    const bool use_tiled = (mode == ZUTIS_DECODE_TILED) || (mode == ZUTIS_DECODE_AUTO && tiled_ok);

    if (use_tiled) {
        p.XR = XR;
        p.QC = Q <= 96 ? ((Q + 3) & ~3) : 96;
        p.QS = p.QC + 4;                                   // == 4 (mod 8) words: distinct banks for <= 8 taps
        if ((p.QS & 7) != 4) p.QS += 4;
        const int nn = p.n * p.n;
        p.hist_in_smem = (hist_partial != nullptr) && (nn * 4 <= 64 * 1024);
        const size_t smem = (size_t)(p.hist_in_smem ? ((nn + 3) & ~3) : 0) * 4 + (size_t)kTiledWarps * 2 * XR * p.QS * 4;
        p.n_items = (long)B * h * p.XB;
        ZUTIS_CUDA(cudaFuncSetAttribute(decode_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        ZUTIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_tiled_kernel, kTiledWarps * 32, smem));
        if (per_sm < 1) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: tiled kernel does not fit (smem %zu)", smem);
        long blocks = (p.n_items + kTiledWarps - 1) / kTiledWarps;
        const long cap = (long)sms * per_sm;
        if (blocks > cap) blocks = cap;
        decode_tiled_kernel<<<(unsigned)blocks, kTiledWarps * 32, smem, stream>>>(p);
        return check_launch("decode_tiled_kernel");
    }
    {
        const long total = (long)B * H * W;
        long blocks = (total + 255) / 256;
        const long cap = (long)sms * 8;
        if (blocks > cap) blocks = cap;
        decode_generic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(p);
        return check_launch("decode_generic_kernel");
    }
}
