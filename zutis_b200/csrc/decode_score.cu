// Fused bilinear-upsample -> argmax -> int16 labels -> confusion histogram (sm_100a).
//
// Replaces, without ever writing full-resolution logits to HBM:
//   F.interpolate(lo, size, "bilinear") ; torch.argmax(dim=1)     networks/zutis.py:366-372
//   RunningScore._fast_hist                                        utils/running_score.py:10-16
//
// Three kernels (the third, decode_cells_kernel with exact per-cell candidate pruning, lives in decode_cells.cu and is
// what ZUTIS_DECODE_AUTO picks for pixel-major logits and >= 4x up-sampling; the dispatcher is at the end of this file):
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
#include "decode.cuh"

namespace zutis {


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
//
// Work unit = (image, row group, 32-column block, category chunk).  A row group is <= 8 consecutive
// output rows that share their two low-res source rows (a "cell row", split if longer than 8); the
// table of groups is built once per CTA in shared memory.  Per unit a warp
//   1. prefetches the NEXT unit's taps with cp.async into the other half of its private double buffer,
//   2. runs the chunked argmax over the current tile:  for every 4 categories and every row
//         v0..v3 = lerp (FMUL+FFMA each, FMA pipe);  m = max(v0..v3);  if (m > best) {best = m; group = g}
//      i.e. the half-rate ALU pipe sees ~1.25 instructions per (category,pixel) instead of 3,
//   3. after the last chunk re-evaluates the 4 categories of each row's winning group to recover the
//      exact first-max index (bit-identical values, so `==` finds it), writes int16 labels and
//      counts (gt,pred) pairs.
// A tile that contains a non-finite tap takes the plain per-category loop with torch's NaN ordering.


__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int NR, bool NAN_AWARE>
__device__ __forceinline__ void tile_rows_argmax(const float* __restrict__ ra, const float* __restrict__ rb,
                                                 int row_stride, int qc, int q0, float lx0, float lx1,
                                                 const float (&ly0)[8], const float (&ly1)[8],
                                                 float (&best)[8], int (&idx)[8]) {
    const float* rc = ra + row_stride;
    const float* rd = rb + row_stride;
#pragma unroll 2
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


// chunked max tracking: `grp[r]` = index (in units of 4 categories, counted from category 0) of the
// first group whose maximum is the running maximum of row r
template <int NR>
__device__ __forceinline__ void tile_rows_groupmax(const float* __restrict__ ra, const float* __restrict__ rb,
                                                   int row_stride, int ngroups, int g0, float lx0, float lx1,
                                                   const float (&ly0)[8], const float (&ly1)[8],
                                                   float (&best)[8], int (&grp)[8]) {
    const float4* pa = reinterpret_cast<const float4*>(ra);
    const float4* pb = reinterpret_cast<const float4*>(rb);
    const float4* pc = reinterpret_cast<const float4*>(ra + row_stride);
    const float4* pd = reinterpret_cast<const float4*>(rb + row_stride);
    const unsigned long long LX0 = pack2(lx0, lx0), LX1 = pack2(lx1, lx1);
#pragma unroll 2
    for (int g = 0; g < ngroups; ++g) {
        const float4 a = pa[g], b = pb[g], c = pc[g], d = pd[g];
        // t = fma(lx0, a, lx1*b), u = fma(lx0, c, lx1*d) for the 4 categories of the group, two per instruction
        const unsigned long long t01 = fma2(LX0, pack2(a.x, a.y), mul2(LX1, pack2(b.x, b.y)));
        const unsigned long long t23 = fma2(LX0, pack2(a.z, a.w), mul2(LX1, pack2(b.z, b.w)));
        const unsigned long long u01 = fma2(LX0, pack2(c.x, c.y), mul2(LX1, pack2(d.x, d.y)));
        const unsigned long long u23 = fma2(LX0, pack2(c.z, c.w), mul2(LX1, pack2(d.z, d.w)));
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const unsigned long long LY0 = pack2(ly0[r], ly0[r]), LY1 = pack2(ly1[r], ly1[r]);
            float v0, v1, v2, v3;
            unpack2(fma2(LY0, t01, mul2(LY1, u01)), v0, v1);      // v = fma(ly0, t, ly1*u)
            unpack2(fma2(LY0, t23, mul2(LY1, u23)), v2, v3);
            // running maximum folded into two 3-input max instructions; the group index moves only when the
            // maximum moved (strictly greater, so the first group holding the maximum wins)
            const float nb = max3(v2, v3, max3(v0, v1, best[r]));
            if (nb != best[r]) grp[r] = g0 + g;
            best[r] = nb;
        }
    }
}

constexpr int kTiledWarps = 8;
constexpr int kMaxGroups = 1024;      // row groups kept in shared memory (host checks)
constexpr int kMaxTableRows = 4096;   // output rows whose vertical weights are tabulated in shared memory

struct Unit {                         // one (image, row group, column block, category chunk)
    int b, cy, Y0, nr, xb, rx_lo, ch;
};


// Labels + histogram for the rows of one finished item.  GT loads are issued first and consumed last so
// that their latency hides behind the index refinement.
template <typename GT>
__device__ __forceinline__ void finish_rows(const DecodeParams& p, const Unit& un, int X, bool xvalid, int* hist,
                                            const int (&label)[8]) {
    const size_t pix0 = ((size_t)un.b * p.H + un.Y0) * p.W + X;
    if (hist) {
        const GT* gt = reinterpret_cast<const GT*>(p.gt) + (size_t)un.b * p.gt_sb + (size_t)un.Y0 * p.W + X;
        GT g[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) g[r] = (xvalid && r < un.nr) ? gt[(size_t)r * p.W] : (GT)0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r < un.nr) {
                if (p.labels && xvalid) p.labels[pix0 + (size_t)r * p.W] = (int16_t)label[r];
                const int cls = xvalid ? class_of<GT>(g[r], p.n) : -1;
                warp_hist_add(hist, cls >= 0 ? cls * p.n + label[r] : -1);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < un.nr && xvalid) p.labels[pix0 + (size_t)r * p.W] = (int16_t)label[r];
    }
}


__global__ void __launch_bounds__(kTiledWarps * 32, 2) decode_tiled_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nn = p.n * p.n;
    int* s_hist = reinterpret_cast<int*>(smem);
    const int hist_words = p.hist_in_smem ? ((nn + 3) & ~3) : 0;
    int* s_ystart = reinterpret_cast<int*>(smem) + hist_words;                   // [h+1]
    int4* s_groups = reinterpret_cast<int4*>(s_ystart + ((p.h + 1 + 3) & ~3));   // [n_groups]: cy, Y0, nrows
    float2* s_ly = reinterpret_cast<float2*>(s_groups + p.n_groups);             // [H]: (ly0, ly1)
    const int row_stride = p.XR * p.QS;                                          // floats between the two staged rows
    const int tile_floats = 2 * row_stride;
    float* tiles = reinterpret_cast<float*>(s_ly + ((p.H + 1) & ~1)) + warp * (2 * tile_floats);

    // ---- per-CTA tables
    for (int i = threadIdx.x; i < nn && p.hist_in_smem; i += blockDim.x) s_hist[i] = 0;
    for (int cy = threadIdx.x; cy <= p.h; cy += blockDim.x) s_ystart[cy] = first_dst_with_tap_ge(cy, p.h, p.H, p.scale_y);
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) {
        const AxisTap ty = axis_tap(Y, p.h, p.H, p.scale_y);
        s_ly[Y] = make_float2(ty.l0, ty.l1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0;
        for (int cy = 0; cy < p.h; ++cy)
            for (int y = s_ystart[cy]; y < s_ystart[cy + 1] && g < p.n_groups; y += 8)
                s_groups[g++] = make_int4(cy, y, min(8, s_ystart[cy + 1] - y), 0);
        for (; g < p.n_groups; ++g) s_groups[g] = make_int4(0, 0, 0, 0);     // cannot happen (host counts the same way)
    }
    __syncthreads();
    int* hist = p.hist ? (p.hist_in_smem ? s_hist : p.hist) : nullptr;

    const int nchunk = (p.Q + p.QC - 1) / p.QC;
    // A warp owns whole items (its registers carry the running maxima across the chunks of an item):
    // its s-th unit is chunk s % nchunk of item first_item + (s / nchunk) * item_stride.
    // (32-bit index math: the host guarantees n_items * nchunk < 2^31 and per-image offsets < 2^31)
    const unsigned first_item = blockIdx.x * kTiledWarps + warp;
    const unsigned item_stride = gridDim.x * kTiledWarps;
    const unsigned total_items = (unsigned)p.n_items;
    const unsigned my_items = first_item < total_items ? (total_items - first_item + item_stride - 1) / item_stride : 0;
    const unsigned n_units = my_items * nchunk;
    const int sx = (int)p.sx, sy = (int)p.sy, sq = (int)p.sq;

    auto decode_unit = [&](unsigned s_) {
        Unit un;
        unsigned item;
        if (nchunk == 1) { un.ch = 0; item = first_item + s_ * item_stride; }
        else { un.ch = (int)(s_ % (unsigned)nchunk); item = first_item + (s_ / (unsigned)nchunk) * item_stride; }
        un.xb = (int)(item % (unsigned)p.XB);
        const unsigned t = item / (unsigned)p.XB;
        const int4 grp = s_groups[t % (unsigned)p.n_groups];
        un.b = (int)(t / (unsigned)p.n_groups);
        un.cy = grp.x; un.Y0 = grp.y; un.nr = grp.z;
        un.rx_lo = axis_tap(un.xb * 32, p.w, p.W, p.scale_x).i0;
        return un;
    };
    auto stage = [&](const Unit& un, float* tile) {
        const int cy1 = un.cy + (un.cy < p.h - 1 ? 1 : 0);
        const int q0 = un.ch * p.QC;
        const int qc = min(p.QC, p.Q - q0);
        const float* img = p.logits + (long)un.b * p.sb + (long)q0 * p.sq;
        const int off0 = un.cy * sy, off1 = cy1 * sy;
        if (p.vec_stage) {
            const int cpp = (qc + 3) >> 2;                               // 16-byte chunks per pixel (<= 32)
            const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tile) + (uint32_t)lane * 16u;
            if (lane < cpp) {
                for (int rx = 0; rx < p.XR; ++rx) {
                    const int gx = min(un.rx_lo + rx, p.w - 1) * sx + lane * 4;
                    cp_async16(tbase + (uint32_t)(rx * p.QS) * 4u, img + off0 + gx);
                    cp_async16(tbase + (uint32_t)((p.XR + rx) * p.QS) * 4u, img + off1 + gx);
                }
            }
        } else {
            for (int rx = 0; rx < p.XR; ++rx) {
                const int gx = min(un.rx_lo + rx, p.w - 1) * sx;
                for (int j = lane; j < qc; j += 32) {
                    tile[rx * p.QS + j] = __ldg(img + off0 + gx + j * sq);
                    tile[(p.XR + rx) * p.QS + j] = __ldg(img + off1 + gx + j * sq);
                }
            }
        }
        cp_async_commit();
        if (p.hist && un.ch == nchunk - 1) {
            // pull this item's ground-truth rows towards L2 so the loads in finish_rows are short
            const int X = min(un.xb * 32 + lane, p.W - 1);
            const char* g = reinterpret_cast<const char*>(p.gt) + ((size_t)un.b * p.gt_sb + (size_t)un.Y0 * p.W + X) * p.gt_bytes;
            if ((lane & 3) == 0)
                for (int r = 0; r < un.nr; ++r)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g + (size_t)r * p.W * p.gt_bytes));
        }
    };

    unsigned u = 0;
    int buf = 0;
    Unit cur;
    if (n_units > 0) { cur = decode_unit(0); stage(cur, tiles); }

    float ly0[8], ly1[8], best[8];
    int sel[8];                        // winning group (fast path) or category (NaN path) per row
    bool item_finite = true;

    for (; u < n_units; ++u, buf ^= 1) {
        float* tile = tiles + buf * tile_floats;
        const Unit un = cur;
        const int q0 = un.ch * p.QC;
        const int qc = min(p.QC, p.Q - q0);
        const int ngroups = (qc + 3) >> 2;
        const int X = un.xb * 32 + lane;
        const bool xvalid = X < p.W;
        const AxisTap tx = axis_tap(xvalid ? X : p.W - 1, p.w, p.W, p.scale_x);

        cp_async_wait_all();
        __syncwarp();
        if (u + 1 < n_units) { cur = decode_unit(u + 1); stage(cur, tiles + (buf ^ 1) * tile_floats); }

        // Pad the last group with -inf and look for non-finite taps: x * 0 + acc stays (+-)0 for finite x and turns
        // into NaN for NaN / +-inf, which costs one packed FMA-pipe instruction per two taps.
        unsigned long long poison = 0ull;                                 // packed (0.f, 0.f)
        {
            const int step4 = p.QS >> 2, npix = 2 * p.XR;
            if (lane < ngroups) {
                float4* q4 = reinterpret_cast<float4*>(tile) + lane;
                if (!((lane == ngroups - 1) && (qc & 3))) {
#pragma unroll 4
                    for (int pix = 0; pix < npix; ++pix) {
                        const float4 v = q4[pix * step4];
                        poison = fma2(pack2(v.x, v.y), 0ull, poison);
                        poison = fma2(pack2(v.z, v.w), 0ull, poison);
                    }
                } else {
                    const int keep = qc & 3;
                    for (int pix = 0; pix < npix; ++pix) {
                        float4 v = q4[pix * step4];
                        poison = fma2(pack2(v.x, keep > 1 ? v.y : 0.f), 0ull, poison);
                        poison = fma2(pack2(keep > 2 ? v.z : 0.f, 0.f), 0ull, poison);
                        if (keep < 2) v.y = -INFINITY;
                        if (keep < 3) v.z = -INFINITY;
                        v.w = -INFINITY;
                        q4[pix * step4] = v;
                    }
                }
            }
            for (int g = lane + 32; g < ngroups; g += 32) {               // QC > 128 never happens, kept for safety
                for (int pix = 0; pix < npix; ++pix) {
                    const float4 v = reinterpret_cast<const float4*>(tile)[pix * step4 + g];
                    poison = fma2(pack2(v.x, v.y), 0ull, poison);
                    poison = fma2(pack2(v.z, v.w), 0ull, poison);
                }
            }
        }
        float poison_x, poison_y;
        unpack2(poison, poison_x, poison_y);
        const bool chunk_finite = __all_sync(0xffffffffu, poison_x == 0.f && poison_y == 0.f);   // NaN compares false
        __syncwarp();

        if (un.ch == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 l = s_ly[min(un.Y0 + r, p.H - 1)];
                ly0[r] = l.x; ly1[r] = l.y; best[r] = -INFINITY; sel[r] = 0;
            }
            item_finite = true;
            if (nchunk > 1) {
                // several chunks: every chunk must use the same ordering rule, so look at all taps of the item now
                bool all_ok = true;
                const int cy1 = un.cy + (un.cy < p.h - 1 ? 1 : 0);
                const float* img = p.logits + (long)un.b * p.sb;
                for (int pix = 0; pix < 2 * p.XR; ++pix) {
                    const int ry = pix >= p.XR;
                    const int gx = min(un.rx_lo + (pix - ry * p.XR), p.w - 1);
                    const float* src = img + (ry ? cy1 : un.cy) * sy + gx * sx;
                    for (int j = lane; j < p.Q; j += 32) all_ok = all_ok && (fabsf(__ldg(src + j * sq)) <= 3.402823466e38f);
                }
                item_finite = __all_sync(0xffffffffu, all_ok);
            }
        }
        if (nchunk == 1) item_finite = chunk_finite;

        const float* ra = tile + (tx.i0 - un.rx_lo) * p.QS;
        const float* rb = tile + (tx.i1 - un.rx_lo) * p.QS;
        const int nr = un.nr;
        if (item_finite) {
            if (nr > 4) tile_rows_groupmax<8>(ra, rb, row_stride, ngroups, q0 >> 2, tx.l0, tx.l1, ly0, ly1, best, sel);
            else if (nr > 2) tile_rows_groupmax<4>(ra, rb, row_stride, ngroups, q0 >> 2, tx.l0, tx.l1, ly0, ly1, best, sel);
            else if (nr > 1) tile_rows_groupmax<2>(ra, rb, row_stride, ngroups, q0 >> 2, tx.l0, tx.l1, ly0, ly1, best, sel);
            else tile_rows_groupmax<1>(ra, rb, row_stride, ngroups, q0 >> 2, tx.l0, tx.l1, ly0, ly1, best, sel);
        } else {
            if (nr > 4) tile_rows_argmax<8, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, sel);
            else if (nr > 2) tile_rows_argmax<4, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, sel);
            else if (nr > 1) tile_rows_argmax<2, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, sel);
            else tile_rows_argmax<1, true>(ra, rb, row_stride, qc, q0, tx.l0, tx.l1, ly0, ly1, best, sel);
        }
        if (un.ch != nchunk - 1) continue;

        // ---- last chunk of the item: exact index inside the winning group of 4, labels, histogram
        int label[8];
        if (item_finite) {
            const int cy1 = un.cy + (un.cy < p.h - 1 ? 1 : 0);
            const float* img = p.logits + (long)un.b * p.sb;
            const float* g_a = img + un.cy * sy + tx.i0 * sx;
            const float* g_b = img + un.cy * sy + tx.i1 * sx;
            const float* g_c = img + cy1 * sy + tx.i0 * sx;
            const float* g_d = img + cy1 * sy + tx.i1 * sx;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int qb = sel[r] * 4;
                float4 a, bb, c, d;
                if (nchunk == 1) {
                    a = *reinterpret_cast<const float4*>(ra + qb); bb = *reinterpret_cast<const float4*>(rb + qb);
                    c = *reinterpret_cast<const float4*>(ra + row_stride + qb); d = *reinterpret_cast<const float4*>(rb + row_stride + qb);
                } else {
                    const float ninf = -INFINITY;
                    a = make_float4(__ldg(g_a + qb * sq), qb + 1 < p.Q ? __ldg(g_a + (qb + 1) * sq) : ninf, qb + 2 < p.Q ? __ldg(g_a + (qb + 2) * sq) : ninf, qb + 3 < p.Q ? __ldg(g_a + (qb + 3) * sq) : ninf);
                    bb = make_float4(__ldg(g_b + qb * sq), qb + 1 < p.Q ? __ldg(g_b + (qb + 1) * sq) : ninf, qb + 2 < p.Q ? __ldg(g_b + (qb + 2) * sq) : ninf, qb + 3 < p.Q ? __ldg(g_b + (qb + 3) * sq) : ninf);
                    c = make_float4(__ldg(g_c + qb * sq), qb + 1 < p.Q ? __ldg(g_c + (qb + 1) * sq) : ninf, qb + 2 < p.Q ? __ldg(g_c + (qb + 2) * sq) : ninf, qb + 3 < p.Q ? __ldg(g_c + (qb + 3) * sq) : ninf);
                    d = make_float4(__ldg(g_d + qb * sq), qb + 1 < p.Q ? __ldg(g_d + (qb + 1) * sq) : ninf, qb + 2 < p.Q ? __ldg(g_d + (qb + 2) * sq) : ninf, qb + 3 < p.Q ? __ldg(g_d + (qb + 3) * sq) : ninf);
                }
                // bit-identical re-evaluation of the four candidates: `==` recovers the first maximum
                const float v0 = __fmaf_rn(ly0[r], lerp_w(tx.l0, a.x, tx.l1, bb.x), __fmul_rn(ly1[r], lerp_w(tx.l0, c.x, tx.l1, d.x)));
                const float v1 = __fmaf_rn(ly0[r], lerp_w(tx.l0, a.y, tx.l1, bb.y), __fmul_rn(ly1[r], lerp_w(tx.l0, c.y, tx.l1, d.y)));
                const float v2 = __fmaf_rn(ly0[r], lerp_w(tx.l0, a.z, tx.l1, bb.z), __fmul_rn(ly1[r], lerp_w(tx.l0, c.z, tx.l1, d.z)));
                int sub = 3;
                if (v2 == best[r]) sub = 2;
                if (v1 == best[r]) sub = 1;
                if (v0 == best[r]) sub = 0;
                label[r] = qb + sub;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) label[r] = sel[r];
        }
        switch (p.gt_dtype) {
            case ZUTIS_GT_U8: finish_rows<uint8_t>(p, un, X, xvalid, hist, label); break;
            case ZUTIS_GT_I16: finish_rows<int16_t>(p, un, X, xvalid, hist, label); break;
            case ZUTIS_GT_I32: finish_rows<int32_t>(p, un, X, xvalid, hist, label); break;
            default: finish_rows<long long>(p, un, X, xvalid, hist, label); break;
        }
    }
    cp_async_wait_all();
    if (p.hist && p.hist_in_smem) {
        __syncthreads();
        flush_shared_hist(s_hist, p.hist, nn);
    }
}


// ------------------------------------------------------------------------- tiled threshold kernel
// interp(probabilities) > threshold -> bit-packed masks (networks/zutis.py:422-425).  Same work decomposition and
// tap staging as decode_tiled_kernel (warp = 32 output columns x <= 8 rows of one cell row, double-buffered
// cp.async tiles); per 4 queries and row the four comparisons are turned into four 32-pixel words by warp ballots,
// lane (4*row + j) keeps word (row, query 4g+j) and the warp issues ONE 32-address store per group of 4 queries.
struct ThresholdParams {
    const float* probs;
    long sb, sq, sy, sx;
    int B, Q, h, w, H, W;
    float scale_y, scale_x, threshold;
    uint32_t* bits;      // [B,Q,H,words]
    int* areas;          // [B,Q] or null
    int words, XB, XR, QC, QS, n_groups, vec_stage;
    long n_items;
};

__global__ void __launch_bounds__(kTiledWarps * 32, 2) threshold_tiled_kernel(const ThresholdParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int* s_ystart = reinterpret_cast<int*>(smem);                                // [h+1]
    int4* s_groups = reinterpret_cast<int4*>(s_ystart + ((p.h + 1 + 3) & ~3));   // [n_groups]: cy, Y0, nrows
    float2* s_ly = reinterpret_cast<float2*>(s_groups + p.n_groups);             // [H]: (ly0, ly1)
    const int row_stride = p.XR * p.QS;
    const int tile_floats = 2 * row_stride;
    float* tiles = reinterpret_cast<float*>(s_ly + ((p.H + 1) & ~1)) + warp * (2 * tile_floats);

    for (int cy = threadIdx.x; cy <= p.h; cy += blockDim.x) s_ystart[cy] = first_dst_with_tap_ge(cy, p.h, p.H, p.scale_y);
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) {
        const AxisTap ty = axis_tap(Y, p.h, p.H, p.scale_y);
        s_ly[Y] = make_float2(ty.l0, ty.l1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0;
        for (int cy = 0; cy < p.h; ++cy)
            for (int y = s_ystart[cy]; y < s_ystart[cy + 1] && g < p.n_groups; y += 8)
                s_groups[g++] = make_int4(cy, y, min(8, s_ystart[cy + 1] - y), 0);
        for (; g < p.n_groups; ++g) s_groups[g] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();

    const int nchunk = (p.Q + p.QC - 1) / p.QC;
    const unsigned n_units_total = (unsigned)p.n_items * nchunk;      // chunks are independent here: any warp may take any unit
    const unsigned first = blockIdx.x * kTiledWarps + warp;
    const unsigned stride = gridDim.x * kTiledWarps;
    const int sx = (int)p.sx, sy = (int)p.sy, sq = (int)p.sq;

    auto decode_unit = [&](unsigned u) {
        Unit un;
        un.ch = (int)(u % (unsigned)nchunk);
        const unsigned item = u / (unsigned)nchunk;
        un.xb = (int)(item % (unsigned)p.XB);
        const unsigned t = item / (unsigned)p.XB;
        const int4 grp = s_groups[t % (unsigned)p.n_groups];
        un.b = (int)(t / (unsigned)p.n_groups);
        un.cy = grp.x; un.Y0 = grp.y; un.nr = grp.z;
        un.rx_lo = axis_tap(un.xb * 32, p.w, p.W, p.scale_x).i0;
        return un;
    };
    auto stage = [&](const Unit& un, float* tile) {
        const int cy1 = un.cy + (un.cy < p.h - 1 ? 1 : 0);
        const int q0 = un.ch * p.QC;
        const int qc = min(p.QC, p.Q - q0);
        const float* img = p.probs + (long)un.b * p.sb + (long)q0 * p.sq;
        const int off0 = un.cy * sy, off1 = cy1 * sy;
        if (p.vec_stage) {
            const int cpp = (qc + 3) >> 2;
            const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tile) + (uint32_t)lane * 16u;
            if (lane < cpp) {
                for (int rx = 0; rx < p.XR; ++rx) {
                    const int gx = min(un.rx_lo + rx, p.w - 1) * sx + lane * 4;
                    cp_async16(tbase + (uint32_t)(rx * p.QS) * 4u, img + off0 + gx);
                    cp_async16(tbase + (uint32_t)((p.XR + rx) * p.QS) * 4u, img + off1 + gx);
                }
            }
        } else {
            for (int rx = 0; rx < p.XR; ++rx) {
                const int gx = min(un.rx_lo + rx, p.w - 1) * sx;
                for (int j = lane; j < ((qc + 3) & ~3); j += 32) {
                    tile[rx * p.QS + j] = j < qc ? __ldg(img + off0 + gx + j * sq) : 0.0f;
                    tile[(p.XR + rx) * p.QS + j] = j < qc ? __ldg(img + off1 + gx + j * sq) : 0.0f;
                }
            }
        }
        cp_async_commit();
    };

    int buf = 0;
    Unit cur;
    if (first < n_units_total) { cur = decode_unit(first); stage(cur, tiles); }
    const int my_row = lane >> 2, my_j = lane & 3;            // the (row, query-in-group) whose word this lane keeps

    for (unsigned u = first; u < n_units_total; u += stride, buf ^= 1) {
        float* tile = tiles + buf * tile_floats;
        const Unit un = cur;
        const int q0 = un.ch * p.QC;
        const int qc = min(p.QC, p.Q - q0);
        const int ngroups = (qc + 3) >> 2;
        const int X = un.xb * 32 + lane;
        const bool xvalid = X < p.W;
        const AxisTap tx = axis_tap(xvalid ? X : p.W - 1, p.w, p.W, p.scale_x);
        cp_async_wait_all();
        __syncwarp();
        if (u + stride < n_units_total) { cur = decode_unit(u + stride); stage(cur, tiles + (buf ^ 1) * tile_floats); }

        unsigned long long LY0[8], LY1[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 l = s_ly[min(un.Y0 + r, p.H - 1)];
            LY0[r] = pack2(l.x, l.x); LY1[r] = pack2(l.y, l.y);
        }
        const unsigned long long LX0 = pack2(tx.l0, tx.l0), LX1 = pack2(tx.l1, tx.l1);
        const float4* pa = reinterpret_cast<const float4*>(tile + (tx.i0 - un.rx_lo) * p.QS);
        const float4* pb = reinterpret_cast<const float4*>(tile + (tx.i1 - un.rx_lo) * p.QS);
        const float4* pc = reinterpret_cast<const float4*>(tile + (tx.i0 - un.rx_lo) * p.QS + row_stride);
        const float4* pd = reinterpret_cast<const float4*>(tile + (tx.i1 - un.rx_lo) * p.QS + row_stride);
        const float thr = p.threshold;
        // word (row my_row, query q0 + 4g + my_j) goes to bits[((b*Q + q)*H + Y)*words + xb]
        const bool row_ok = my_row < un.nr;
        uint32_t* out = p.bits + (((size_t)un.b * p.Q + q0 + my_j) * p.H + un.Y0 + my_row) * p.words + un.xb;
        const size_t q_step = (size_t)4 * p.H * p.words;
#pragma unroll 1
        for (int g = 0; g < ngroups; ++g) {
            const float4 a = pa[g], b = pb[g], c = pc[g], d = pd[g];
            const unsigned long long t01 = fma2(LX0, pack2(a.x, a.y), mul2(LX1, pack2(b.x, b.y)));
            const unsigned long long t23 = fma2(LX0, pack2(a.z, a.w), mul2(LX1, pack2(b.z, b.w)));
            const unsigned long long u01 = fma2(LX0, pack2(c.x, c.y), mul2(LX1, pack2(d.x, d.y)));
            const unsigned long long u23 = fma2(LX0, pack2(c.z, c.w), mul2(LX1, pack2(d.z, d.w)));
            unsigned mine = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float v0, v1, v2, v3;
                unpack2(fma2(LY0[r], t01, mul2(LY1[r], u01)), v0, v1);
                unpack2(fma2(LY0[r], t23, mul2(LY1[r], u23)), v2, v3);
                const unsigned w0 = __ballot_sync(0xffffffffu, xvalid && v0 > thr);
                const unsigned w1 = __ballot_sync(0xffffffffu, xvalid && v1 > thr);
                const unsigned w2 = __ballot_sync(0xffffffffu, xvalid && v2 > thr);
                const unsigned w3 = __ballot_sync(0xffffffffu, xvalid && v3 > thr);
                if (my_row == r) mine = my_j == 0 ? w0 : (my_j == 1 ? w1 : (my_j == 2 ? w2 : w3));
            }
            const bool q_ok = 4 * g + my_j < qc;
            if (row_ok && q_ok) out[(size_t)g * q_step] = mine;
            if (p.areas) {
                int cnt = (row_ok && q_ok) ? __popc(mine) : 0;
                cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
                if (lane < 4 && q_ok && cnt) atomicAdd(p.areas + (size_t)un.b * p.Q + q0 + 4 * g + lane, cnt);
            }
        }
    }
    cp_async_wait_all();
}

// Host side of the tiled threshold kernel; returns ZUTIS_ERR_UNSUPPORTED when the shape needs the generic kernel.
int launch_threshold_tiled(const float* probs, long sb, long sq, long sy, long sx, int B, int Q, int h, int w, int H, int W,
                           float threshold, uint32_t* bits, int* areas, cudaStream_t stream) {
    if (H < h || W < w || (H == h && W == w)) return ZUTIS_ERR_UNSUPPORTED;
    ThresholdParams p;
    p.probs = probs; p.sb = sb; p.sq = sq; p.sy = sy; p.sx = sx;
    p.B = B; p.Q = Q; p.h = h; p.w = w; p.H = H; p.W = W;
    p.scale_y = axis_scale(h, H); p.scale_x = axis_scale(w, W); p.threshold = threshold;
    p.bits = bits; p.areas = areas; p.words = (W + 31) / 32; p.XB = p.words;
    int XR = 0;
    for (int xb = 0; xb < p.XB; ++xb) {
        const int lo = axis_tap(xb * 32, w, W, p.scale_x).i0;
        const int last = (xb * 32 + 31 < W) ? xb * 32 + 31 : W - 1;
        const int hi = axis_tap(last, w, W, p.scale_x).i1;
        if (hi - lo + 1 > XR) XR = hi - lo + 1;
    }
    if (XR > 8) return ZUTIS_ERR_UNSUPPORTED;
    p.XR = XR;
    p.QC = Q <= 128 ? ((Q + 3) & ~3) : 128;
    p.QS = p.QC;
    while ((p.QS & 7) != 4) p.QS += 4;
    int groups = 0, prev = 0;
    for (int cy = 0; cy < h; ++cy) {
        int lo = prev, hi = H;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (axis_tap(mid, h, H, p.scale_y).i0 >= cy + 1) hi = mid; else lo = mid + 1; }
        groups += (lo - prev + 7) / 8;
        prev = lo;
    }
    p.n_groups = groups;
    p.n_items = (long)B * groups * p.XB;
    p.vec_stage = (sq == 1) && ((sx & 3) == 0) && ((sy & 3) == 0) && ((sb & 3) == 0) && (sx >= ((Q + 3) & ~3)) &&
                  ((reinterpret_cast<uintptr_t>(probs) & 15) == 0);
    const int nchunk = (Q + p.QC - 1) / p.QC;
    const size_t smem = (size_t)((h + 1 + 3) & ~3) * 4 + (size_t)groups * 16 + (size_t)((H + 1) & ~1) * 8 +
                        (size_t)kTiledWarps * 2 * 2 * XR * p.QS * 4;
    const long extent = (long)(h - 1) * sy + (long)(w - 1) * sx + (long)(Q - 1) * sq;
    if (groups > kMaxGroups || H > kMaxTableRows || smem > 200 * 1024 || extent >= 2147483647L || sx < 0 || sy < 0 || sq < 0 ||
        p.n_items * nchunk >= 2147483647L)
        return ZUTIS_ERR_UNSUPPORTED;
    ZUTIS_CUDA(cudaFuncSetAttribute(threshold_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    ZUTIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, threshold_tiled_kernel, kTiledWarps * 32, smem));
    if (per_sm < 1) return ZUTIS_ERR_UNSUPPORTED;
    long blocks = (p.n_items * nchunk + kTiledWarps - 1) / kTiledWarps;
    const long cap = (long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    threshold_tiled_kernel<<<(unsigned)blocks, kTiledWarps * 32, smem, stream>>>(p);
    return check_launch("threshold_tiled_kernel");
}

}  // namespace zutis

using namespace zutis;

// workspace of the cell kernel: its global run counter and the finished-CTA count (two words, 16 bytes reserved)
extern "C" size_t zutis_decode_workspace_bytes(int B, int Q, int h, int w, int H, int W) {
    (void)B; (void)Q; (void)h; (void)w; (void)H; (void)W;
    return 16;
}

static int decode_score_impl(const float* logits, long sb, long sq, long sy, long sx,
                             int B, int Q, int h, int w, int H, int W,
                             const void* gt, int gt_dtype, long gt_sb,
                             int16_t* labels, int32_t* hist_partial, int n_classes,
                             int mode, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // the caller may promise that the workspace words are zero (fresh cudaMemset, or last used by this entry point)
    const bool workspace_zeroed = (mode & ZUTIS_DECODE_WORKSPACE_ZEROED) != 0;
    mode &= ~ZUTIS_DECODE_WORKSPACE_ZEROED;
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
    ZUTIS_REQUIRE(mode >= ZUTIS_DECODE_AUTO && mode <= ZUTIS_DECODE_CELLS, "zutis_decode_score: bad mode %d", mode);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;

    DecodeParams p;
    p.logits = logits; p.sb = sb; p.sq = sq; p.sy = sy; p.sx = sx;
    p.B = B; p.Q = Q; p.h = h; p.w = w; p.H = H; p.W = W;
    p.scale_y = axis_scale(h, H); p.scale_x = axis_scale(w, W);
    p.gt = gt; p.gt_dtype = gt_dtype; p.gt_sb = gt_sb;
    p.labels = labels; p.hist = hist_partial; p.n = hist_partial ? n_classes : 1;
    p.identity = (H == h && W == w);
    p.XB = (W + 31) / 32; p.XR = 0; p.QC = 0; p.QS = 0; p.hist_in_smem = 0; p.n_items = 0; p.n_groups = 0; p.vec_stage = 0; p.gt_bytes = gt ? gt_dtype_bytes(gt_dtype) : 0;

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
    const long extent_px = (long)(h - 1) * sy + (long)(w - 1) * sx + (long)(Q - 1);
    // ---- cell kernel: exact per-cell pruning; pixel-major logits (category index contiguous, 16-byte aligned pixels) and
    // cells of >= 4x4 output pixels.  No workspace, every image (non-finite taps take its brute-force path).
    const bool cells_ok = !p.identity && sq == 1 && sx > 0 && sy > 0 && H >= 4 * h && W >= 4 * w && Q >= 2 && Q <= 65535 &&
                          extent_px < 2147483647L && (sx & 3) == 0 && (sy & 3) == 0 && (sb & 3) == 0 && sx >= ((Q + 3) & ~3) &&
                          (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && H <= 16384 && W <= 16384;
    if (mode == ZUTIS_DECODE_CELLS && !cells_ok)
        return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: cell kernel needs contiguous categories with 16-byte aligned pixels and >= 4x up-sampling");
    if (cells_ok && (mode == ZUTIS_DECODE_CELLS || mode == ZUTIS_DECODE_AUTO)) {
        // with a workspace the runs are handed out through a global counter (dynamic balance over the SMs)
        unsigned* counter = nullptr;
        if (workspace && workspace_bytes >= 16 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0) {
            counter = reinterpret_cast<unsigned*>(workspace);
            if (!workspace_zeroed) ZUTIS_CUDA(cudaMemsetAsync(counter, 0, 16, stream));
        }
        bool launched = false;
        st = launch_decode_cells(p, mode == ZUTIS_DECODE_CELLS, hist_partial ? gt_dtype : ZUTIS_GT_I64, counter, sms, stream, &launched);
        if (st != ZUTIS_OK || launched) return st;
    }
    const bool use_tiled = (mode == ZUTIS_DECODE_TILED) || (mode == ZUTIS_DECODE_AUTO && tiled_ok);

    if (use_tiled) {
        p.XR = XR;
        p.QC = Q <= 128 ? ((Q + 3) & ~3) : 128;
        p.QS = p.QC;                                        // pitch == 4 (mod 8) words: LDS.128 of <= 8 taps hit disjoint banks
        while ((p.QS & 7) != 4) p.QS += 4;
        const int nn = p.n * p.n;
        p.hist_in_smem = (hist_partial != nullptr) && (nn * 4 <= 64 * 1024);
        // row groups, counted exactly as the kernel builds them
        int groups = 0;
        {
            int prev = 0;
            for (int cy = 0; cy < h; ++cy) {
                int lo = prev, hi = H;                       // first Y whose tap index is >= cy+1
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (axis_tap(mid, h, H, p.scale_y).i0 >= cy + 1) hi = mid; else lo = mid + 1; }
                groups += (lo - prev + 7) / 8;
                prev = lo;
            }
        }
        p.n_groups = groups;
        p.n_items = (long)B * groups * p.XB;
        p.vec_stage = (sq == 1) && ((sx & 3) == 0) && ((sy & 3) == 0) && ((sb & 3) == 0) && (sx >= ((Q + 3) & ~3)) &&
                      ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
        const size_t smem = (size_t)(p.hist_in_smem ? ((nn + 3) & ~3) : 0) * 4 + (size_t)((h + 1 + 3) & ~3) * 4 + (size_t)groups * 16 +
                            (size_t)((H + 1) & ~1) * 8 + (size_t)kTiledWarps * 2 * 2 * XR * p.QS * 4;
        const long extent = (long)(h - 1) * sy + (long)(w - 1) * sx + (long)(Q - 1) * sq;      // per-image offsets stay 32-bit in the kernel
        if (groups > kMaxGroups || H > kMaxTableRows || smem > 200 * 1024 || extent >= 2147483647L || sx < 0 || sy < 0 || sq < 0 ||
            p.n_items * ((Q + p.QC - 1) / p.QC) >= 2147483647L) {
            if (mode == ZUTIS_DECODE_TILED)
                return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: tiled kernel does not fit this shape (groups=%d smem=%zu)", groups, smem);
        } else {
            ZUTIS_CUDA(cudaFuncSetAttribute(decode_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 1;
            ZUTIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_tiled_kernel, kTiledWarps * 32, smem));
            if (per_sm < 1) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: tiled kernel does not fit (smem %zu)", smem);
            const int nchunk = (Q + p.QC - 1) / p.QC;
            long blocks = (p.n_items * nchunk + kTiledWarps - 1) / kTiledWarps;
            const long cap = (long)sms * per_sm;
            if (blocks > cap) blocks = cap;
            decode_tiled_kernel<<<(unsigned)blocks, kTiledWarps * 32, smem, stream>>>(p);
            return check_launch("decode_tiled_kernel");
        }
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

extern "C" int zutis_decode_score(const float* logits, long sb, long sq, long sy, long sx,
                                  int B, int Q, int h, int w, int H, int W,
                                  const void* gt, int gt_dtype, long gt_sb,
                                  int16_t* labels, int32_t* hist_partial, int n_classes,
                                  int mode, void* stream) {
    return decode_score_impl(logits, sb, sq, sy, sx, B, Q, h, w, H, W, gt, gt_dtype, gt_sb, labels, hist_partial, n_classes, mode,
                             nullptr, 0, stream);
}

extern "C" int zutis_decode_score_ws(const float* logits, long sb, long sq, long sy, long sx,
                                     int B, int Q, int h, int w, int H, int W,
                                     const void* gt, int gt_dtype, long gt_sb,
                                     int16_t* labels, int32_t* hist_partial, int n_classes,
                                     int mode, void* workspace, size_t workspace_bytes, void* stream) {
    return decode_score_impl(logits, sb, sq, sy, sx, B, Q, h, w, H, W, gt, gt_dtype, gt_sb, labels, hist_partial, n_classes, mode,
                             workspace, workspace_bytes, stream);
}
