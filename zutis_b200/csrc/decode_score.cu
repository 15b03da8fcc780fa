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
    int n_groups;    // row groups (<= 8 rows sharing their source rows) per image
    int vec_stage;   // taps can be staged with 16-byte cp.async (category index contiguous and aligned)
    int gt_bytes;    // sizeof one ground-truth label
    long n_items;    // B * n_groups * XB
    // image selection between the tiled and the pruned kernel (workspace path only)
    int select;              // 0: every image; 1: images the pruned kernel does NOT take; 2: images it takes
    int agree_min;           // an image is pruned when it is finite and >= agree_min neighbouring low-res pixels share their champion
    int* img_stats;          // [B] neighbour agreements | [B] non-finite flags | [B] bits of max |logit| (champion_kernel) | work counter
    const int* champ;        // [B*h*w] first-max category per low-res pixel
    const float* lead;       // [B*h*w] its value minus the largest value of a category with a smaller index
    long lead_delta;         // lead - champ in 4-byte elements
    int cap;                 // candidate slots per warp in the pruned kernel
    // byte offsets of the pruned kernel's tables in dynamic shared memory (computed by the host so that the kernel can
    // re-derive a pointer with one add instead of a chain of size computations when registers run out)
    int off_ystart, off_xstart, off_ly, off_lx, off_img, off_warp;
};

__device__ __forceinline__ bool image_is_pruned(const DecodeParams& p, int b) {
    return p.img_stats[p.B + b] == 0 && p.img_stats[b] >= p.agree_min;
}

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

// smallest d in [0,out] whose first tap index is >= target
__device__ __forceinline__ int first_dst_with_tap_ge(int target, int in, int out, float scale) {
    int lo = 0, hi = out;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (axis_tap(mid, in, out, scale).i0 >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

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

// Packed fp32x2 arithmetic (sm_100 FMUL2 / FFMA2): two IEEE round-to-nearest operations per issued instruction,
// bit-identical to the scalar __fmul_rn / __fmaf_rn they replace.  The decode kernel is issue-bound, so halving the
// FMA-pipe instruction count of the interpolation is worth ~1.5x on the main loop.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
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

template <typename GT>
__device__ __forceinline__ int class_of(GT g, int n) {
    return (g >= (GT)0 && (long long)g < (long long)n) ? (int)g : -1;
}

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

// the images this kernel owns, in ascending order (select: 1 = not pruned, 2 = pruned); one warp, B <= 1024
__device__ __forceinline__ void build_image_list(const DecodeParams& p, int want_pruned, int* s_img, int* s_nimg) {
    if (threadIdx.x < 32) {
        int n = 0;
        for (int b0 = 0; b0 < p.B; b0 += 32) {
            const int b = b0 + (int)threadIdx.x;
            const bool mine = b < p.B && (image_is_pruned(p, b) == (want_pruned != 0));
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (mine) s_img[n + __popc(bal & ((1u << threadIdx.x) - 1u))] = b;
            n += __popc(bal);
        }
        if (threadIdx.x == 0) *s_nimg = n;
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
    int* s_img = reinterpret_cast<int*>(s_ly + ((p.H + 1) & ~1));                // [B] images of this launch (select != 0)
    float* tiles = reinterpret_cast<float*>(s_img + (p.select ? ((p.B + 3) & ~3) : 0)) + warp * (2 * tile_floats);
    __shared__ int s_nimg;

    // ---- which images are ours (the pruned kernel takes the finite, spatially coherent ones)
    if (p.select) {
        // this launch follows the pruned kernel on the stream: re-arm its work counter so that the same workspace
        // (champions included) can serve another decode of the same logits
        if (blockIdx.x == 0 && threadIdx.x == 0) p.img_stats[3 * p.B] = 0;
        build_image_list(p, p.select == 2, s_img, &s_nimg);
        __syncthreads();
        if (s_nimg == 0) return;
    }

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
    const unsigned total_items = p.select ? (unsigned)s_nimg * (unsigned)p.n_groups * (unsigned)p.XB : (unsigned)p.n_items;
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
        if (p.select) un.b = s_img[un.b];
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
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            const int v = s_hist[i];
            if (v) atomicAdd(p.hist + i, v);
        }
    }
}


// ------------------------------------------------------------------------------ pruned decode
// Exact candidate pruning per low-res CELL (the output pixels that share their top-left tap (cy, cx)).
//
// Every interpolant is  v_q = fma(ly0, fma(lx0, A_q, lx1*B_q), ly1 * fma(lx0, C_q, lx1*D_q))  of the cell's four
// corner logits with non-negative weights, and every rounding in that expression is monotone in the taps.  Hence if
// A_k >= A_j, B_k >= B_j, C_k >= C_j and D_k >= D_j then v_k >= v_j at EVERY pixel of the cell, in floating point.
// Category j can therefore never be the first maximum anywhere in the cell when some k dominates it that way and
//   * k < j (a tie still goes to k), or
//   * k > j and the dominance holds with a margin m = 2^-20 * M, M = max|logit of the image|: the exact difference
//     of the two interpolants is then >= m * (1 - 2^-22) (the weights sum to >= 1 - 2^-23, the margin test itself
//     rounds once), while the roundings of one interpolant (two products and three fmas, each relative 2^-24 on terms
//     whose weighted magnitudes sum to <= 4 M) move it by <= 2^-22 M, the pair by <= 2^-21 M < m: v_k > v_j strictly.
// Only the four corner champions (first maxima of the corner pixels, found once per low-res pixel by champion_kernel)
// are tried as dominators.  On model-like logits 6 of 81 categories survive per cell (22 of 920); the survivors'
// corner values are compacted into shared memory and the warp (lane = pixel of an 8x8 tile, 2 pixels per lane) walks
// only those, in ascending category order with a strict compare = torch.argmax's first maximum.
// A warp walks a run of kCellRun horizontally adjacent cells (runs are handed out through an atomic counter): the right corners (B, D) of one cell are the left
// corners (A, C) of the next and stay in registers (NQ = ceil(Q/32) values per lane and corner, NQ = 0: wide Q, taps
// re-read per cell).  The ground truth of a cell's first tile is requested before the pruning work so that its
// latency is hidden.
// Images with a non-finite logit (NaN ordering) or without spatial coherence (pruning would not pay) are left to the
// tiled kernel; both kernels derive the same image split from champion_kernel's per-image counters.
constexpr int kPrunedWarps = 8;
constexpr int kCellRun = 4;

// Per low-res pixel: first-max category and its lead over the categories in front of it; per image: the number of
// horizontally adjacent pixels that share their champion, a non-finite flag and max |logit|.  One block per (image, low-res row); 8 lanes per pixel read the pixel's
// categories as float4 (category index contiguous, 16-byte aligned pixels), 4 pixels per warp at a time.
__global__ void __launch_bounds__(256) champion_kernel(const float* __restrict__ logits, long sb, long sy, long sx, int B, int Q,
                                                       int h, int w, int* __restrict__ champ, float* __restrict__ lead,
                                                       int* __restrict__ stats) {
    extern __shared__ int s_row[];                            // [w]
    __shared__ int s_agree[8];
    const int b = blockIdx.x / h, y = blockIdx.x % h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, grp = lane >> 3;
    const float* row = logits + (long)b * sb + (long)y * sy;
    const int chunks = (Q + 3) >> 2;
    bool bad = false;
    float row_amax = 0.f;
    for (int x0 = warp * 4; x0 < w; x0 += 32) {
        const int x = x0 + grp;
        float best = -INFINITY, amax = 0.f;
        int idx = 0x7fffffff;
        if (x < w) {
            const float4* v = reinterpret_cast<const float4*>(row + (long)x * sx);
            for (int c = sub; c < chunks; c += 8) {
                const float4 f = __ldg(v + c);
                const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = c * 4 + j;
                    if (q < Q) {
                        const float af = fabsf(e[j]);
                        bad = bad || !(af <= 3.402823466e38f);
                        amax = fmaxf(amax, af);
                        if (e[j] > best || idx == 0x7fffffff) { best = e[j]; idx = q; }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
        }
        row_amax = fmaxf(row_amax, amax);
        // second sweep (L1 hits): the best of the categories in front of the champion
        float prev = -INFINITY;
        if (x < w) {
            const float4* v = reinterpret_cast<const float4*>(row + (long)x * sx);
            for (int c = sub; c < chunks; c += 8) {
                const float4 f = __ldg(v + c);
                const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c * 4 + j < idx) prev = fmaxf(prev, e[j]);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) prev = fmaxf(prev, __shfl_xor_sync(0xffffffffu, prev, o));
        if (sub == 0 && x < w) {
            champ[((long)b * h + y) * w + x] = idx;
            lead[((long)b * h + y) * w + x] = __fsub_rn(best, prev);
            s_row[x] = idx;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    int amax_bits = __float_as_int(row_amax);
    for (int o = 16; o > 0; o >>= 1) amax_bits = max(amax_bits, __shfl_xor_sync(0xffffffffu, amax_bits, o));
    if (lane == 0) {
        if (bad) atomicOr(stats + B + b, 1);
        atomicMax(stats + 2 * B + b, amax_bits);
    }
    __syncthreads();
    int agree = 0;
    for (int x = threadIdx.x; x + 1 < w; x += blockDim.x) agree += (s_row[x] == s_row[x + 1]);
    for (int o = 16; o > 0; o >>= 1) agree += __shfl_xor_sync(0xffffffffu, agree, o);
    if (lane == 0) s_agree[warp] = agree;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += s_agree[k];
        if (t) atomicAdd(stats + b, t);
    }
}

template <typename GT, int NQ>
__global__ void __launch_bounds__(kPrunedWarps * 32, 3) decode_pruned_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nn = p.n * p.n;
    char* smem_c = reinterpret_cast<char*>(smem);
    int* s_hist = reinterpret_cast<int*>(smem);                                   // [n*n] when the histogram fits
    int* s_ystart = reinterpret_cast<int*>(smem_c + p.off_ystart);                // [h+1] first output row of each cell row
    int* s_xstart = reinterpret_cast<int*>(smem_c + p.off_xstart);                // [w+1]
    float2* s_ly = reinterpret_cast<float2*>(smem_c + p.off_ly);                  // [H] (ly0, ly1)
    float2* s_lx = reinterpret_cast<float2*>(smem_c + p.off_lx);                  // [W] (lx0, lx1)
    int* s_img = reinterpret_cast<int*>(smem_c + p.off_img);                      // [B] images of this launch
    // per-warp areas behind the tables: survivors' corner values (A, C, B, D), their categories, the champions' corner values
    char* warp_area = smem_c + p.off_warp;
    float4* s_val = reinterpret_cast<float4*>(warp_area) + warp * p.cap;
    int* s_list = reinterpret_cast<int*>(warp_area + (size_t)kPrunedWarps * p.cap * 16) + warp * p.cap;
    float* s_champ = reinterpret_cast<float*>(warp_area + (size_t)kPrunedWarps * p.cap * 20) + warp * 16;
    __shared__ int s_nimg;

    build_image_list(p, 1, s_img, &s_nimg);
    __syncthreads();
    if (s_nimg == 0) return;

    for (int i = threadIdx.x; i < nn && p.hist_in_smem; i += blockDim.x) s_hist[i] = 0;
    for (int c = threadIdx.x; c <= p.h; c += blockDim.x) s_ystart[c] = first_dst_with_tap_ge(c, p.h, p.H, p.scale_y);
    for (int c = threadIdx.x; c <= p.w; c += blockDim.x) s_xstart[c] = first_dst_with_tap_ge(c, p.w, p.W, p.scale_x);
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) { const AxisTap t = axis_tap(Y, p.h, p.H, p.scale_y); s_ly[Y] = make_float2(t.l0, t.l1); }
    for (int X = threadIdx.x; X < p.W; X += blockDim.x) { const AxisTap t = axis_tap(X, p.w, p.W, p.scale_x); s_lx[X] = make_float2(t.l0, t.l1); }
    __syncthreads();
    int* hist = p.hist ? (p.hist_in_smem ? s_hist : p.hist) : nullptr;
    const GT* gt_base = reinterpret_cast<const GT*>(p.gt);

    const int sx = (int)p.sx, sy = (int)p.sy;
    const unsigned runs_per_row = (unsigned)(p.w + kCellRun - 1) / kCellRun;
    const unsigned runs_per_image = runs_per_row * (unsigned)p.h;
    const unsigned total = (unsigned)s_nimg * runs_per_image;
    // runs are handed out dynamically (their cost follows the number of survivors): one atomic per run of kCellRun cells
    unsigned* work_counter = reinterpret_cast<unsigned*>(p.img_stats + 3 * p.B);
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const unsigned slot = item / runs_per_image;
        const unsigned rr = item - slot * runs_per_image;
        const int cy = (int)(rr / runs_per_row);
        const int cx_begin = (int)(rr - (unsigned)cy * runs_per_row) * kCellRun;
        const int cx_end = min(cx_begin + kCellRun, p.w);
        const int b = s_img[slot];
        const int ys = s_ystart[cy], ye = s_ystart[cy + 1];
        if (ys >= ye) continue;
        const int cy1 = min(cy + 1, p.h - 1);
        const float* row0 = p.logits + (long)b * p.sb + cy * sy;      // taps of the cells' upper corners
        const float* row1 = p.logits + (long)b * p.sb + cy1 * sy;     //                    lower corners
        const int* ch0 = p.champ + ((size_t)b * p.h + cy) * p.w;
        const int* ch1 = p.champ + ((size_t)b * p.h + cy1) * p.w;

        // 2^-20 * max|logit of the image|, never 0: a champion must not dominate itself (its differences are exactly 0)
        const float margin = fmaxf(__int_as_float(p.img_stats[2 * p.B + b]) * 9.5367431640625e-07f, 1e-37f);
        // per-lane bases: lane = (row lane/8 [+4], column lane%8) of an 8x8 pixel tile; category lane (+32, +64, ..) of a tap
        const size_t lane_px = (size_t)(ys + (lane >> 3)) * p.W + (lane & 7);
        int16_t* lbl_lane = p.labels ? p.labels + (size_t)b * p.H * p.W + lane_px : nullptr;
        const GT* gt_lane = gt_base + (size_t)b * p.gt_sb + lane_px;
        const float* row0_lane = row0 + lane;
        const float* row1_lane = row1 + lane;
        const int W4 = 4 * p.W;
        if (hist) {
            // pull the run's ground truth towards L2 now (the whole-cell shortcut has nothing to hide its latency behind):
            // lane = (row, 128-byte segment) of the run's pixel rectangle
            const int xs_run = s_xstart[cx_begin], span = (s_xstart[cx_end] - xs_run) * (int)sizeof(GT);
            const int segs = (span + 127) >> 7, r = lane / max(segs, 1), sgm = lane - r * max(segs, 1);
            if (segs > 0 && ys + r < ye && r < 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(gt_base + (size_t)b * p.gt_sb + (size_t)(ys + r) * p.W + xs_run) + sgm * 128));
        }

        // champions of the current cell's four corners; the next cell's right corners are requested one cell ahead
        int hA = __ldg(ch0 + cx_begin), hC = __ldg(ch1 + cx_begin);
        int hB = __ldg(ch0 + min(cx_begin + 1, p.w - 1)), hD = __ldg(ch1 + min(cx_begin + 1, p.w - 1));

        for (int cx = cx_begin; cx < cx_end; ++cx) {
            const int cx1 = min(cx + 1, p.w - 1);
            const int xs = s_xstart[cx], xe = s_xstart[cx + 1];
            const float* pA = row0 + cx * sx;
            const float* pB = row0 + cx1 * sx;
            const float* pC = row1 + cx * sx;
            const float* pD = row1 + cx1 * sx;
            const int cx2 = min(cx + 2, p.w - 1);
            const int hBn = __ldg(ch0 + cx2), hDn = __ldg(ch1 + cx2);
            // One champion at all four corners that leads every category in front of it by the margin at each of them wins
            // the whole cell (the categories behind it can at best tie, and ties go to the smaller index): no taps, no
            // survivor list, no evaluation.  Real segmentation maps are mostly such cells.
            if (hA == hB && hA == hC && hA == hD) {
                // the leads live lead_delta elements behind the champions (one constant instead of two more row pointers)
                const float* ld0 = reinterpret_cast<const float*>(ch0) + p.lead_delta;
                const float* ld1 = reinterpret_cast<const float*>(ch1) + p.lead_delta;
                const float l = fminf(fminf(__ldg(ld0 + cx), __ldg(ld0 + cx1)), fminf(__ldg(ld1 + cx), __ldg(ld1 + cx1)));
                if (l >= margin) {
                    for (int ty = ys; ty < ye; ty += 8) {
                        for (int tx = xs; tx < xe; tx += 8) {
                            const bool okx = tx + (lane & 7) < xe, ok0 = okx && ty + (lane >> 3) < ye, ok1 = okx && ty + (lane >> 3) + 4 < ye;
                            const int tile_off = (ty - ys) * p.W + tx;
                            if (lbl_lane) {
                                if (ok0) lbl_lane[tile_off] = (int16_t)hA;
                                if (ok1) lbl_lane[tile_off + W4] = (int16_t)hA;
                            }
                            if (hist) {
                                const int c0 = ok0 ? class_of<GT>(gt_lane[tile_off], p.n) : -1;
                                const int c1 = ok1 ? class_of<GT>(gt_lane[tile_off + W4], p.n) : -1;
                                warp_hist_add(hist, c0 >= 0 ? c0 * p.n + hA : -1);
                                warp_hist_add(hist, c1 >= 0 ? c1 * p.n + hA : -1);
                            }
                        }
                    }
                    hA = hB; hC = hD; hB = hBn; hD = hDn;
                    continue;
                }
            }
            // the four corner taps of this lane's categories (the left pair was the previous cell's right pair: L1 hits)
            float la_[NQ > 0 ? NQ : 1], lc_[NQ > 0 ? NQ : 1], rb_[NQ > 0 ? NQ : 1], rd_[NQ > 0 ? NQ : 1];
            if (NQ > 0) {
#pragma unroll
                for (int it = 0; it < NQ; ++it) {
                    const bool qv = it * 32 + lane < p.Q;
                    la_[it] = qv ? __ldg(row0_lane + cx * sx + it * 32) : 0.f;
                    lc_[it] = qv ? __ldg(row1_lane + cx * sx + it * 32) : 0.f;
                    rb_[it] = qv ? __ldg(row0_lane + cx1 * sx + it * 32) : 0.f;
                    rd_[it] = qv ? __ldg(row1_lane + cx1 * sx + it * 32) : 0.f;
                }
            }
            // ground truth of the first tile, requested now, used after the evaluation
            const int Xf = xs + (lane & 7), Yf0 = ys + (lane >> 3), Yf1 = Yf0 + 4;
            const bool okxf = Xf < xe, okf0 = okxf && Yf0 < ye, okf1 = okxf && Yf1 < ye;
            GT g0 = (GT)0, g1 = (GT)0;
            if (hist) {
                if (okf0) g0 = gt_lane[xs];
                if (okf1) g1 = gt_lane[xs + W4];
            }
            if (xs < xe) {
                const int kk[4] = {hA, hB, hC, hD};
                const bool use[4] = {true, kk[1] != kk[0], kk[2] != kk[0] && kk[2] != kk[1], kk[3] != kk[0] && kk[3] != kk[1] && kk[3] != kk[2]};

                int n = 0;
                // the champions' values at the four corners: lane 4c+r fetches corner r of champion c
                __syncwarp();
                if (lane < 16) {
                    const int c = lane >> 2, r = lane & 3;
                    const int k = c == 0 ? kk[0] : c == 1 ? kk[1] : c == 2 ? kk[2] : kk[3];
                    const float* pr = r == 0 ? pA : r == 1 ? pB : r == 2 ? pC : pD;
                    s_champ[lane] = __ldg(pr + k);
                }
                __syncwarp();
                const unsigned long long MINUS1 = pack2(-1.0f, -1.0f);

                // survivors, ascending category order
                const int iters = NQ > 0 ? NQ : (p.Q + 31) >> 5;
#pragma unroll
                for (int it = 0; it < iters; ++it) {
                    const int q0 = it * 32, q = q0 + lane;
                    const bool valid = q < p.Q;
                    float a, bq, c_, d;
                    if (NQ > 0) { a = la_[it]; bq = rb_[it]; c_ = lc_[it]; d = rd_[it]; }
                    else {
                        a = bq = c_ = d = 0.f;
                        if (valid) { a = __ldg(pA + q); bq = __ldg(pB + q); c_ = __ldg(pC + q); d = __ldg(pD + q); }
                    }
                    bool dom = false;
                    const unsigned long long ac = pack2(a, c_), bd = pack2(bq, d);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (use[c]) {                              // warp-uniform
                            // smallest of the champion's four leads over q: RN(champion - q) is >= 0 exactly when champion >= q.
                            // A champion with a smaller index may tie; one with a larger index (or q itself) must lead by the margin.
                            const float4 t4 = reinterpret_cast<const float4*>(s_champ)[c];    // champion c at (A, B, C, D): broadcast read
                            float d0, d1, d2, d3;
                            unpack2(fma2(ac, MINUS1, pack2(t4.x, t4.z)), d0, d1);
                            unpack2(fma2(bd, MINUS1, pack2(t4.y, t4.w)), d2, d3);
                            const float lead = fminf(fminf(d0, d1), fminf(d2, d3));
                            dom = dom || (lead >= (kk[c] < q ? 0.f : margin));
                        }
                    }
                    const bool keep = valid && !dom;
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const int pos = n + __popc(bal & ((1u << lane) - 1u));
                        if (pos < p.cap) { s_val[pos] = make_float4(a, c_, bq, d); s_list[pos] = q; }
                    }
                    n += __popc(bal);
                }
                __syncwarp();

                // evaluate the survivors on 8x8 pixel tiles of the cell: lane = (row lane/8 and +4, column lane%8)
                for (int ty = ys; ty < ye; ty += 8) {
                    for (int tx = xs; tx < xe; tx += 8) {
                        const int X = tx + (lane & 7), Y0 = ty + (lane >> 3), Y1 = Y0 + 4;
                        const bool okx = X < xe, ok0 = okx && Y0 < ye, ok1 = okx && Y1 < ye;
                        const float2 lx = s_lx[min(X, p.W - 1)], la = s_ly[min(Y0, p.H - 1)], lb = s_ly[min(Y1, p.H - 1)];
                        float best0 = -INFINITY, best1 = -INFINITY;
                        int i0 = 0, i1 = 0;
                        if (n <= p.cap) {
                            const unsigned long long LX0 = pack2(lx.x, lx.x), LX1 = pack2(lx.y, lx.y);
                            // walk the survivor list by shared-memory address (the winner is remembered as an address too)
                            const uint32_t first = (uint32_t)__cvta_generic_to_shared(s_val), last = first + (uint32_t)n * 16u;
                            uint32_t w0 = first, w1 = first;
#pragma unroll 2
                            for (uint32_t at = first; at < last; at += 16u) {
                                float4 v;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(at));
                                float tt, uu;
                                unpack2(fma2(LX0, pack2(v.x, v.y), mul2(LX1, pack2(v.z, v.w))), tt, uu);   // t = fma(lx0,A,lx1*B), u = fma(lx0,C,lx1*D)
                                const float v0 = __fmaf_rn(la.x, tt, __fmul_rn(la.y, uu));
                                const float v1 = __fmaf_rn(lb.x, tt, __fmul_rn(lb.y, uu));
                                if (v0 > best0) { best0 = v0; w0 = at; }
                                if (v1 > best1) { best1 = v1; w1 = at; }
                            }
                            i0 = s_list[(w0 - first) >> 4]; i1 = s_list[(w1 - first) >> 4];
                        } else {
                            // more survivors than slots (incoherent cell of a wide-Q image): every category, taps from global memory
                            for (int q = 0; q < p.Q; ++q) {
                                const float tt = lerp_w(lx.x, __ldg(pA + q), lx.y, __ldg(pB + q));
                                const float uu = lerp_w(lx.x, __ldg(pC + q), lx.y, __ldg(pD + q));
                                const float v0 = __fmaf_rn(la.x, tt, __fmul_rn(la.y, uu));
                                const float v1 = __fmaf_rn(lb.x, tt, __fmul_rn(lb.y, uu));
                                if (v0 > best0) { best0 = v0; i0 = q; }
                                if (v1 > best1) { best1 = v1; i1 = q; }
                            }
                        }
                        const int tile_off = (ty - ys) * p.W + tx;    // relative to the lane's base pixel
                        if (lbl_lane) {
                            if (ok0) lbl_lane[tile_off] = (int16_t)i0;
                            if (ok1) lbl_lane[tile_off + W4] = (int16_t)i1;
                        }
                        if (hist) {
                            if (ty != ys || tx != xs) {           // later tiles of a large cell: the ground truth was not requested ahead
                                g0 = ok0 ? gt_lane[tile_off] : (GT)0;
                                g1 = ok1 ? gt_lane[tile_off + W4] : (GT)0;
                            }
                            const int c0 = ok0 ? class_of<GT>(g0, p.n) : -1, c1 = ok1 ? class_of<GT>(g1, p.n) : -1;
                            warp_hist_add(hist, c0 >= 0 ? c0 * p.n + i0 : -1);
                            warp_hist_add(hist, c1 >= 0 ? c1 * p.n + i1 : -1);
                        }
                    }
                }
            }
            // slide right
            hA = hB; hC = hD; hB = hBn; hD = hDn;
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

typedef void (*PrunedKernel)(const DecodeParams);
template <typename GT>
static PrunedKernel pruned_kernel_for_q(int Q) {
    const int nq = (Q + 31) / 32;
    switch (nq) {
        case 1: return decode_pruned_kernel<GT, 1>;
        case 2: return decode_pruned_kernel<GT, 2>;
        case 3: return decode_pruned_kernel<GT, 3>;
        case 4: return decode_pruned_kernel<GT, 4>;
        default: return decode_pruned_kernel<GT, 0>;
    }
}
static PrunedKernel pruned_kernel_for(int gt_dtype, int Q) {
    switch (gt_dtype) {
        case ZUTIS_GT_U8: return pruned_kernel_for_q<uint8_t>(Q);
        case ZUTIS_GT_I16: return pruned_kernel_for_q<int16_t>(Q);
        case ZUTIS_GT_I32: return pruned_kernel_for_q<int32_t>(Q);
        default: return pruned_kernel_for_q<long long>(Q);
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

// workspace of the pruned path: champions [B*h*w] int2 | per-image counters [2*B] int
static size_t decode_workspace_bytes(int B, int h, int w) { return decode_ws_bytes(B, (long)h * w); }

extern "C" size_t zutis_decode_workspace_bytes(int B, int Q, int h, int w, int H, int W) {
    (void)Q; (void)H; (void)W;
    if (B <= 0 || h <= 0 || w <= 0) return 0;
    return decode_workspace_bytes(B, h, w);
}

static int decode_score_impl(const float* logits, long sb, long sq, long sy, long sx,
                             int B, int Q, int h, int w, int H, int W,
                             const void* gt, int gt_dtype, long gt_sb,
                             int16_t* labels, int32_t* hist_partial, int n_classes,
                             int mode, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // the contraction may have left the champions in the workspace already (zutis_gemm_logits_champions)
    const bool champions_ready = (mode & ZUTIS_DECODE_CHAMPIONS_READY) != 0;
    mode &= ~ZUTIS_DECODE_CHAMPIONS_READY;
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
    p.XB = (W + 31) / 32; p.XR = 0; p.QC = 0; p.QS = 0; p.hist_in_smem = 0; p.n_items = 0; p.n_groups = 0; p.vec_stage = 0; p.gt_bytes = gt ? gt_dtype_bytes(gt_dtype) : 0;
    p.select = 0; p.agree_min = 0; p.img_stats = nullptr; p.champ = nullptr; p.lead = nullptr; p.lead_delta = 0; p.cap = 0;
    p.off_ystart = p.off_xstart = p.off_ly = p.off_lx = p.off_img = p.off_warp = 0;

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
    // ---- can the pruned kernel take the coherent images?  (category index contiguous, cells of >= 4x4 pixels, workspace)
    const size_t ws_need = decode_workspace_bytes(B, h, w);
    const long extent_px = (long)(h - 1) * sy + (long)(w - 1) * sx + (long)(Q - 1);
    bool pruned_ok = tiled_ok && sq == 1 && sx > 0 && sy > 0 && H >= 4 * h && W >= 4 * w && Q >= 8 && B <= 1024 &&
                     extent_px < 2147483647L && (long)B * h * w < 2147483647L && workspace != nullptr && workspace_bytes >= ws_need &&
                     (reinterpret_cast<uintptr_t>(workspace) & 7) == 0 && w <= 8192 &&
                     (sx & 3) == 0 && (sy & 3) == 0 && (sb & 3) == 0 && sx >= ((Q + 3) & ~3) && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    if (mode == ZUTIS_DECODE_PRUNED && !pruned_ok)
        return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: pruned kernel needs contiguous categories, >= 4x up-sampling, B <= 1024 and a workspace of %zu bytes (zutis_decode_workspace_bytes)", ws_need);
    const bool use_pruned = pruned_ok && (mode == ZUTIS_DECODE_PRUNED || mode == ZUTIS_DECODE_AUTO);
    const bool use_tiled = use_pruned || (mode == ZUTIS_DECODE_TILED) || (mode == ZUTIS_DECODE_AUTO && tiled_ok);

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
                            (size_t)((H + 1) & ~1) * 8 + (size_t)kTiledWarps * 2 * 2 * XR * p.QS * 4 + (use_pruned ? (size_t)((B + 3) & ~3) * 4 : 0);
        const long extent = (long)(h - 1) * sy + (long)(w - 1) * sx + (long)(Q - 1) * sq;      // per-image offsets stay 32-bit in the kernel
        if (groups > kMaxGroups || H > kMaxTableRows || smem > 200 * 1024 || extent >= 2147483647L || sx < 0 || sy < 0 || sq < 0 ||
            p.n_items * ((Q + p.QC - 1) / p.QC) >= 2147483647L) {
            if (mode == ZUTIS_DECODE_TILED || mode == ZUTIS_DECODE_PRUNED)
                return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: tiled kernel does not fit this shape (groups=%d smem=%zu)", groups, smem);
        } else {
            // ---- pruned path: champions per low-res pixel, then the pruned kernel on the finite + coherent images and the
            // tiled kernel on the rest (each kernel returns at once when it has no image)
            size_t psmem = 0;
            if (use_pruned) {
                p.cap = Q <= 128 ? ((Q + 3) & ~3) : 256;
                p.off_ystart = (p.hist_in_smem ? ((nn + 3) & ~3) : 0) * 4;
                p.off_xstart = p.off_ystart + ((h + 1 + 3) & ~3) * 4;
                p.off_ly = p.off_xstart + ((w + 1 + 3) & ~3) * 4;
                p.off_lx = p.off_ly + ((H + 1) & ~1) * 8;
                p.off_img = p.off_lx + ((W + 1) & ~1) * 8;
                p.off_warp = p.off_img + ((B + 3) & ~3) * 4;  // every offset is a multiple of 16 bytes
                const size_t tables = (size_t)p.off_warp;
                psmem = tables + (size_t)kPrunedWarps * p.cap * 20 + (size_t)kPrunedWarps * 64;
                if (psmem > 100 * 1024) {
                    if (mode == ZUTIS_DECODE_PRUNED)
                        return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: pruned kernel does not fit this shape (smem=%zu)", psmem);
                    psmem = 0;                                   // AUTO: plain tiled launch below
                }
            }
            if (psmem) {
                int* champ = decode_ws_champ(workspace);
                float* lead = decode_ws_lead(workspace, B, (long)h * w);
                int* stats = decode_ws_stats(workspace, B, (long)h * w);
                if (!champions_ready) {
                    ZUTIS_CUDA(cudaMemsetAsync(stats, 0, decode_ws_counter_bytes(B), stream));
                    champion_kernel<<<(unsigned)(B * h), 256, (size_t)w * 4, stream>>>(logits, sb, sy, sx, B, Q, h, w, champ, lead, stats);
                    st = check_launch("champion_kernel");
                    if (st != ZUTIS_OK) return st;
                }
                p.champ = champ; p.lead = lead; p.lead_delta = lead - reinterpret_cast<const float*>(champ); p.img_stats = stats;
                // AUTO: an image is worth pruning when >= 10 % of its horizontally adjacent low-res pixels share their champion
                // (model outputs: ~30 %; i.i.d. noise: 1 %); PRUNED forces every finite image through the pruned kernel
                p.agree_min = mode == ZUTIS_DECODE_PRUNED ? 0 : (int)(((long)h * (w - 1) + 9) / 10);
                p.select = 2;
                PrunedKernel pk = pruned_kernel_for(hist_partial ? gt_dtype : ZUTIS_GT_I64, Q);
                ZUTIS_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
                int pper = 1;
                ZUTIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pper, pk, kPrunedWarps * 32, psmem));
                if (pper < 1) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: pruned kernel does not fit (smem %zu)", psmem);
                long pblocks = ((long)B * h * ((w + kCellRun - 1) / kCellRun) + kPrunedWarps - 1) / kPrunedWarps;
                if (pblocks > (long)sms * pper) pblocks = (long)sms * pper;
                pk<<<(unsigned)pblocks, kPrunedWarps * 32, psmem, stream>>>(p);
                st = check_launch("decode_pruned_kernel");
                if (st != ZUTIS_OK) return st;
                p.select = 1;
            }
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
