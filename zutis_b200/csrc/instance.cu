// Instance-path and on-request kernels (sm_100a):
//   upsample_kernel         F.interpolate(...) materialised for return_logits=True   networks/zutis.py:366-370
//   threshold_kernel        interp(probabilities) > threshold -> bit-packed masks    networks/zutis.py:422-425, :390
//   unpack_bits_kernel      bit-packed -> one byte per pixel for legacy consumers
//   pair_inter_kernel       popcount(mask_i & mask_j): the counts behind compute_iou  utils/iou.py:31-33 (NMS, zutis.py:258-259)
//   mask_rle_kernel         column-major (COCO) run lengths + bounding box of a mask  networks/zutis.py:290, :294
//   lowres_stats_kernel     mask sizes, in-mask probability sums, masked mean tokens  networks/zutis.py:390-406
//   (the tiled threshold kernel lives in decode_score.cu next to the decode kernel whose staging it shares)
//   categories_kernel       sigmoid(T * cos(text, mean token)) -> argmax / max        networks/zutis.py:409-420
// Compiled with -fmad=false; interpolation uses the same explicit fma pattern as decode_score.cu.
#include "common.cuh"
#include "gemm.cuh"

#include <limits.h>

namespace zutis {

struct PlaneParams {
    const float* in;
    long sb, sq, sy, sx;
    int B, Q, h, w, H, W;
    float scale_y, scale_x;
    int identity;
};

__device__ __forceinline__ float sample_bilinear(const PlaneParams& p, const float* plane, int Y, int X) {
    if (p.identity) return __ldg(plane + (long)Y * p.sy + (long)X * p.sx);
    const AxisTap ty = axis_tap(Y, p.h, p.H, p.scale_y);
    const AxisTap tx = axis_tap(X, p.w, p.W, p.scale_x);
    const float a = __ldg(plane + (long)ty.i0 * p.sy + (long)tx.i0 * p.sx);
    const float b = __ldg(plane + (long)ty.i0 * p.sy + (long)tx.i1 * p.sx);
    const float c = __ldg(plane + (long)ty.i1 * p.sy + (long)tx.i0 * p.sx);
    const float d = __ldg(plane + (long)ty.i1 * p.sy + (long)tx.i1 * p.sx);
    return __fmaf_rn(ty.l0, lerp_w(tx.l0, a, tx.l1, b), __fmul_rn(ty.l1, lerp_w(tx.l0, c, tx.l1, d)));
}

__global__ void __launch_bounds__(256) upsample_kernel(const PlaneParams p, float* out) {
    const long total = (long)p.B * p.Q * p.H * p.W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r = i;
        const int X = (int)(r % p.W); r /= p.W;
        const int Y = (int)(r % p.H); r /= p.H;
        const int q = (int)(r % p.Q);
        const int b = (int)(r / p.Q);
        out[i] = sample_bilinear(p, p.in + (long)b * p.sb + (long)q * p.sq, Y, X);
    }
}

// one warp per (mask, output row, 32-pixel word)
__global__ void __launch_bounds__(256) threshold_kernel(const PlaneParams p, float threshold, uint32_t* bits, int* areas) {
    const int words = (p.W + 31) >> 5;
    const long nwords = (long)p.B * p.Q * p.H * words;
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long wi = warp0; wi < nwords; wi += nwarps) {
        long r = wi;
        const int xw = (int)(r % words); r /= words;
        const int Y = (int)(r % p.H); r /= p.H;      // r = b*Q + q
        const int q = (int)(r % p.Q);
        const int b = (int)(r / p.Q);
        const int X = xw * 32 + lane;
        bool on = false;
        if (X < p.W) on = sample_bilinear(p, p.in + (long)b * p.sb + (long)q * p.sq, Y, X) > threshold;
        const unsigned word = __ballot_sync(0xffffffffu, on);
        if (lane == 0) {
            bits[wi] = word;
            if (areas && word) atomicAdd(areas + r, __popc(word));
        }
    }
}

__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint32_t* bits, long n_masks, int H, int W, uint8_t* out) {
    const int words = (W + 31) >> 5;
    const long total = n_masks * H * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int X = (int)(i % W);
        const long row = i / W;
        out[i] = (bits[row * words + (X >> 5)] >> (X & 31)) & 1u;
    }
}

// block (i, j>=i): inter[i][j] = inter[j][i] = sum_w popc(m_i[w] & m_j[w])
__global__ void __launch_bounds__(256) pair_inter_kernel(const uint32_t* bits, int M, long words, int* inter) {
    const int i = blockIdx.y, j = blockIdx.x;
    if (j < i) return;
    const uint32_t* a = bits + (long)i * words;
    const uint32_t* c = bits + (long)j * words;
    int acc = 0;
    for (long k = threadIdx.x; k < words; k += blockDim.x) acc += __popc(a[k] & c[k]);
    __shared__ int red[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < 8; ++k) s += red[k];
        inter[(long)i * M + j] = s;
        inter[(long)j * M + i] = s;
    }
}


// Greedy hard NMS per category (networks/zutis.py:245-278), one block per image, driven by the intersection counts.
// Every category advances at once: in a round, thread i is picked when it is the highest-scoring live candidate of its
// category; the picked mask then suppresses the live masks of its category whose IoU with it exceeds the threshold
// (IoU = both / ((area_i + area_top - both) + 1e-7) in float64, as utils/iou.py:31-33 computes it), and whatever is
// left with a score <= floor is dropped, exactly like the reference's `if s > 0.001` after the first pick.
// pick_rank[i] = the round in which i was picked (-1: suppressed or dropped); scores are unchanged by hard NMS.
// tie[b] is set when two live candidates of a category carry the same score at a pick: the reference's order then
// depends on numpy's unstable argsort, and the caller replays that image on the host.
__global__ void __launch_bounds__(128) nms_hard_kernel(const int* __restrict__ inter, const int* __restrict__ cats, const float* __restrict__ scores,
                                                       int M, double iou_threshold, float floor, int* __restrict__ pick_rank, int* __restrict__ tie) {
    extern __shared__ int s_mem[];
    int* s_cat = s_mem;                       // [M]
    float* s_score = reinterpret_cast<float*>(s_mem + M);
    int* s_live = s_mem + 2 * M;              // [M] 1 = candidate still in play
    int* s_top = s_mem + 3 * M;               // [M] the query picked for this thread's category in the current round (-1: none)
    __shared__ int s_any, s_tie;
    const int b = blockIdx.x;
    const int* in = inter + (long)b * M * M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        s_cat[i] = cats[(long)b * M + i];
        s_score[i] = scores[(long)b * M + i];
        s_live[i] = s_cat[i] != 0;            // the background category is skipped (zutis.py:246)
        pick_rank[(long)b * M + i] = -1;
    }
    if (threadIdx.x == 0) s_tie = 0;
    __syncthreads();
    for (int round = 0; round < M; ++round) {
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        // who is the top of its category?
        for (int i = threadIdx.x; i < M; i += blockDim.x) {
            bool top = s_live[i] != 0;
            if (top) {
                const float si = s_score[i];
                if (si != si) s_tie = 1;      // NaN scores have no order: host replay
                for (int j = 0; j < M; ++j) {
                    if (j == i || !s_live[j] || s_cat[j] != s_cat[i]) continue;
                    const float sj = s_score[j];
                    if (sj > si) { top = false; break; }
                    if (sj == si) { s_tie = 1; if (j < i) { top = false; break; } }      // flagged; any deterministic choice
                }
            }
            s_top[i] = top ? 1 : 0;
            if (top) s_any = 1;
        }
        __syncthreads();
        if (!s_any) break;
        // the picks leave the game; everything else in their category is tested against them
        for (int i = threadIdx.x; i < M; i += blockDim.x) {
            if (s_top[i]) { pick_rank[(long)b * M + i] = round; continue; }
            if (!s_live[i]) continue;
            int t = -1;
            for (int j = 0; j < M; ++j)
                if (s_top[j] && s_cat[j] == s_cat[i]) { t = j; break; }
            if (t < 0) continue;
            const int both = in[(long)i * M + t];
            const double iou = (double)both / ((double)(in[(long)i * M + i] + in[(long)t * M + t] - both) + 1e-7);
            const float s_new = iou > iou_threshold ? 0.0f : s_score[i];                  // s * weight, weight in {0, 1}
            if (!(s_new > floor)) s_live[i] = 0;                                           // `if s > 0.001` (zutis.py:274)
        }
        __syncthreads();
        for (int i = threadIdx.x; i < M; i += blockDim.x)
            if (s_top[i]) s_live[i] = 0;
        __syncthreads();
    }
    if (threadIdx.x == 0) tie[b] = s_tie;
}

// COCO run-length encoding and bounding box of bit-packed masks, on the device.
// Replaces, per kept mask, pycocotools.mask.encode(np.asfortranarray(m)) (networks/zutis.py:290; the run lengths that
// cocoapi's rleEncode produces: the mask flattened COLUMN by column, alternating runs starting with a zero run that
// may be empty) and torchvision.ops.masks_to_boxes (:294).  The reference ships every boolean mask to the host
// (307 KB each at 480x640) and walks it there.
// One block per mask:
//   1. 32x32 bit-block transposes by warp ballots turn the row-packed mask (bit x of word [y][x/32]) into
//      column-packed words in shared memory (bit y of word [x][y/32]);
//   2. thread = column: t = c ^ ((c << 1) | carry) marks the positions where the column-major bit stream changes
//      value (the carry into a column is the last bit of the previous column, 0 before the first); popcounts give the
//      transitions per column, first/last set bits the box;
//   3. a block scan turns the counts into output offsets and carries the last transition position across columns;
//   4. (write pass) the set bits of t are walked again and every transition position p_k emits run k = p_k - p_(k-1),
//      p_(-1) = 0; the closing run is H*W - p_(K-1).  A mask with K transitions has K+1 runs.
// The caller runs the kernel twice: runs == nullptr counts (n_runs, boxes), then with exact offsets it writes.
constexpr int kRleThreads = 256;

__global__ void __launch_bounds__(kRleThreads) mask_rle_kernel(const uint32_t* __restrict__ bits, long mask_stride,
                                                               const int* __restrict__ mask_ids, int H, int W, int words,
                                                               int hw, int hwp, const long* __restrict__ run_offsets,
                                                               uint32_t* __restrict__ runs, int* __restrict__ n_runs,
                                                               int* __restrict__ boxes) {
    extern __shared__ uint32_t s_col[];                      // [words*32][hwp] column-packed bits, hwp odd
    const int wpad = words * 32;
    int* s_cnt = reinterpret_cast<int*>(s_col + (size_t)wpad * hwp);    // [wpad] transitions per column -> exclusive offsets
    int* s_last = s_cnt + wpad;                              // [wpad] last transition position of the column -> of all earlier columns
    __shared__ int s_box[4];
    __shared__ int s_warp_sum[kRleThreads / 32], s_warp_max[kRleThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long mask = mask_ids ? mask_ids[blockIdx.x] : (long)blockIdx.x;
    const uint32_t* m = bits + mask * mask_stride;
    if (threadIdx.x == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MAX; s_box[2] = -1; s_box[3] = -1; }

    // ---- 1. transpose
    const uint32_t tail_x = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
    for (int blk = warp; blk < hw * words; blk += kRleThreads / 32) {
        const int yb = blk / words, xb = blk % words;
        const int y = yb * 32 + lane;
        uint32_t r = y < H ? __ldg(m + (long)y * words + xb) : 0u;
        if (xb == words - 1) r &= tail_x;
        uint32_t mine = 0;
#pragma unroll
        for (int x = 0; x < 32; ++x) {
            const uint32_t v = __ballot_sync(0xffffffffu, (r >> x) & 1u);
            if (lane == x) mine = v;
        }
        s_col[(size_t)(xb * 32 + lane) * hwp + yb] = mine;
    }
    __syncthreads();

    // ---- 2. transitions per column, box
    const uint32_t tail_y = (H & 31) ? ((1u << (H & 31)) - 1u) : 0xffffffffu;
    int xmin = INT_MAX, ymin = INT_MAX, xmax = -1, ymax = -1;
    for (int x = threadIdx.x; x < wpad; x += kRleThreads) {
        int cnt = 0, last = -1;
        if (x < W) {
            uint32_t carry = x > 0 ? (s_col[(size_t)(x - 1) * hwp + ((H - 1) >> 5)] >> ((H - 1) & 31)) & 1u : 0u;
            for (int j = 0; j < hw; ++j) {
                const uint32_t c = s_col[(size_t)x * hwp + j];
                uint32_t t = c ^ ((c << 1) | carry);
                if (j == hw - 1) t &= tail_y;
                carry = c >> 31;
                cnt += __popc(t);
                if (t) last = x * H + j * 32 + (31 - __clz(t));
                if (c) {
                    ymin = min(ymin, j * 32 + __ffs(c) - 1);
                    ymax = max(ymax, j * 32 + 31 - __clz(c));
                    xmin = min(xmin, x);
                    xmax = max(xmax, x);
                }
            }
        }
        s_cnt[x] = cnt;
        s_last[x] = last;
    }
    if (xmax >= 0) { atomicMin(&s_box[0], xmin); atomicMin(&s_box[1], ymin); atomicMax(&s_box[2], xmax); atomicMax(&s_box[3], ymax); }
    __syncthreads();

    // ---- 3. exclusive scan of the counts / running maximum of the last positions; thread = a contiguous segment of columns
    const int seg = (wpad + kRleThreads - 1) / kRleThreads;
    const int x0 = threadIdx.x * seg, x1 = min(x0 + seg, wpad);
    int sum = 0, mx = -1;
    for (int x = x0; x < x1; ++x) { sum += s_cnt[x]; mx = max(mx, s_last[x]); }
    int inc = sum, incmax = mx;                               // inclusive scans across the warp
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc, o), bmax = __shfl_up_sync(0xffffffffu, incmax, o);
        if (lane >= o) { inc += a; incmax = max(incmax, bmax); }
    }
    if (lane == 31) { s_warp_sum[warp] = inc; s_warp_max[warp] = incmax; }
    __syncthreads();
    int base = 0, basemax = -1, total = 0;
    for (int k = 0; k < kRleThreads / 32; ++k) {
        if (k < warp) { base += s_warp_sum[k]; basemax = max(basemax, s_warp_max[k]); }
        total += s_warp_sum[k];
    }
    int ex = base + inc - sum;                                // transitions before this thread's segment
    int exmax = max(basemax, __shfl_up_sync(0xffffffffu, incmax, 1));
    if (lane == 0) exmax = basemax;
    for (int x = x0; x < x1; ++x) {
        const int c = s_cnt[x], l = s_last[x];
        s_cnt[x] = ex; s_last[x] = exmax;
        ex += c; exmax = max(exmax, l);
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        n_runs[blockIdx.x] = total + 1;
        if (boxes) {
            const bool any = s_box[2] >= 0;
            boxes[4 * blockIdx.x + 0] = any ? s_box[0] : -1; boxes[4 * blockIdx.x + 1] = any ? s_box[1] : -1;
            boxes[4 * blockIdx.x + 2] = s_box[2]; boxes[4 * blockIdx.x + 3] = s_box[3];
        }
    }
    if (!runs) return;

    // ---- 4. emit the runs
    uint32_t* out = runs + run_offsets[blockIdx.x];
    for (int x = threadIdx.x; x < W; x += kRleThreads) {
        int k = s_cnt[x];
        int prev = max(s_last[x], 0);
        uint32_t carry = x > 0 ? (s_col[(size_t)(x - 1) * hwp + ((H - 1) >> 5)] >> ((H - 1) & 31)) & 1u : 0u;
        for (int j = 0; j < hw; ++j) {
            const uint32_t c = s_col[(size_t)x * hwp + j];
            uint32_t t = c ^ ((c << 1) | carry);
            if (j == hw - 1) t &= tail_y;
            carry = c >> 31;
            while (t) {
                const int b = __ffs(t) - 1;
                t &= t - 1;
                const int pos = x * H + j * 32 + b;
                out[k++] = (uint32_t)(pos - prev);
                prev = pos;
            }
        }
    }
    // closing run: everything after the last transition (s_last of a virtual column W = running max over all columns)
    if (threadIdx.x == kRleThreads - 1) {
        // ex / exmax of the last thread cover every column after its loop above
        out[total] = (uint32_t)(H * W - max(exmax, 0));
    }
}

// cocoapi rleToString on the device (the `counts` bytes pycocotools.mask.encode returns, networks/zutis.py:290).
// From the fourth run of a mask on, the value written is the difference to the run two places back; a value becomes
// little-endian groups of 5 bits, bit 5 = "more groups follow", +48; a group with bit 4 set ends the number once the
// remaining sign-extended value is -1.  One block per mask: thread = a contiguous chunk of runs; sweep 1 counts the
// characters, a block scan places the chunks, the block reserves its span of the output with one atomic on a cursor
// (so the strings are packed back to back in completion order), sweep 2 writes.
__device__ __forceinline__ int rle_chars(long long x, uint8_t* out) {
    int n = 0;
    bool more;
    do {
        int c = (int)(x & 0x1f);
        x >>= 5;
        more = (c & 0x10) ? (x != -1) : (x != 0);
        if (more) c |= 0x20;
        if (out) out[n] = (uint8_t)(c + 48);
        ++n;
    } while (more);
    return n;
}

__global__ void __launch_bounds__(kRleThreads) rle_string_kernel(const uint32_t* __restrict__ runs, const long* __restrict__ run_offsets,
                                                                 const int* __restrict__ n_runs, uint8_t* __restrict__ strings,
                                                                 long capacity, unsigned long long* cursor,
                                                                 long* __restrict__ string_offsets, int* __restrict__ string_lengths) {
    __shared__ int s_warp[kRleThreads / 32];
    __shared__ long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = n_runs[blockIdx.x];
    const uint32_t* r = runs + run_offsets[blockIdx.x];
    const int chunk = (n + kRleThreads - 1) / kRleThreads;
    const int i0 = min(threadIdx.x * chunk, n), i1 = min(i0 + chunk, n);
    int len = 0;
    for (int i = i0; i < i1; ++i) len += rle_chars((long long)r[i] - (i > 2 ? (long long)r[i - 2] : 0ll), nullptr);
    int inc = len;
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int before = 0, total = 0;
    for (int k = 0; k < kRleThreads / 32; ++k) {
        if (k < warp) before += s_warp[k];
        total += s_warp[k];
    }
    if (threadIdx.x == 0) {
        const long base = (long)atomicAdd(cursor, (unsigned long long)total);
        s_base = base;
        string_offsets[blockIdx.x] = base;
        string_lengths[blockIdx.x] = total;
    }
    __syncthreads();
    if (s_base + total > capacity) return;                   // reported through the cursor; the caller sized the buffer
    uint8_t* out = strings + s_base + before + inc - len;
    for (int i = i0; i < i1; ++i) out += rle_chars((long long)r[i] - (i > 2 ? (long long)r[i - 2] : 0ll), out);
}

// Low-resolution instance statistics (networks/zutis.py:390-406).  One block per (image, tile of kStatQ queries,
// slice of kStatD channels):
//   phase 1  threads sweep the h*w probabilities of the tile's queries once: per pixel a 16-bit word of ">" bits goes
//            to shared memory, per query the mask size and the in-mask probability sum are block-reduced (fixed tree);
//   phase 2  the image's tokens are streamed once per query TILE (not once per query): thread = one channel and one
//            of 4 pixel phases, the token value is added to the accumulators of the queries whose bit is set; the 4
//            phases are then combined in a fixed order (deterministic).
// The reference materialises this as a [B,100,h,w,512] broadcast (0.98 GB per image).
constexpr int kStatQ = 16;
constexpr int kStatD = 64;

__global__ void __launch_bounds__(256) lowres_stats_kernel(const float* probs, long sb, long sq, long sy, long sx,
                                                           const float* tokens, int Q, int h, int w, int D, float threshold,
                                                           int* sizes, float* psum, float* mean_tokens) {
    extern __shared__ unsigned short s_bits[];     // [h*w] one bit per query of the tile
    __shared__ int r_cnt[8][kStatQ];
    __shared__ float r_sum[8][kStatQ];
    __shared__ float s_denom[kStatQ];
    __shared__ float s_part[3][kStatQ][kStatD];    // partial sums of pixel phases 1..3
    const int tiles_per_image = (Q + kStatQ - 1) / kStatQ;
    const int b = blockIdx.x / tiles_per_image;
    const int q0 = (blockIdx.x % tiles_per_image) * kStatQ;
    const int nq = min(kStatQ, Q - q0);
    const int hw = h * w;
    const float* img = probs + (long)b * sb + (long)q0 * sq;
    {
        int cnt[kStatQ];
        float sum[kStatQ];
#pragma unroll
        for (int j = 0; j < kStatQ; ++j) { cnt[j] = 0; sum[j] = 0.0f; }
        for (int i = threadIdx.x; i < hw; i += blockDim.x) {
            const float* px = img + (long)(i / w) * sy + (long)(i % w) * sx;
            unsigned bits = 0;
#pragma unroll
            for (int j = 0; j < kStatQ; ++j) {
                if (j < nq) {
                    const float v = px[(long)j * sq];
                    if (v > threshold) { bits |= 1u << j; ++cnt[j]; sum[j] = __fadd_rn(sum[j], v); }
                }
            }
            s_bits[i] = (unsigned short)bits;
        }
#pragma unroll
        for (int j = 0; j < kStatQ; ++j) {
            for (int o = 16; o > 0; o >>= 1) {
                cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], o);
                sum[j] = __fadd_rn(sum[j], __shfl_xor_sync(0xffffffffu, sum[j], o));
            }
            if ((threadIdx.x & 31) == 0) { r_cnt[threadIdx.x >> 5][j] = cnt[j]; r_sum[threadIdx.x >> 5][j] = sum[j]; }
        }
    }
    __syncthreads();
    if (threadIdx.x < nq) {
        int total = 0;
        float tsum = 0.0f;
        for (int k = 0; k < 8; ++k) { total += r_cnt[k][threadIdx.x]; tsum = __fadd_rn(tsum, r_sum[k][threadIdx.x]); }
        if (blockIdx.y == 0) {
            sizes[(long)b * Q + q0 + threadIdx.x] = total;
            psum[(long)b * Q + q0 + threadIdx.x] = tsum;
        }
        s_denom[threadIdx.x] = __fadd_rn((float)total, 1e-7f);    // (mask_sizes + 1e-7) promotes to fp32, zutis.py:406
    }
    __syncthreads();
    if (!mean_tokens) return;
    const int c = threadIdx.x & (kStatD - 1);      // channel within the slice
    const int phase = threadIdx.x / kStatD;        // 0..3: pixels i = phase, phase+4, ...
    const int d = blockIdx.y * kStatD + c;
    const float* tok = tokens + (long)b * hw * D + d;
    float acc[kStatQ];
#pragma unroll
    for (int j = 0; j < kStatQ; ++j) acc[j] = 0.0f;
    if (d < D) {
        // batches of 8 pixels: all token loads of a batch are issued before any is consumed (the loop is otherwise
        // one exposed global-load latency per pixel)
        for (int i0 = phase; i0 < hw; i0 += 32) {
            float v[8];
            unsigned bits[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + 4 * k;
                bits[k] = i < hw ? s_bits[i] : 0u;
                v[k] = bits[k] ? __ldg(tok + (long)i * D) : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (bits[k] == 0) continue;                      // warp-uniform: a warp shares its pixel phase
#pragma unroll
                for (int j = 0; j < kStatQ; ++j)
                    if (bits[k] & (1u << j)) acc[j] = __fadd_rn(acc[j], v[k]);
            }
        }
    }
    if (phase > 0) {
#pragma unroll
        for (int j = 0; j < kStatQ; ++j) s_part[phase - 1][j][c] = acc[j];
    }
    __syncthreads();
    if (phase == 0 && d < D) {
#pragma unroll
        for (int j = 0; j < kStatQ; ++j) {
            if (j < nq) {
                const float total = __fadd_rn(__fadd_rn(acc[j], s_part[0][j][c]), __fadd_rn(s_part[1][j][c], s_part[2][j][c]));
                mean_tokens[((long)b * Q + q0 + j) * D + d] = total / s_denom[j];
            }
        }
    }
}


// ------------------------------------------------------------------ masked average on the tensor cores
// sum_hw tokens[b,hw,:] * mask[b,q,hw] (zutis.py:404-406, a [B,100,h,w,512] broadcast in the reference) is a
// [Q, hw] x [hw, D] contraction per image.  The 0/1 mask is exact in any number format and the tokens go through the
// contraction kernel's hi/lo split, so the sums are fp32-grade.  Both operands must be K-contiguous (K = pixels):
// the mask is written that way, the channel-last tokens are transposed once.
__global__ void __launch_bounds__(256) mask_matrix_kernel(const float* probs, long sb, long sq, long sy, long sx, int Q, int h, int w,
                                                          int hwp, float threshold, float* mask) {
    const int b = blockIdx.z, q = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hwp) return;
    float m = 0.0f;
    if (i < h * w) m = probs[(long)b * sb + (long)q * sq + (long)(i / w) * sy + (long)(i % w) * sx] > threshold ? 1.0f : 0.0f;
    mask[((long)b * Q + q) * hwp + i] = m;
}

// tokens [B, hw, D] -> [B, D, hwp] (zero-padded pixels), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_tokens_kernel(const float* tokens, int hw, int D, int hwp, float* out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int pp = p0 + r, d = d0 + tx;
        tile[r][tx] = (pp < hw && d < D) ? tokens[((long)b * hw + pp) * D + d] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, pp = p0 + tx;
        if (d < D && pp < hwp) out[((long)b * D + d) * hwp + pp] = tile[tx][r];
    }
}

__global__ void __launch_bounds__(256) divide_rows_kernel(float* sums, const int* sizes, long n_rows, int D) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows * D) sums[i] = sums[i] / __fadd_rn((float)sizes[i / D], 1e-7f);       // (mask_sizes + 1e-7), zutis.py:406
}

// block per (b,q): cosine of the mean token with every text row, sigmoid(T*cos), first-max category
// (networks/zutis.py:409-420).  A warp takes one category at a time: lanes stride the channel dimension (coalesced
// reads of the text row), partial sums are combined by a shuffle tree.
__global__ void __launch_bounds__(128) categories_kernel(const float* mean_tokens, const float* text, int n_cat, int D,
                                                         float temperature, int* category, float* max_prob) {
    extern __shared__ float s_tok[];               // [D] normalised token, then [n_cat] probabilities
    float* s_prob = s_tok + D;
    const int bq = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* t = mean_tokens + (long)bq * D;
    float ss = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { const float v = t[d]; s_tok[d] = v; ss = __fmaf_rn(v, v, ss); }
    __shared__ float red[4];
    for (int o = 16; o > 0; o >>= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    const float norm = __fadd_rn(sqrtf(__fadd_rn(__fadd_rn(red[0], red[1]), __fadd_rn(red[2], red[3]))), 1e-7f);   // zutis.py:412
    for (int d = threadIdx.x; d < D; d += blockDim.x) s_tok[d] = s_tok[d] / norm;
    __syncthreads();
    for (int n = warp; n < n_cat; n += 4) {
        const float* e = text + (long)n * D;
        float acc = 0.0f;
        for (int d = lane; d < D; d += 32) acc = __fmaf_rn(__ldg(e + d), s_tok[d], acc);
        for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) s_prob[n] = 1.0f / (1.0f + expf(-__fmul_rn(acc, temperature)));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = s_prob[0];
        int idx = 0;
        for (int n = 1; n < n_cat; ++n)
            if (better_nan_aware(s_prob[n], best)) { best = s_prob[n]; idx = n; }
        category[bq] = idx;
        max_prob[bq] = best;
    }
}

static int make_plane_params(PlaneParams* p, const float* in, long sb, long sq, long sy, long sx,
                             int B, int Q, int h, int w, int H, int W, const char* who) {
    if (!in) return fail(ZUTIS_ERR_BAD_ARG, "%s: input is NULL", who);
    if (!(B > 0 && Q > 0 && h > 0 && w > 0 && H > 0 && W > 0))
        return fail(ZUTIS_ERR_BAD_ARG, "%s: non-positive shape B=%d Q=%d h=%d w=%d H=%d W=%d", who, B, Q, h, w, H, W);
    p->in = in; p->sb = sb; p->sq = sq; p->sy = sy; p->sx = sx;
    p->B = B; p->Q = Q; p->h = h; p->w = w; p->H = H; p->W = W;
    p->scale_y = axis_scale(h, H); p->scale_x = axis_scale(w, W);
    p->identity = (H == h && W == w);
    return current_device_ok();
}

static unsigned grid_for(long work_items, int per_block) {
    long blocks = (work_items + per_block - 1) / per_block;
    const long cap = (long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace zutis

using namespace zutis;

extern "C" int zutis_upsample_bilinear(const float* in, long sb, long sq, long sy, long sx,
                                       int B, int Q, int h, int w, int H, int W, float* out, void* stream) {
    PlaneParams p;
    int st = make_plane_params(&p, in, sb, sq, sy, sx, B, Q, h, w, H, W, "zutis_upsample_bilinear");
    if (st != ZUTIS_OK) return st;
    ZUTIS_REQUIRE(out != nullptr, "zutis_upsample_bilinear: out is NULL");
    upsample_kernel<<<grid_for((long)B * Q * H * W, 256), 256, 0, (cudaStream_t)stream>>>(p, out);
    return check_launch("upsample_kernel");
}

extern "C" int zutis_decode_threshold(const float* probs, long sb, long sq, long sy, long sx,
                                      int B, int Q, int h, int w, int H, int W, float threshold,
                                      uint32_t* mask_bits, int32_t* areas, void* stream) {
    PlaneParams p;
    int st = make_plane_params(&p, probs, sb, sq, sy, sx, B, Q, h, w, H, W, "zutis_decode_threshold");
    if (st != ZUTIS_OK) return st;
    ZUTIS_REQUIRE(mask_bits != nullptr, "zutis_decode_threshold: mask_bits is NULL");
    {
        // fast path: warp-tile kernel with staged taps (up-sampling by >= ~5x); anything else takes the generic kernel
        const int tiled = launch_threshold_tiled(probs, sb, sq, sy, sx, B, Q, h, w, H, W, threshold, mask_bits, areas, (cudaStream_t)stream);
        if (tiled != ZUTIS_ERR_UNSUPPORTED) return tiled;
    }
    const long nwords = (long)B * Q * H * ((W + 31) / 32);
    threshold_kernel<<<grid_for(nwords, 8), 256, 0, (cudaStream_t)stream>>>(p, threshold, mask_bits, areas);
    return check_launch("threshold_kernel");
}

extern "C" int zutis_unpack_mask_bits(const uint32_t* mask_bits, long n_masks, int H, int W,
                                      uint8_t* out_bytes, void* stream) {
    ZUTIS_REQUIRE(mask_bits && out_bytes, "zutis_unpack_mask_bits: NULL pointer");
    ZUTIS_REQUIRE(n_masks >= 0 && H > 0 && W > 0, "zutis_unpack_mask_bits: bad shape");
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    if (n_masks == 0) return ZUTIS_OK;
    unpack_bits_kernel<<<grid_for(n_masks * H * W, 256), 256, 0, (cudaStream_t)stream>>>(mask_bits, n_masks, H, W, out_bytes);
    return check_launch("unpack_bits_kernel");
}

extern "C" int zutis_pairwise_mask_intersections(const uint32_t* mask_bits, int M, long words_per_mask,
                                                 int32_t* inter, void* stream) {
    ZUTIS_REQUIRE(mask_bits && inter, "zutis_pairwise_mask_intersections: NULL pointer");
    ZUTIS_REQUIRE(M > 0 && M <= 65535 && words_per_mask > 0, "zutis_pairwise_mask_intersections: bad shape M=%d words=%ld", M, words_per_mask);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    pair_inter_kernel<<<dim3(M, M), 256, 0, (cudaStream_t)stream>>>(mask_bits, M, words_per_mask, inter);
    return check_launch("pair_inter_kernel");
}


extern "C" int zutis_instance_nms_hard(const int32_t* inter, const int32_t* categories, const float* scores, int B, int M,
                                       double iou_threshold, float score_floor, int32_t* pick_rank, int32_t* tie, void* stream) {
    ZUTIS_REQUIRE(inter && categories && scores && pick_rank && tie, "zutis_instance_nms_hard: NULL pointer");
    ZUTIS_REQUIRE(B > 0 && M > 0 && M <= 8192, "zutis_instance_nms_hard: bad shape B=%d M=%d", B, M);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    // the reference compares `iou > 0.3` with both sides in float64, and `s > 0.001` in float32
    nms_hard_kernel<<<(unsigned)B, 128, (size_t)M * 16, (cudaStream_t)stream>>>(inter, categories, scores, M, iou_threshold, score_floor, pick_rank, tie);
    return check_launch("nms_hard_kernel");
}

extern "C" int zutis_mask_rle(const uint32_t* mask_bits, long mask_stride_words, const int32_t* mask_ids, int n_masks,
                              int H, int W, const int64_t* run_offsets, uint32_t* runs, int32_t* n_runs, int32_t* boxes,
                              void* stream) {
    ZUTIS_REQUIRE(mask_bits && n_runs, "zutis_mask_rle: NULL pointer");
    ZUTIS_REQUIRE(n_masks >= 0 && H > 0 && W > 0, "zutis_mask_rle: bad shape");
    ZUTIS_REQUIRE((long)H * W < 2147483647L, "zutis_mask_rle: H*W=%ld does not fit 31 bits", (long)H * W);
    ZUTIS_REQUIRE(!runs || run_offsets, "zutis_mask_rle: the write pass needs run_offsets");
    static_assert(sizeof(long) == sizeof(int64_t), "LP64 expected");
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    if (n_masks == 0) return ZUTIS_OK;
    const int words = (W + 31) / 32, hw = (H + 31) / 32, hwp = hw | 1;
    const size_t smem = ((size_t)words * 32 * hwp + 2 * (size_t)words * 32) * 4;
    if (smem > 200 * 1024)
        return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_mask_rle: a %dx%d mask needs %zu bytes of shared memory (limit 200 KB)", H, W, smem);
    if (smem > 48 * 1024)
        ZUTIS_CUDA(cudaFuncSetAttribute(mask_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mask_rle_kernel<<<(unsigned)n_masks, kRleThreads, smem, (cudaStream_t)stream>>>(
        mask_bits, mask_stride_words, mask_ids, H, W, words, hw, hwp, reinterpret_cast<const long*>(run_offsets), runs, n_runs, boxes);
    return check_launch("mask_rle_kernel");
}

extern "C" int zutis_rle_to_string(const uint32_t* runs, const int64_t* run_offsets, const int32_t* n_runs, int n_masks,
                                   uint8_t* strings, int64_t capacity, uint64_t* cursor, int64_t* string_offsets,
                                   int32_t* string_lengths, void* stream) {
    ZUTIS_REQUIRE(runs && run_offsets && n_runs && strings && cursor && string_offsets && string_lengths, "zutis_rle_to_string: NULL pointer");
    ZUTIS_REQUIRE(n_masks >= 0 && capacity >= 0, "zutis_rle_to_string: bad shape");
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    if (n_masks == 0) return ZUTIS_OK;
    rle_string_kernel<<<(unsigned)n_masks, kRleThreads, 0, (cudaStream_t)stream>>>(
        runs, reinterpret_cast<const long*>(run_offsets), n_runs, strings, (long)capacity,
        reinterpret_cast<unsigned long long*>(cursor), reinterpret_cast<long*>(string_offsets), string_lengths);
    return check_launch("rle_string_kernel");
}

extern "C" int zutis_instance_lowres_stats(const float* probs, long sb, long sq, long sy, long sx,
                                           const float* tokens, int B, int Q, int h, int w, int D,
                                           float threshold, int32_t* sizes, float* psum, float* mean_tokens,
                                           void* stream) {
    ZUTIS_REQUIRE(probs && sizes && psum, "zutis_instance_lowres_stats: NULL pointer");
    ZUTIS_REQUIRE(B > 0 && Q > 0 && h > 0 && w > 0, "zutis_instance_lowres_stats: bad shape");
    ZUTIS_REQUIRE(!mean_tokens || (tokens && D > 0), "zutis_instance_lowres_stats: mean_tokens needs tokens and D");
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    const size_t smem = (size_t)h * w * 2;
    ZUTIS_REQUIRE(smem <= 180 * 1024, "zutis_instance_lowres_stats: h*w=%d too large", h * w);
    if (smem > 48 * 1024)
        ZUTIS_CUDA(cudaFuncSetAttribute(lowres_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 blocks((unsigned)(B * ((Q + kStatQ - 1) / kStatQ)), mean_tokens ? (unsigned)((D + kStatD - 1) / kStatD) : 1u);
    lowres_stats_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(probs, sb, sq, sy, sx, tokens, Q, h, w, D, threshold,
                                                                       sizes, psum, mean_tokens);
    return check_launch("lowres_stats_kernel");
}


extern "C" size_t zutis_instance_stats_workspace_bytes(int B, int Q, int h, int w, int D) {
    if (B <= 0 || Q <= 0 || h <= 0 || w <= 0 || D <= 0) return 0;
    const long hwp = ((long)h * w + 31) & ~31L;
    const size_t mask = (size_t)B * Q * hwp * 4, tok = (size_t)B * D * hwp * 4;
    const size_t gemm = gemm_tcgen05_workspace_bytes(Q, D, (int)hwp, B, ZUTIS_GEMM_TF32X3);
    return ((mask + 255) & ~(size_t)255) + ((tok + 255) & ~(size_t)255) + gemm;
}

extern "C" int zutis_instance_lowres_stats_ws(const float* probs, long sb, long sq, long sy, long sx,
                                              const float* tokens, int B, int Q, int h, int w, int D,
                                              float threshold, int32_t* sizes, float* psum, float* mean_tokens,
                                              void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // sizes and in-mask probability sums: the statistics kernel without its token phase
    int st = zutis_instance_lowres_stats(probs, sb, sq, sy, sx, nullptr, B, Q, h, w, 0, threshold, sizes, psum, nullptr, stream_);
    if (st != ZUTIS_OK || !mean_tokens) return st;
    ZUTIS_REQUIRE(tokens && D > 0, "zutis_instance_lowres_stats_ws: mean_tokens needs tokens and D");
    const long hw = (long)h * w, hwp = (hw + 31) & ~31L;
    GemmParams g;
    const size_t mask_bytes = ((size_t)B * Q * hwp * 4 + 255) & ~(size_t)255, tok_bytes = ((size_t)B * D * hwp * 4 + 255) & ~(size_t)255;
    const size_t need = zutis_instance_stats_workspace_bytes(B, Q, h, w, D);
    float* mask = reinterpret_cast<float*>(workspace);
    float* tokT = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + mask_bytes);
    g.A = mask; g.lda = hwp; g.strideA = (long)Q * hwp;
    g.Bm = tokT; g.ldb = hwp; g.strideB = (long)D * hwp;
    g.C = mean_tokens; g.stride_cn = D; g.stride_cp = 1; g.strideC = (long)Q * D;
    g.M = Q; g.N = D; g.K = (int)hwp; g.sigmoid = 0;
    const bool tensor_path = workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
                             hwp < 2147483647L && B <= 65535 && Q <= 65535 && gemm_tcgen05_supports(g, B, ZUTIS_GEMM_TF32X3);
    if (!tensor_path)      // no (or too small a) workspace, or a shape the contraction kernel does not take: the SIMT kernel
        return zutis_instance_lowres_stats(probs, sb, sq, sy, sx, tokens, B, Q, h, w, D, threshold, sizes, psum, mean_tokens, stream_);
    mask_matrix_kernel<<<dim3((unsigned)((hwp + 255) / 256), (unsigned)Q, (unsigned)B), 256, 0, stream>>>(probs, sb, sq, sy, sx, Q, h, w, (int)hwp, threshold, mask);
    st = check_launch("mask_matrix_kernel");
    if (st != ZUTIS_OK) return st;
    transpose_tokens_kernel<<<dim3((unsigned)(hwp / 32), (unsigned)((D + 31) / 32), (unsigned)B), 256, 0, stream>>>(tokens, (int)hw, D, (int)hwp, tokT);
    st = check_launch("transpose_tokens_kernel");
    if (st != ZUTIS_OK) return st;
    st = launch_gemm_tcgen05(g, B, ZUTIS_GEMM_TF32X3, reinterpret_cast<char*>(workspace) + mask_bytes + tok_bytes,
                             workspace_bytes - mask_bytes - tok_bytes, stream);
    if (st != ZUTIS_OK) return st;
    const long n = (long)B * Q * D;
    divide_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(mean_tokens, sizes, (long)B * Q, D);
    return check_launch("divide_rows_kernel");
}

extern "C" int zutis_instance_categories(const float* mean_tokens, long n_rows, const float* text, int n_categories, int D,
                                         float temperature, int32_t* category, float* max_prob, void* stream) {
    ZUTIS_REQUIRE(mean_tokens && text && category && max_prob, "zutis_instance_categories: NULL pointer");
    ZUTIS_REQUIRE(n_rows > 0 && n_categories > 0 && D > 0, "zutis_instance_categories: bad shape");
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    const size_t smem = (size_t)(D + n_categories) * sizeof(float);
    ZUTIS_REQUIRE(smem <= 200 * 1024, "zutis_instance_categories: D + n_categories too large");
    if (smem > 48 * 1024)
        ZUTIS_CUDA(cudaFuncSetAttribute(categories_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    categories_kernel<<<(unsigned)n_rows, 128, smem, (cudaStream_t)stream>>>(mean_tokens, text, n_categories, D, temperature, category, max_prob);
    return check_launch("categories_kernel");
}
