// Shared declarations of the decode kernels (decode_score.cu: generic + tiled + threshold kernels and the dispatcher;
// decode_pruned.cu: champion pass and the candidate-pruning kernel).
#pragma once
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

// smallest d in [0,out] whose first tap index is >= target
__device__ __forceinline__ int first_dst_with_tap_ge(int target, int in, int out, float scale) {
    int lo = 0, hi = out;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (axis_tap(mid, in, out, scale).i0 >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
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

template <typename GT>
__device__ __forceinline__ int class_of(GT g, int n) {
    return (g >= (GT)0 && (long long)g < (long long)n) ? (int)g : -1;
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

// Add a CTA's shared-memory histogram into the global int32 partial.  Two neighbouring bins travel in ONE 64-bit atomic
// when the partial is 8-byte aligned: counts are non-negative and the whole partial stays below 2^31 (the entry point
// checks B*H*W), so the low word never carries into the high one.  All CTAs flush at about the same time, which makes
// this a burst of (#CTAs x non-zero bins) L2 atomics; halving it is measurable.
__device__ __forceinline__ void flush_shared_hist(const int* s_hist, int* hist, int nn) {
    if ((reinterpret_cast<uintptr_t>(hist) & 7) == 0) {
        const int pairs = nn >> 1;
        for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
            const int2 v = reinterpret_cast<const int2*>(s_hist)[i];
            if (v.x | v.y)
                atomicAdd(reinterpret_cast<unsigned long long*>(hist) + i,
                          (unsigned long long)(unsigned)v.x | ((unsigned long long)(unsigned)v.y << 32));
        }
        if ((nn & 1) && threadIdx.x == 0 && s_hist[nn - 1]) atomicAdd(hist + nn - 1, s_hist[nn - 1]);
    } else {
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            const int v = s_hist[i];
            if (v) atomicAdd(hist + i, v);
        }
    }
}

// decode_pruned.cu.  Sets up and launches the pruned kernel (and the champion pass unless the workspace already
// holds the champions of these logits) for the images it owns; on success *launched = true and p.select = 1, so that
// the tiled kernel launched next takes the remaining images.  forced = ZUTIS_DECODE_PRUNED (every finite image).
// *launched = false with ZUTIS_OK means "does not fit, use the tiled kernel alone" (only when not forced).
int launch_decode_pruned(DecodeParams& p, bool forced, bool champions_ready, int label_dtype, void* workspace, int sms,
                         cudaStream_t stream, bool* launched);

}  // namespace zutis
