// Shared declarations of the decode kernels (decode_score.cu: generic + tiled + threshold kernels and the dispatcher;
// decode_cells.cu: the per-cell candidate-pruning kernel).
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
};

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

// decode_cells.cu: the per-cell pruning kernel.  counter: two zeroed words of workspace for dynamic work distribution
// (NULL: static).  *launched = false with ZUTIS_OK means "shape not taken, use another kernel" (only when not forced).
int launch_decode_cells(const DecodeParams& p, bool forced, int label_dtype, unsigned* counter, int sms, cudaStream_t stream,
                        bool* launched);

}  // namespace zutis
