// Shared host/device helpers for libzutis_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/zutis_b200.h"

namespace zutis {

// ---------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);
int sm_count();            // SMs of the current device (cached per device)
int current_device_ok();   // ZUTIS_OK iff current device is sm_100

// decode_score.cu: tiled interp(probabilities) > threshold kernel; ZUTIS_ERR_UNSUPPORTED => use the generic kernel
int launch_threshold_tiled(const float* probs, long sb, long sq, long sy, long sx, int B, int Q, int h, int w, int H, int W,
                           float threshold, uint32_t* bits, int* areas, cudaStream_t stream);

#define ZUTIS_REQUIRE(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) return ::zutis::fail(ZUTIS_ERR_BAD_ARG, __VA_ARGS__); \
    } while (0)

#define ZUTIS_CUDA(expr)                                           \
    do {                                                           \
        int _st = ::zutis::check_cuda((expr), #expr);              \
        if (_st != ZUTIS_OK) return _st;                           \
    } while (0)

// ------------------------------------------------------------ bilinear source coordinates
// ATen upsample_bilinear2d, align_corners=False, size= given (networks/zutis.py:367):
//   src = scale*(dst+0.5)-0.5 as ONE fma, clamped at 0; i0 = min(floor(src), in-1);
//   i1 = i0 + (i0 < in-1); l1 = clamp(src - i0, 0, 1); l0 = 1 - l1.   in == out: identity taps.
// `scale` = (float)in / (float)out is computed once on the host and passed in, so host and
// device tables are bit-identical to oracle/zutis_oracle.c:zo_axis_table.
struct AxisTap {
    int i0, i1;
    float l0, l1;
};

__host__ __device__ __forceinline__ AxisTap axis_tap(int d, int in, int out, float scale) {
    AxisTap t;
    if (in == out) {
        t.i0 = d; t.i1 = d; t.l0 = 1.0f; t.l1 = 0.0f;
        return t;
    }
#ifdef __CUDA_ARCH__
    float src = __fmaf_rn(scale, __fadd_rn((float)d, 0.5f), -0.5f);
#else
    float src = fmaf(scale, (float)d + 0.5f, -0.5f);
#endif
    src = src < 0.0f ? 0.0f : src;
    int a = (int)floorf(src);
    a = a > in - 1 ? in - 1 : a;
    t.i0 = a;
    t.i1 = a + (a < in - 1 ? 1 : 0);
#ifdef __CUDA_ARCH__
    float f = __fsub_rn(src, (float)a);
#else
    float f = src - (float)a;
#endif
    f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    t.l1 = f;
#ifdef __CUDA_ARCH__
    t.l0 = __fsub_rn(1.0f, f);
#else
    t.l0 = 1.0f - f;
#endif
    return t;
}

inline float axis_scale(int in, int out) { return (float)in / (float)out; }

#ifdef __CUDACC__
// out = fma(ly0, fma(lx0,a, lx1*b), ly1 * fma(lx0,c, lx1*d)): width first, then height, in the
// contraction pattern that is bit-exact with ATen's CPU kernel.
__device__ __forceinline__ float lerp_w(float l0, float a, float l1, float b) {
    return __fmaf_rn(l0, a, __fmul_rn(l1, b));
}

// torch.argmax ordering: strictly greater wins; NaN beats everything, first NaN kept.
__device__ __forceinline__ bool better_nan_aware(float v, float best) {
    return (v > best) || (v != v && best == best);
}

__device__ __forceinline__ long long load_label(const void* p, int dtype, size_t i) {
    switch (dtype) {
        case ZUTIS_GT_U8: return (long long)((const uint8_t*)p)[i];
        case ZUTIS_GT_I16: return (long long)((const int16_t*)p)[i];
        case ZUTIS_GT_I32: return (long long)((const int32_t*)p)[i];
        default: return ((const long long*)p)[i];
    }
}

// Add one count per lane to hist[key] (key < 0: lane has nothing to add), aggregating equal keys
// of a warp into a single atomic.  All 32 lanes must call.
__device__ __forceinline__ void warp_hist_add(int* hist, int key) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(full, key);
    if (key >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(hist + key, __popc(peers));
}
#endif

inline int gt_dtype_bytes(int dtype) {
    switch (dtype) {
        case ZUTIS_GT_U8: return 1;
        case ZUTIS_GT_I16: return 2;
        case ZUTIS_GT_I32: return 4;
        case ZUTIS_GT_I64: return 8;
        default: return 0;
    }
}

}  // namespace zutis
