// Helpers shared by the cell decode kernels (decode_cells.cu: the fused prune + evaluate kernel; decode_split.cu: the
// thread-per-cell pruning kernel and the list-driven evaluation kernel): shared-memory PTX accessors, packed fp32x2
// arithmetic, division by launch-invariant divisors, label pairs.
#pragma once
#include "decode.cuh"
#include "tma.cuh"

namespace zutis {

constexpr unsigned kFull = 0xffffffffu;

// Make a value opaque to the optimiser: it stays in its register instead of being re-derived from kernel parameters
// and special registers at every use (the compiler otherwise rematerialises shared-memory base addresses all over).
#define ZUTIS_KEEP(x) asm volatile("" : "+r"(x))

__device__ __forceinline__ bool elect_one_lane() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// 0 <= g < n as one unsigned compare (running_score.py:12)
template <typename GT>
__device__ __forceinline__ bool label_in_range(GT g, int n) { return (unsigned long long)(long long)g < (unsigned long long)n; }

__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f4_lo(const float4& v) { return pack2(v.x, v.y); }
__device__ __forceinline__ unsigned long long f4_hi(const float4& v) { return pack2(v.z, v.w); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u16(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// predicated store: no branch, whatever the compiler thinks of the condition
__device__ __forceinline__ void sts_u16_if(bool cond, uint32_t a, unsigned v) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %2, 0;\n"
        "@p st.shared.u16 [%0], %1;\n"
        "}\n"
        ::"r"(a), "h"((unsigned short)v), "r"((unsigned)cond) : "memory");
}

__device__ __forceinline__ void red_shared_add(uint32_t a, int v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename GT> struct Pair;
template <> struct Pair<uint8_t> { typedef uchar2 type; };
template <> struct Pair<int16_t> { typedef short2 type; };
template <> struct Pair<int32_t> { typedef int2 type; };
template <> struct Pair<long long> { typedef longlong2 type; };

// n / d for n < 2^31 with a divisor fixed per launch: q = umulhi(n, mul) >> shr (mul == 0: d == 1)
struct FastDiv {
    unsigned mul, shr;
};
__device__ __forceinline__ unsigned fast_div(unsigned n, FastDiv f) { return f.mul ? (__umulhi(n, f.mul) >> f.shr) : n; }
inline FastDiv make_fast_div(unsigned d) {
    FastDiv f;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    unsigned l = 0;
    while ((1u << l) < d) ++l;
    const unsigned p = 31 + l;
    f.mul = (unsigned)((((unsigned long long)1 << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}


}  // namespace zutis
