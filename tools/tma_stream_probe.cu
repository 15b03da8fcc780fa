// Stand-alone probe (not part of the product): how fast can 148 persistent CTAs stream a [rows][512] fp32 tensor
// from HBM into shared memory with nothing consuming it, by request shape?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_stream_probe tools/tma_stream_probe.cu -lcuda
//   run  : /tmp/tma_stream_probe
// Modes: 0 = 2-D TMA box {32 floats, 128 rows}, 128B swizzle (what gemm_tcgen05.cu issues), k-slab inner loop
//        1 = same box, row-tile inner loop (each CTA walks down its rows for a fixed k slab)
//        2 = 2-D TMA box {64 floats, 64 rows}, no swizzle
//        3 = 2-D TMA box {256 floats, 16 rows}, no swizzle
//        4 = 1-D bulk copy of 16 KB contiguous
//        5 = coalesced LDG.128, 512 threads
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int K = 512;
constexpr int REQ_BYTES = 16384;

// rows per CTA chunk = 128; each chunk = 128 rows x 2 KB = 256 KB = 16 requests of 16 KB
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const float* src, long rows, int mode, int depth) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + (uint32_t)depth * REQ_BYTES;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long chunks = rows / 128;
    long n = 0;
    for (long c = blockIdx.x; c < chunks; c += gridDim.x) {
        for (int r = 0; r < 16; ++r, ++n) {
            const int slot = (int)(n % depth);
            const uint32_t bar = bars + 8 * slot;
            if (n >= depth) mbar_wait(bar, (uint32_t)((n / depth - 1) & 1));
            mbar_expect(bar, REQ_BYTES);
            const uint32_t dst = base + (uint32_t)slot * REQ_BYTES;
            if (mode == 0) tma_2d(dst, &map, bar, r * 32, (int)(c * 128));
            else if (mode == 2) tma_2d(dst, &map, bar, (r & 7) * 64, (int)(c * 128 + (r >> 3) * 64));
            else if (mode == 3) tma_2d(dst, &map, bar, (r & 1) * 256, (int)(c * 128 + (r >> 1) * 16));
            else if (mode == 4) bulk_1d(dst, src + (c * 128 * K) + (long)r * (REQ_BYTES / 4), REQ_BYTES, bar);
        }
    }
    // drain
    for (long m = (n > depth ? n - depth : 0); m < n; ++m) mbar_wait(bars + 8 * (int)(m % depth), (uint32_t)((m / depth) & 1));
}

// mode 1: k slab fixed per "pass", CTA walks over its row tiles: the same bytes in a different order
__global__ void __launch_bounds__(128, 1) stream_kernel_rowinner(const __grid_constant__ CUtensorMap map, long rows, int depth) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + (uint32_t)depth * REQ_BYTES;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long chunks = rows / 128;
    long n = 0;
    for (int r = 0; r < 16; ++r)
        for (long c = blockIdx.x; c < chunks; c += gridDim.x, ++n) {
            const int slot = (int)(n % depth);
            const uint32_t bar = bars + 8 * slot;
            if (n >= depth) mbar_wait(bar, (uint32_t)((n / depth - 1) & 1));
            mbar_expect(bar, REQ_BYTES);
            tma_2d(base + (uint32_t)slot * REQ_BYTES, &map, bar, r * 32, (int)(c * 128));
        }
    for (long m = (n > depth ? n - depth : 0); m < n; ++m) mbar_wait(bars + 8 * (int)(m % depth), (uint32_t)((m / depth) & 1));
}

__global__ void __launch_bounds__(512, 1) ldg_kernel(const float4* src, long n4, float* sink) {
    float acc = 0.f;
    // per CTA: contiguous 256 KB chunks like the TMA modes
    const long chunk4 = 128L * K / 4;
    const long chunks = n4 / chunk4;
    for (long c = blockIdx.x; c < chunks; c += gridDim.x) {
        const float4* p = src + c * chunk4;
#pragma unroll 8
        for (int i = threadIdx.x; i < chunk4; i += 512) {
            const float4 v = __ldcs(p + i);
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const long rows = 64L * 1600;                 // cfg2: 64 images x 40x40 pixels, 512 channels
    const size_t bytes = (size_t)rows * K * 4;
    float* buf[3];
    for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], 1, bytes)); }
    float* sink; CK(cudaMalloc(&sink, 4));
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(stream_kernel_rowinner, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));

    struct Shape { int mode; cuuint32_t bx, by; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2; const char* name; };
    const Shape shapes[] = {
        {0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tma {32,128} sw128 l2-256 k-inner"},
        {0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "tma {32,128} sw128 l2-128 k-inner"},
        {0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, "tma {32,128} sw128 l2-none k-inner"},
        {0, 32, 128, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tma {32,128} noswz l2-256 k-inner"},
        {1, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tma {32,128} sw128 l2-256 row-inner"},
        {2, 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tma {64,64} noswz"},
        {3, 256, 16, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tma {256,16} noswz"},
        {4, 0, 0, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "bulk 1-D 16 KB contiguous"},
        {5, 0, 0, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "LDG.128 coalesced 512 thr"},
    };
    for (const Shape& s : shapes) {
        for (int depth : {4, 8, 13}) {
            if (s.mode == 5 && depth != 4) continue;
            CUtensorMap maps[3];
            if (s.bx) {
                for (int i = 0; i < 3; ++i) {
                    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
                    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
                    cuuint32_t box[2] = {s.bx, s.by};
                    cuuint32_t estr[2] = {1, 1};
                    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf[i], dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, s.sw, s.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, s.name); return 1; }
                }
            } else {
                for (int i = 0; i < 3; ++i) maps[i] = CUtensorMap{};
            }
            const size_t smem = (size_t)depth * REQ_BYTES + 1024 + 256;
            const int iters = 30;
            float ms = 0;
            for (int it = -3; it < iters; ++it) {
                if (it == 0) CK(cudaEventRecord(e0));
                const int b = ((it % 3) + 3) % 3;
                if (s.mode == 5) ldg_kernel<<<148, 512>>>((const float4*)buf[b], (long)(bytes / 16), sink);
                else if (s.mode == 1) stream_kernel_rowinner<<<148, 128, smem>>>(maps[b], rows, depth);
                else stream_kernel<<<148, 128, smem>>>(maps[b], buf[b], rows, s.mode, depth);
            }
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1000.0 / iters;
            printf("%-40s depth %2d : %7.2f us  %7.1f GB/s\n", s.name, depth, us, bytes / us * 1e-3);
        }
    }
    return 0;
}
