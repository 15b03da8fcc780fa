// Stand-alone probe (not part of the product): issue/execute rate of tcgen05.mma kind::tf32 on sm_100a by tile width,
// operand source (A from TMEM vs shared memory) and accumulator reuse.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_rate_probe.bin tools/umma_rate_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int kind_bits) {   // kind_bits: 2 = tf32, 1 = bf16
    return (1u << 4) | ((uint32_t)kind_bits << 7) | ((uint32_t)kind_bits << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_ts_tf32(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// mode 0: TS tf32 (A in TMEM), 1: SS tf32, 2: SS bf16.  acc_mode 0: every MMA accumulates into the same D,
// 1: alternate between two accumulators.  batch: MMAs between commits (+ wait) -- models a stage.
template <int MODE, int BATCH>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int acc_mode, int reps, long* out) {
    constexpr int mode = MODE, batch = BATCH;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base;                 // 128 rows x 128 B
    const uint32_t b_smem = base + 16384;         // 256 rows x 128 B
    const uint32_t bar = base + 16384 + 32768;
    const uint32_t tslot = bar + 8;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(128, n, mode == 2 ? 1 : 2);
        const uint64_t adesc = make_desc_sw128(a_smem), bdesc = make_desc_sw128(b_smem);
        const uint32_t a_tmem = tmem + 448;       // 64 columns of A at the end
        uint32_t ph = 0;
        const long t0 = clock64();
        long issue = 0;
        for (int r = 0; r < reps; ++r) {
            const long i0 = clock64();
            if (elect_one()) {
#pragma unroll
            for (int m = 0; m < batch; ++m) {
                const uint32_t d = tmem + ((acc_mode && (m & 1)) ? (uint32_t)n : 0u);
                const uint64_t adv = (uint64_t)(((m & 3) * 32) >> 4);
                if (mode == 0) umma_ts_tf32(d, a_tmem + (m & 3) * 8, bdesc + adv, idesc, 1u);
                else if (mode == 1) umma_ss_tf32(d, adesc + adv, bdesc + adv, idesc, 1u);
                else umma_ss_bf16(d, adesc + adv, bdesc + adv, idesc, 1u);
            }
            umma_commit(bar);
            }
            __syncwarp();
            issue += clock64() - i0;
            mbar_wait(bar, ph);
            ph ^= 1;
        }
        const long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = issue; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE, int BATCH>
void run(int n, int acc_mode, long* out) {
    const char* names[] = {"TS tf32", "SS tf32", "SS bf16"};
    CK(cudaFuncSetAttribute(rate_kernel<MODE, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int reps = 1536 / BATCH;
    for (int w = 0; w < 2; ++w) rate_kernel<MODE, BATCH><<<148, 128, 64 * 1024>>>(n, acc_mode, reps, out);
    CK(cudaDeviceSynchronize());
    long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    const double per = (double)h[0] / (reps * BATCH);
    const double fma_per_clk = 128.0 * n * (MODE == 2 ? 16 : 8) / per;
    printf("%s N=%3d acc=%s batch=%3d : %6.1f cyc/MMA (issue part %6.1f)  %6.0f FMA/clk/SM\n", names[MODE], n,
           acc_mode ? "alt" : "one", BATCH, per, (double)h[1] / (reps * BATCH), fma_per_clk);
}

int main() {
    long* out; CK(cudaMalloc(&out, 16));
    for (int n : {64, 96, 128, 192, 256}) {
        run<0, 4>(n, 0, out); run<0, 12>(n, 0, out); run<0, 48>(n, 0, out);
        if (2 * n <= 448) run<0, 12>(n, 1, out);
        run<1, 4>(n, 0, out); run<1, 12>(n, 0, out); run<1, 48>(n, 0, out);
        run<2, 12>(n, 0, out); run<2, 48>(n, 0, out);
    }
    return 0;
}
