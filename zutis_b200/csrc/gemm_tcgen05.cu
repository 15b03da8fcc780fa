// tcgen05 / TMA contraction kernel for sm_100a: fp32-grade logits on the 5th-gen tensor cores.
//
//   out[b][n][p] = act( sum_k A[b][n][k] * Bm[b][p][k] )         (n: categories/queries, p: pixels)
//
// replaces torch.einsum("nc,bchw->bnhw") (networks/zutis.py:361-365) and the mask-proposal einsums
// (:184-186, :196-198, + sigmoid :209).  Pixels ride the UMMA M dimension (128 rows per tile), the
// categories the N dimension (padded to 16, <= 256 per tile), K is consumed in slabs of 32 fp32
// = one 128-byte swizzled shared-memory row.
//
// Precision.  The label bar (>= 99.99 % agreement) needs fp32-grade logits; kind::tf32 keeps 11
// significand bits.  Each operand is therefore split  x = hi + lo,  hi = tf32_rn(x), lo = tf32_rn(x - hi),
// and  D += hi*hi + hi*lo + lo*hi  goes into ONE fp32 TMEM accumulator (dropped lo*lo term ~ 2^-22).
//   * The category operand (81..920 rows) is split once per call by split_operand_kernel into a
//     zero-padded workspace [hi | lo] and arrives by TMA.
//   * The pixel operand is split ON ITS WAY INTO TENSOR MEMORY: TMA lands the raw fp32 tile in shared
//     memory, eight converter warps read it once (un-swizzling their own row), and tcgen05.st the `hi`
//     and `lo` halves into per-stage TMEM columns; the MMAs then take A from TMEM (tcgen05.mma [d],[a],b)
//     and only the category operand B is ever re-read from shared memory.  (r01c measured the previous
//     all-in-smem variant as shared-memory-port bound: 172 KB moved per K slab; this moves 92 KB.)
//   ZUTIS_GEMM_TF32 runs the single pass hi*hi only: the reduced-precision mode that meets the 2e-2 logit bar.
//
// Warp roles (512 threads, 1 CTA / SM, persistent over tiles) and the three decoupled rings they run:
//   warp 0      pixel-tile TMA producer   A ring  (smem, SA slots of 16 KB):  waits a_free[i]  -> arms a_full[i]
//   warp 3      category TMA producer     B ring  (smem, SB slots hi|lo):    waits b_free[j]  -> arms b_full[j]
//   warps 8-15  converters                A ring -> T ring: wait a_full[i]; ld.shared; split; wait t_free[k];
//                                         tcgen05.st; arrive t_ready[k], then a_free[i]
//   warp 1      MMA issuer (one lane)     waits t_ready[k], b_full[j], tmem_empty[a]; 3 MMAs per k-step (2 when the
//                                         hi and lo category halves are issued as one double-width MMA, see `stacked`);
//                                         commits t_free[k], b_free[j], and tmem_full[a] after the last slab
//   warp 2      TMEM allocator / deallocator
//   warps 4-7   epilogue: tcgen05.ld -> (sigmoid) -> global stores, any (stride_cn, stride_cp)
// Decoupling matters: with one ring the HBM stream of pixel tiles had to wait for MMA completion + commit
// before every refill and the kernel was latency-bound at ~1400 cycles per slab whatever the MMA count.
// TMEM map (512 columns): [0, acc_bufs*acc_cols) accumulators (double-buffered when they fit, so the epilogue
// of tile i overlaps the MMAs of tile i+1), then ST slots of 64 columns (A_hi | A_lo).
#include "gemm.cuh"

#include <cuda.h>

#include <mutex>

namespace zutis {

namespace {

constexpr int BLOCK_M = 128;          // pixels per tile (UMMA M)
constexpr int BLOCK_K = 32;           // fp32 per slab = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;             // tf32 MMA K
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;     // 16 KB
constexpr int NUM_THREADS = 512;
constexpr int MAX_A = 8, MAX_B = 5, MAX_T = 5;   // ring depth limits
constexpr int NUM_CONVERTER_WARPS = 8;
constexpr int SMEM_LIMIT = 232448;    // 227 KB opt-in maximum per CTA

struct TcParams {
    float* C;
    long stride_cn, stride_cp, strideC;
    int M;            // valid categories
    long N;           // pixels per image
    int K;
    int batch;
    int umma_n;       // categories per N tile (multiple of 16, <= 256)
    int acc_cols;     // TMEM columns of one accumulator: umma_n, or 2 * umma_n when hi*hi and hi*lo are issued as one MMA (stacked)
    int stacked;
    int n_tiles;      // N tiles
    int p_tiles;      // pixel tiles per image
    int a_rows_per_image;   // rows of the split workspace per image (0 => shared by the batch)
    int sa, sb, st;   // ring depths: smem pixel tiles, smem category tiles, TMEM split-pixel slots
    int b_slot_bytes;
    int tmem_cols;
    int acc_bufs;     // accumulator buffers in TMEM (2 when they fit next to the A stages)
    int a_col0;       // first TMEM column of the per-stage A operand (64 columns per stage: hi | lo)
    int passes;       // 3 = hi*hi + hi*lo + lo*hi, 1 = hi*hi only
    int sigmoid;
    // raw fp32 operands, for the exact re-computation of pixels whose products are not finite (epilogue)
    const float* A; long lda, strideA;
    const float* Bm; long ldb, strideB;
};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// one lane of a CONVERGED warp; tcgen05.mma / commit / TMA issued under this predicate compile to a single
// UTCHMMA / UTCBAR / UTMALDG (inside `if (lane == 0)` the compiler wraps each in an elect-and-retry loop, which
// made the one-thread MMA issuer the kernel's critical path)
__device__ __forceinline__ bool elect_one() {
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, single CTA: the A operand lives in tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t to_tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
// Same rounding (nearest, ties away) for the converter warps' inner loop, spelled so that it costs one FMA-pipe
// integer add and one LOP3 instead of cvt.rna's four ALU-pipe instructions (the ALU pipe is half rate on sm_100
// and was the contraction kernel's limiter).  +inf stays +inf; NaN stays NaN unless its payload is all ones.
__device__ __forceinline__ uint32_t to_tf32_rn_fast(float x) {
    uint32_t u;
    asm("mad.lo.u32 %0, %1, 1, 0x1000;" : "=r"(u) : "r"(__float_as_uint(x)));
    return u & 0xffffe000u;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major; 1) | [32,46) SBO >> 4 = 1024 B between
//   8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------- operand split
// dst_hi/dst_lo: [rows_total][K]; rows beyond M (padding up to rows_per_image) are zero.
__global__ void __launch_bounds__(256) split_operand_kernel(const float* A, long lda, long strideA, int M, int K,
                                                            int rows_per_image, int images, uint32_t* dst_hi, uint32_t* dst_lo) {
    const long total = (long)images * rows_per_image * K;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const long r = i / K;
        const int row = (int)(r % rows_per_image);
        const long img = r / rows_per_image;
        uint32_t hi = 0, lo = 0;
        if (row < M) {
            const float x = A[img * strideA + (long)row * lda + k];
            hi = to_tf32_rn(x);
            const float rest = (fabsf(x) <= 3.402823466e38f) ? __fsub_rn(x, __uint_as_float(hi)) : 0.0f;
            lo = to_tf32_rn(rest);
        }
        dst_hi[i] = hi;
        dst_lo[i] = lo;
    }
}

// ----------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_pix, const __grid_constant__ CUtensorMap map_cat_hi,
                    const __grid_constant__ CUtensorMap map_cat_lo, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B wants 1024-B alignment
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b_tile_bytes = p.umma_n * BLOCK_K * 4;

    // shared-memory map: A ring | B ring | barriers | tmem pointer
    auto slot_a = [&](int i) { return base + (uint32_t)i * A_TILE_BYTES; };
    const uint32_t b_base = base + (uint32_t)p.sa * A_TILE_BYTES;
    auto slot_b_hi = [&](int j) { return b_base + (uint32_t)j * p.b_slot_bytes; };
    auto slot_b_lo = [&](int j) { return b_base + (uint32_t)j * p.b_slot_bytes + b_tile_bytes; };
    const uint32_t bar_base = b_base + (uint32_t)p.sb * p.b_slot_bytes;
    auto bar_a_full = [&](int i) { return bar_base + 8u * i; };
    auto bar_a_free = [&](int i) { return bar_base + 8u * (MAX_A + i); };
    auto bar_b_full = [&](int j) { return bar_base + 8u * (2 * MAX_A + j); };
    auto bar_b_free = [&](int j) { return bar_base + 8u * (2 * MAX_A + MAX_B + j); };
    auto bar_t_ready = [&](int k) { return bar_base + 8u * (2 * MAX_A + 2 * MAX_B + k); };
    auto bar_t_free = [&](int k) { return bar_base + 8u * (2 * MAX_A + 2 * MAX_B + MAX_T + k); };
    auto bar_tmem_full = [&](int a) { return bar_base + 8u * (2 * MAX_A + 2 * MAX_B + 2 * MAX_T + a); };
    auto bar_tmem_empty = [&](int a) { return bar_base + 8u * (2 * MAX_A + 2 * MAX_B + 2 * MAX_T + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_A + 2 * MAX_B + 2 * MAX_T + 4);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_pix); prefetch_tmap(&map_cat_hi); prefetch_tmap(&map_cat_lo);
        for (int i = 0; i < p.sa; ++i) { mbar_init(bar_a_full(i), 1); mbar_init(bar_a_free(i), NUM_CONVERTER_WARPS); }
        for (int j = 0; j < p.sb; ++j) { mbar_init(bar_b_full(j), 1); mbar_init(bar_b_free(j), 1); }
        for (int k = 0; k < p.st; ++k) { mbar_init(bar_t_ready(k), NUM_CONVERTER_WARPS); mbar_init(bar_t_free(k), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tmem_full(a), 1); mbar_init(bar_tmem_empty(a), 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int num_k = p.K / BLOCK_K;
    const long total_tiles = (long)p.batch * p.p_tiles * p.n_tiles;

    if (warp == 0) {
        // ========================= pixel-tile TMA producer (A ring) =========================
        {
            int i = 0;
            uint32_t ph = 0;
            for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const long r = t / p.n_tiles;
                const int pix_row = (int)((long)(r / p.p_tiles) * p.N + (long)(r % p.p_tiles) * BLOCK_M);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(bar_a_free(i), ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_a_full(i), (uint32_t)A_TILE_BYTES);
                        tma_load_2d(slot_a(i), &map_pix, bar_a_full(i), kb * BLOCK_K, pix_row);
                    }
                    __syncwarp();
                    if (++i == p.sa) { i = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ========================= category TMA producer (B ring) ==========================
        {
            int j = 0;
            uint32_t ph = 0;
            const uint32_t tx = (p.passes == 3 ? 2u : 1u) * (uint32_t)b_tile_bytes;
            for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = (int)(t % p.n_tiles);
                const int b = (int)((t / p.n_tiles) / p.p_tiles);
                const int cat_row = b * p.a_rows_per_image + nt * p.umma_n;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(bar_b_free(j), ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_b_full(j), tx);
                        tma_load_2d(slot_b_hi(j), &map_cat_hi, bar_b_full(j), kb * BLOCK_K, cat_row);
                        if (p.passes == 3) tma_load_2d(slot_b_lo(j), &map_cat_lo, bar_b_full(j), kb * BLOCK_K, cat_row);
                    }
                    __syncwarp();
                    if (++j == p.sb) { j = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        {
            const uint32_t idesc = make_idesc_tf32(BLOCK_M, p.umma_n);
            uint32_t acc_it = 0;
            int j = 0, k = 0;
            uint32_t phj = 0, phk = 0;
            for (long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++acc_it) {
                const int a = acc_it % p.acc_bufs;
                const uint32_t aph = (acc_it / p.acc_bufs) & 1;
                mbar_wait(bar_tmem_empty(a), aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * p.acc_cols);
                const uint32_t idesc2 = make_idesc_tf32(BLOCK_M, 2 * p.umma_n);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(bar_b_full(j), phj);
                    mbar_wait(bar_t_ready(k), phk);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + (uint32_t)(p.a_col0 + k * 64);
                    const uint32_t a_lo = a_hi + 32;
                    const uint64_t b_hi = make_desc_sw128(slot_b_hi(j));
                    const uint64_t b_lo = make_desc_sw128(slot_b_lo(j));
                    if (elect_one()) {
                        if (p.stacked) {
                            // The category tile's hi and lo halves lie back to back in shared memory (whole 8-row groups), so
                            // A_hi x [B_hi ; B_lo] is ONE MMA of twice the width: hi*hi lands in columns [0, n), hi*lo in
                            // [n, 2n) of the accumulator, and the epilogue adds the halves.  Same pipe time and the same
                            // launch time (8 instead of 12 MMAs to issue per slab), but the small hi*lo terms are summed
                            // apart from the large ones: the measured logits error drops from 4.6e-6 to 3.1e-6 of max|logit|.
#pragma unroll
                            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                                const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);
                                umma_tf32_ts(tmem_d, a_hi + kk * UMMA_K, b_hi + adv, idesc2, (kb | kk) != 0 ? 1u : 0u);
                                umma_tf32_ts(tmem_d, a_lo + kk * UMMA_K, b_hi + adv, idesc, 1u);
                            }
                        } else {
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                            const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4); // +32 bytes per k-step inside the swizzle row
                            umma_tf32_ts(tmem_d, a_hi + kk * UMMA_K, b_hi + adv, idesc, (kb | kk) != 0 ? 1u : 0u);
                            if (p.passes == 3) {
                                umma_tf32_ts(tmem_d, a_hi + kk * UMMA_K, b_lo + adv, idesc, 1u);
                                umma_tf32_ts(tmem_d, a_lo + kk * UMMA_K, b_hi + adv, idesc, 1u);
                            }
                        }
                        }
                        umma_commit(bar_t_free(k));     // TMEM slot k and smem slot j may be refilled once these MMAs retire
                        umma_commit(bar_b_free(j));
                    }
                    __syncwarp();
                    if (++j == p.sb) { j = 0; phj ^= 1; }
                    if (++k == p.st) { k = 0; phk ^= 1; }
                }
                if (elect_one()) umma_commit(bar_tmem_full(a));      // accumulator a is complete
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================ epilogue ================================
        const int quad = warp & 3;                      // TMEM lanes [32*quad, 32*quad+32)
        uint32_t acc_it = 0;
        for (long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++acc_it) {
            const int nt = (int)(t % p.n_tiles);
            const long r = t / p.n_tiles;
            const int pt = (int)(r % p.p_tiles);
            const int b = (int)(r / p.p_tiles);
            const int a = acc_it % p.acc_bufs;
            const uint32_t aph = (acc_it / p.acc_bufs) & 1;
            mbar_wait(bar_tmem_full(a), aph);
            tc_fence_after();
            const long pix = (long)pt * BLOCK_M + quad * 32 + lane;
            const bool row_ok = pix < p.N;
            float* crow = p.C + (long)b * p.strideC + pix * p.stride_cp;
            const bool vec_ok = (p.stride_cn == 1) && ((p.stride_cp & 3) == 0) && ((p.strideC & 3) == 0) &&
                                ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
            float row_sum = 0.f;
            for (int c = 0; c < p.umma_n / 16; ++c) {
                uint32_t v[16];
                tmem_ld_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.acc_cols + c * 16), v);
                if (p.stacked) {                         // + the hi*lo half of the accumulator
                    uint32_t v2[16];
                    tmem_ld_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.acc_cols + p.umma_n + c * 16), v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__fadd_rn(__uint_as_float(v[e]), __uint_as_float(v2[e])));
                }
                tmem_ld_wait();
                const int n0 = nt * p.umma_n + c * 16;
                if (row_ok && n0 < p.M) {
                    float f[16];
                    const bool whole16 = n0 + 16 <= p.M;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        f[j] = __uint_as_float(v[j]);
                        if (whole16 || n0 + j < p.M) row_sum = __fadd_rn(row_sum, f[j]);      // NaN / inf anywhere in the row surfaces here
                        if (p.sigmoid) f[j] = sigmoidf_exact(f[j]);
                    }
                    if (vec_ok) {
                        // pixel-major rows: padding columns up to the row pitch hold zeros (zero-padded operand rows)
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            if (n0 + j + 3 < p.stride_cp && n0 + j < p.M)
                                *reinterpret_cast<float4*>(crow + n0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                            else
                                for (int e = 0; e < 4; ++e)
                                    if (n0 + j + e < p.M) crow[n0 + j + e] = f[j + e];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < p.M) crow[(long)(n0 + j) * p.stride_cn] = f[j];
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tmem_empty(a));
            // A non-finite token (or an overflow) cannot go through the hi/lo split: inf - inf = NaN, and inf * 0 in a
            // correction pass is NaN where fp32 arithmetic keeps the infinity.  Such a pixel -- a whole row of the tile,
            // since every category multiplies the same token -- is recomputed here from the raw operands as one
            // ascending-k fmaf chain, which has torch.einsum's NaN / +-inf pattern (zutis.py:361-365).  Rare and slow.
            if (row_ok && !(fabsf(row_sum) <= 3.402823466e38f)) {
                const float* trow = p.Bm + (long)b * p.strideB + pix * p.ldb;
                const int n_end = min(p.M, (nt + 1) * p.umma_n);
                for (int n = nt * p.umma_n; n < n_end; ++n) {
                    const float* arow = p.A + (long)b * p.strideA + (long)n * p.lda;
                    float acc = 0.f;
                    for (int k = 0; k < p.K; ++k) acc = __fmaf_rn(__ldg(arow + k), __ldg(trow + k), acc);
                    crow[(long)n * p.stride_cn] = p.sigmoid ? sigmoidf_exact(acc) : acc;
                }
            }
            __syncwarp();
        }
    } else if (warp >= 8) {
        // =============================== converters ===============================
        // Thread = one pixel row of the tile (TMEM lane) and one half of the 32-wide K slab.  Row m of the
        // TMA-written tile lives at m*128 bytes with its 16-byte chunks XOR-swizzled by (m & 7).
        const int quad = warp & 3;                      // TMEM lanes [32*quad, 32*quad+32)
        const int half = (warp - 8) >> 2;               // K columns [16*half, 16*half+16)
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        int i = 0, k = 0;
        uint32_t phi = 0, phk = 0;
        for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(bar_a_full(i), phi);
                const uint32_t row_base = slot_a(i) + (uint32_t)row * 128u;
                float x[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t chunk = (uint32_t)((half * 4 + c) ^ (row & 7));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(x[c * 4]), "=f"(x[c * 4 + 1]), "=f"(x[c * 4 + 2]), "=f"(x[c * 4 + 3]) : "r"(row_base + chunk * 16u));
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    // x = hi + lo exactly (x - hi is exact in fp32); lo is then rounded to tf32 as well.
                    // A +-inf token gives lo = NaN, i.e. NaN logits where the reference has +-inf.
                    hi[e] = to_tf32_rn_fast(x[e]);
                    lo[e] = to_tf32_rn_fast(__fsub_rn(x[e], __uint_as_float(hi[e])));
                }
                const int i_prev = i;
                if (++i == p.sa) { i = 0; phi ^= 1; }
                mbar_wait(bar_t_free(k), phk ^ 1);
                tc_fence_after();
                const uint32_t col = (uint32_t)(p.a_col0 + k * 64 + half * 16);
                tmem_st_x16(tmem_base + lane_addr + col, hi);
                if (p.passes == 3) tmem_st_x16(tmem_base + lane_addr + col + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                // ONE arrival per warp on each barrier: with one per thread the 512 shared-memory barrier updates per
                // slab (~2 cycles each) were the kernel's real limiter, ~1100 cycles per slab whatever else changed.
                // The smem slot is released only here, after every lane's tcgen05.st of the values it loaded from the
                // slot has completed.  Arriving right after the ld.shared (no register dependence on the loads) let
                // the barrier unit overtake the load unit; the next TMA then overwrote rows that were still being
                // read (seen as wrong pixel rows, but only on a CTA's 2nd+ tile, when the rings run full).
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar_t_ready(k));
                    mbar_arrive(bar_a_free(i_prev));
                }
                if (++k == p.st) { k = 0; phk ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// --------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        (void)cudaGetLastError();
    });
    return fn;
}

// 2-D fp32 tensor [rows][K] with row pitch `ld` floats, box = {32 floats, box_rows}, 128B swizzle, zero OOB fill
int make_map(CUtensorMap* map, const void* ptr, long rows, int K, long ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(ZUTIS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZUTIS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ZUTIS_OK;
}

struct Plan {
    int n_tiles, umma_n, rows_per_image, sa, sb, st, b_slot_bytes, tmem_cols, acc_bufs, a_col0, acc_cols, stacked;
    size_t smem;
};

Plan make_plan(int M, bool three_pass = false) {
    Plan pl;
    pl.n_tiles = (M + 255) / 256;
    const int per = (M + pl.n_tiles - 1) / pl.n_tiles;
    pl.umma_n = (per + 15) & ~15;
    pl.rows_per_image = pl.n_tiles * pl.umma_n;
    pl.b_slot_bytes = 2 * pl.umma_n * BLOCK_K * 4;
    const int budget = SMEM_LIMIT - 1024 - 512;                       // alignment slack, barriers
    pl.sb = (budget - 3 * A_TILE_BYTES) / pl.b_slot_bytes;
    if (pl.sb > MAX_B) pl.sb = MAX_B;
    pl.sa = pl.sb >= 1 ? (budget - pl.sb * pl.b_slot_bytes) / A_TILE_BYTES : 0;
    if (pl.sa > MAX_A) pl.sa = MAX_A;
    pl.tmem_cols = 512;
    // TMEM: accumulators first, then 64 columns (A_hi | A_lo) per slot.  Double-buffer the accumulator when
    // at least 3 slots still fit beside it.
    // 3xTF32 with <= 96 categories per tile: hi*hi and hi*lo as one double-width MMA into a double-width accumulator;
    // two such accumulators (384 columns) leave two 64-column slots for the split pixel operand
    pl.stacked = (three_pass && pl.n_tiles == 1 && 4 * pl.umma_n + 2 * 64 <= 512) ? 1 : 0;
    pl.acc_cols = pl.stacked ? 2 * pl.umma_n : pl.umma_n;
    pl.acc_bufs = pl.stacked ? 2 : ((((2 * pl.umma_n + 31) & ~31) + 3 * 64 <= 512) ? 2 : 1);
    pl.a_col0 = (pl.acc_bufs * pl.acc_cols + 31) & ~31;
    pl.st = (512 - pl.a_col0) / 64;
    if (pl.st > MAX_T) pl.st = MAX_T;
    pl.smem = (size_t)pl.sa * A_TILE_BYTES + (size_t)pl.sb * pl.b_slot_bytes + 1024 + 512;
    return pl;
}

}  // namespace

size_t gemm_tcgen05_workspace_bytes(int M, long, int K, int batch, int) {
    // worst case: per-image category operand (queries); [hi | lo], zero padded
    const Plan pl = make_plan(M);
    return (size_t)2 * batch * pl.rows_per_image * K * 4;
}

bool gemm_tcgen05_supports(const GemmParams& g, int batch, int flags) {
    if ((flags & ZUTIS_GEMM_PRECISION_MASK) == ZUTIS_GEMM_FP32_SIMT) return false;
    if (g.K % BLOCK_K != 0 || g.M > 1024) return false;
    if ((g.ldb & 3) != 0 || (reinterpret_cast<uintptr_t>(g.Bm) & 15) != 0) return false;
    if (batch > 1 && g.strideB != g.N * g.ldb) return false;        // images must be consecutive rows of one 2-D tensor
    if ((long)batch * g.N >= 2147483647L) return false;
    { const Plan pl = make_plan(g.M, (flags & ZUTIS_GEMM_PRECISION_MASK) == ZUTIS_GEMM_TF32X3); if (pl.sa < 2 || pl.sb < 2 || pl.st < 2) return false; }
    return true;
}

int launch_gemm_tcgen05(const GemmParams& g, int batch, int flags, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const Plan pl = make_plan(g.M, (flags & ZUTIS_GEMM_PRECISION_MASK) == ZUTIS_GEMM_TF32X3);
    const bool shared_a = (g.strideA == 0);
    const int images = shared_a ? 1 : batch;
    const size_t half = (size_t)images * pl.rows_per_image * g.K * 4;
    if (!workspace || workspace_bytes < 2 * half)
        return fail(ZUTIS_ERR_WORKSPACE, "zutis_gemm_logits: workspace of %zu bytes needed, %zu given", 2 * half, workspace ? workspace_bytes : (size_t)0);
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
        return fail(ZUTIS_ERR_BAD_ARG, "zutis_gemm_logits: workspace must be 16-byte aligned");
    uint32_t* ws_hi = reinterpret_cast<uint32_t*>(workspace);
    uint32_t* ws_lo = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + half);
    const int sms = sm_count();

    const long split_elems = (long)images * pl.rows_per_image * g.K;
    long sblocks = (split_elems + 255) / 256;
    if (sblocks > (long)sms * 8) sblocks = (long)sms * 8;
    int st = ZUTIS_OK;
    if (!(flags & ZUTIS_GEMM_A_PREPARED)) {
        split_operand_kernel<<<(unsigned)sblocks, 256, 0, stream>>>(g.A, g.lda, g.strideA, g.M, g.K, pl.rows_per_image, images, ws_hi, ws_lo);
        st = check_launch("split_operand_kernel");
        if (st != ZUTIS_OK) return st;
    }

    CUtensorMap map_pix, map_hi, map_lo;
    st = make_map(&map_pix, g.Bm, (long)batch * g.N, g.K, g.ldb, BLOCK_M);
    if (st != ZUTIS_OK) return st;
    st = make_map(&map_hi, ws_hi, (long)images * pl.rows_per_image, g.K, g.K, pl.umma_n);
    if (st != ZUTIS_OK) return st;
    st = make_map(&map_lo, ws_lo, (long)images * pl.rows_per_image, g.K, g.K, pl.umma_n);
    if (st != ZUTIS_OK) return st;

    TcParams p;
    p.C = g.C; p.stride_cn = g.stride_cn; p.stride_cp = g.stride_cp; p.strideC = g.strideC;
    p.M = g.M; p.N = g.N; p.K = g.K; p.batch = batch;
    p.acc_cols = pl.acc_cols; p.stacked = pl.stacked;
    p.umma_n = pl.umma_n; p.n_tiles = pl.n_tiles; p.p_tiles = (int)((g.N + BLOCK_M - 1) / BLOCK_M);
    p.a_rows_per_image = shared_a ? 0 : pl.rows_per_image;
    p.sa = pl.sa; p.sb = pl.sb; p.st = pl.st; p.b_slot_bytes = pl.b_slot_bytes; p.tmem_cols = pl.tmem_cols; p.acc_bufs = pl.acc_bufs; p.a_col0 = pl.a_col0;
    p.passes = ((flags & ZUTIS_GEMM_PRECISION_MASK) == ZUTIS_GEMM_TF32X3) ? 3 : 1;
    p.sigmoid = g.sigmoid;
    p.A = g.A; p.lda = g.lda; p.strideA = g.strideA; p.Bm = g.Bm; p.ldb = g.ldb; p.strideB = g.strideB;

    const size_t smem = pl.smem;
    ZUTIS_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long total_tiles = (long)batch * p.p_tiles * p.n_tiles;
    const unsigned grid = (unsigned)(total_tiles < sms ? total_tiles : sms);
    gemm_tcgen05_kernel<<<grid, NUM_THREADS, smem, stream>>>(map_pix, map_hi, map_lo, p);
    return check_launch("gemm_tcgen05_kernel");
}

}  // namespace zutis
