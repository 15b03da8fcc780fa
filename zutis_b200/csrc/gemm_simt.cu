// fp32 FFMA contraction kernel (sm_100a SIMT path).
//
//   out[b][n][p] = act( sum_k A[b][n][k] * Bm[b][p][k] )
//
// This is the exact-fp32 implementation of the query x patch contraction
// (torch.einsum "nc,bchw->bnhw", networks/zutis.py:361-365; "bqc,bhwc->bqhw", :184-186).  It is
// what ZUTIS_GEMM_FP32_SIMT selects, and what shapes the tcgen05 kernel does not take (K not a
// multiple of 32, tiny problems) run on.  The tensor-core path lives in gemm_tcgen05.cu.
//
// Tiling: 64 pixels x 64 queries per CTA, K in slices of 16, 256 threads x (4 x 4) accumulators,
// operands staged k-major in shared memory so the inner product reads are conflict-free broadcasts.
// Accumulation is a single ascending-k fmaf chain per output: deterministic, fp32-grade.
#include "gemm.cuh"

namespace zutis {

constexpr int TP = 64, TN = 64, TK = 16, PAD = 4;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams g) {
    __shared__ float As[TK][TN + PAD];
    __shared__ float Bs[TK][TP + PAD];
    const int b = blockIdx.z;
    const long p0 = (long)blockIdx.x * TP;
    const int n0 = blockIdx.y * TN;
    const float* A = g.A + (long)b * g.strideA;
    const float* Bm = g.Bm + (long)b * g.strideB;
    const int tid = threadIdx.x;
    const int tp = (tid & 15) * 4;       // 4 consecutive pixels
    const int tn = (tid >> 4) * 4;       // 4 consecutive queries
    const int lrow = tid >> 2;           // 0..63: row this thread stages
    const int lk = (tid & 3) * 4;        // 4 consecutive k
    const bool vec_ok = ((g.K & 3) == 0) && ((g.lda & 3) == 0) && ((g.ldb & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(Bm) & 15) == 0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < g.K; k0 += TK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        const int n = n0 + lrow;
        const long pp = p0 + lrow;
        if (vec_ok && k0 + lk + 3 < g.K) {
            if (n < g.M) { const float4 v = *reinterpret_cast<const float4*>(A + (long)n * g.lda + k0 + lk); av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w; }
            if (pp < g.N) { const float4 v = *reinterpret_cast<const float4*>(Bm + pp * g.ldb + k0 + lk); bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w; }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + lk + i;
                if (k < g.K) {
                    if (n < g.M) av[i] = A[(long)n * g.lda + k];
                    if (pp < g.N) bv[i] = Bm[pp * g.ldb + k];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) { As[lk + i][lrow] = av[i]; Bs[lk + i][lrow] = bv[i]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tn]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tp]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a[i], bb[j], acc[i][j]);
        }
    }
    float* C = g.C + (long)b * g.strideC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + tn + i;
        if (n >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long pp = p0 + tp + j;
            if (pp >= g.N) continue;
            float v = acc[i][j];
            if (g.sigmoid) v = sigmoidf_exact(v);
            C[(long)n * g.stride_cn + pp * g.stride_cp] = v;
        }
    }
}

int launch_gemm_simt(const GemmParams& g, int batch, cudaStream_t stream) {
    dim3 grid((unsigned)((g.N + TP - 1) / TP), (unsigned)((g.M + TN - 1) / TN), (unsigned)batch);
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(g);
    return check_launch("gemm_simt_kernel");
}

}  // namespace zutis

using namespace zutis;

extern "C" size_t zutis_gemm_workspace_bytes(int M, long N, int K, int batch, int flags) {
    if ((flags & ZUTIS_GEMM_PRECISION_MASK) == ZUTIS_GEMM_FP32_SIMT) return 0;
    return gemm_tcgen05_workspace_bytes(M, N, K, batch, flags);
}

static int gemm_logits_impl(const float* A, long lda, long strideA,
                            const float* Bm, long ldb, long strideB,
                            float* C, long stride_cn, long stride_cp, long strideC,
                            int M, long N, int K, int batch, int flags,
                            void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ZUTIS_REQUIRE(A && Bm && C, "zutis_gemm_logits: NULL pointer");
    ZUTIS_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "zutis_gemm_logits: non-positive shape M=%d N=%ld K=%d batch=%d", M, N, K, batch);
    ZUTIS_REQUIRE(lda >= K && ldb >= K, "zutis_gemm_logits: row stride smaller than K");
    ZUTIS_REQUIRE(batch <= 65535, "zutis_gemm_logits: batch=%d too large", batch);
    const int prec = flags & ZUTIS_GEMM_PRECISION_MASK;
    ZUTIS_REQUIRE(prec != 3, "zutis_gemm_logits: bad precision flags %d", flags);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    GemmParams g;
    g.A = A; g.lda = lda; g.strideA = strideA; g.Bm = Bm; g.ldb = ldb; g.strideB = strideB;
    g.C = C; g.stride_cn = stride_cn; g.stride_cp = stride_cp; g.strideC = strideC;
    g.M = M; g.N = N; g.K = K; g.sigmoid = (flags & ZUTIS_GEMM_SIGMOID) ? 1 : 0;
    if (prec == ZUTIS_GEMM_FP32_SIMT) return launch_gemm_simt(g, batch, stream);
    if (!gemm_tcgen05_supports(g, batch, flags))
        return fail(ZUTIS_ERR_UNSUPPORTED,
                    "zutis_gemm_logits: tcgen05 path needs K %% 32 == 0, 16-byte aligned K-contiguous rows and M <= 1024 "
                    "(M=%d N=%ld K=%d lda=%ld ldb=%ld); use ZUTIS_GEMM_FP32_SIMT", M, N, K, lda, ldb);
    return launch_gemm_tcgen05(g, batch, flags, workspace, workspace_bytes, stream);
}

extern "C" int zutis_gemm_logits(const float* A, long lda, long strideA,
                                 const float* Bm, long ldb, long strideB,
                                 float* C, long stride_cn, long stride_cp, long strideC,
                                 int M, long N, int K, int batch, int flags,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    return gemm_logits_impl(A, lda, strideA, Bm, ldb, strideB, C, stride_cn, stride_cp, strideC, M, N, K, batch, flags,
                            workspace, workspace_bytes, stream);
}
