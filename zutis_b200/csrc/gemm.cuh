// Shared declarations of the two contraction kernels (gemm_simt.cu, gemm_tcgen05.cu).
#pragma once
#include "common.cuh"

namespace zutis {

// out[b][n][p] = act( sum_k A[b][n][k] * Bm[b][p][k] ), element (b,n,p) at C[b*strideC + n*stride_cn + p*stride_cp]
struct GemmParams {
    const float* A; long lda, strideA;
    const float* Bm; long ldb, strideB;
    float* C; long stride_cn, stride_cp, strideC;
    int M; long N; int K; int sigmoid;
};

#ifdef __CUDACC__
// torch.sigmoid in fp32 (networks/zutis.py:209)
__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }
#endif

int launch_gemm_simt(const GemmParams& g, int batch, cudaStream_t stream);
// implemented in gemm_tcgen05.cu
int launch_gemm_tcgen05(const GemmParams& g, int batch, int flags, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t gemm_tcgen05_workspace_bytes(int M, long N, int K, int batch, int flags);
bool gemm_tcgen05_supports(const GemmParams& g, int batch, int flags);

}  // namespace zutis
