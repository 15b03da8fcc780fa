// the tail of ZUTIS.image_to_text_space (networks/zutis.py:321-322) on the device:
//   y = F.layer_norm(y, y.shape[1:])                     joint statistics over (h, w, c) of each image, eps 1e-5
//   y = y / (y.norm(dim=-1, keepdim=True) + 1e-7)        per-pixel L2 normalisation
// The projection in front of it (zutis.py:319) is the contraction kernel with A = proj^T.
//
// moments_partial_kernel : per image, kMomBlocks blocks each reduce a contiguous slice to (sum, sum of squares) in
//                          double, fixed order -> workspace [B][kMomBlocks][2]
// normalize_kernel       : every block first folds its image's partials in a fixed order (deterministic mean / rstd),
//                          then one warp per pixel: load D channels, (x - mean) * rstd, sum of squares by shuffle tree,
//                          divide by (norm + eps), store.  Reads x twice and writes it once: HBM-bound.
#include "common.cuh"

namespace zutis {

constexpr int kMomBlocks = 64;

__global__ void __launch_bounds__(256) moments_partial_kernel(const float* x, long per_image, double* partial) {
    const int b = blockIdx.y;
    const long chunk = (per_image + kMomBlocks - 1) / kMomBlocks;
    const long lo = (long)blockIdx.x * chunk;
    const long hi = lo + chunk < per_image ? lo + chunk : per_image;
    const float* p = x + (long)b * per_image;
    double s = 0.0, ss = 0.0;
    for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double v = (double)p[i];
        s += v; ss += v * v;
    }
    __shared__ double rs[8], rq[8];
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < 8; ++k) { a += rs[k]; c += rq[k]; }
        partial[((long)b * kMomBlocks + blockIdx.x) * 2] = a;
        partial[((long)b * kMomBlocks + blockIdx.x) * 2 + 1] = c;
    }
}

__global__ void __launch_bounds__(256) normalize_kernel(float* x, long pixels, int D, const double* partial, int layer_norm,
                                                        float ln_eps, float l2_eps) {
    const int b = blockIdx.y;
    __shared__ float s_mean, s_rstd;
    if (threadIdx.x == 0) {
        float mean = 0.0f, rstd = 1.0f;
        if (layer_norm) {
            double s = 0.0, ss = 0.0;
            for (int k = 0; k < kMomBlocks; ++k) { s += partial[((long)b * kMomBlocks + k) * 2]; ss += partial[((long)b * kMomBlocks + k) * 2 + 1]; }
            const double n = (double)pixels * D;
            const double m = s / n;
            double var = ss / n - m * m;                       // biased variance, as F.layer_norm
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            rstd = (float)(1.0 / sqrt(var + (double)ln_eps));
        }
        s_mean = mean; s_rstd = rstd;
    }
    __syncthreads();
    const float mean = s_mean, rstd = s_rstd;
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long px = warp0; px < pixels; px += nwarps) {
        float* row = x + ((long)b * pixels + px) * D;
        float ss = 0.0f;
        for (int d = lane; d < D; d += 32) {
            const float v = (row[d] - mean) * rstd;
            ss = __fmaf_rn(v, v, ss);
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float inv = 1.0f / (sqrtf(ss) + l2_eps);
        for (int d = lane; d < D; d += 32) row[d] = ((row[d] - mean) * rstd) * inv;
    }
}

}  // namespace zutis

using namespace zutis;

extern "C" size_t zutis_image_norm_workspace_bytes(int B, long, int) { return (size_t)(B > 0 ? B : 0) * kMomBlocks * 2 * sizeof(double); }

extern "C" int zutis_image_layernorm_l2norm(float* x, int B, long pixels, int D, int layer_norm, float ln_eps, float l2_eps,
                                            void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ZUTIS_REQUIRE(x != nullptr, "zutis_image_layernorm_l2norm: x is NULL");
    ZUTIS_REQUIRE(B > 0 && pixels > 0 && D > 0 && B <= 65535, "zutis_image_layernorm_l2norm: bad shape B=%d pixels=%ld D=%d", B, pixels, D);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    double* partial = reinterpret_cast<double*>(workspace);
    if (layer_norm) {
        if (!workspace || workspace_bytes < zutis_image_norm_workspace_bytes(B, pixels, D))
            return fail(ZUTIS_ERR_WORKSPACE, "zutis_image_layernorm_l2norm: workspace of %zu bytes needed", zutis_image_norm_workspace_bytes(B, pixels, D));
        ZUTIS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "zutis_image_layernorm_l2norm: workspace must be 8-byte aligned");
        moments_partial_kernel<<<dim3(kMomBlocks, B), 256, 0, stream>>>(x, pixels * D, partial);
        st = check_launch("moments_partial_kernel");
        if (st != ZUTIS_OK) return st;
    }
    long blocks = (pixels + 7) / 8;
    const long cap = ((long)sm_count() * 8 + B - 1) / B;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    normalize_kernel<<<dim3((unsigned)blocks, B), 256, 0, stream>>>(x, pixels, D, partial, layer_norm, ln_eps, l2_eps);
    return check_launch("normalize_kernel");
}
