// End-to-end entry with host buffers: copy in -> contraction -> fused decode+score -> merge -> copy out.
//
// This is the call a holder of host (numpy) arrays makes -- the reference hands host int64 ground truth
// to RunningScore.update (trainer.py:320,347) -- and the leg bench.py reports as `e2e`.  Images are
// processed in chunks on two streams so the host->device copy of chunk i+1 overlaps the kernels of
// chunk i.  Device scratch comes from the stream-ordered allocator (the pool keeps it between calls).
#include "gemm.cuh"

#include <mutex>

using namespace zutis;

namespace {

std::once_flag g_pool_once[64];

void keep_pool_memory(int device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
}

struct Lane {
    cudaStream_t stream = nullptr;
    float* tokens = nullptr;
    void* gt = nullptr;
    float* logits = nullptr;
    int16_t* labels = nullptr;
    int32_t* partial = nullptr;
    void* decode_ws = nullptr;       // run counter of the cell decode kernel
};

}  // namespace

extern "C" int zutis_semantic_eval_host(const float* text, const float* tokens, const void* gt, int gt_dtype,
                                        int B, int Q, int D, int h, int w, int H, int W,
                                        long long* hist_host, int16_t* labels_host, int gemm_flags, int device) {
    ZUTIS_REQUIRE(text && tokens, "zutis_semantic_eval_host: NULL input");
    ZUTIS_REQUIRE(hist_host || labels_host, "zutis_semantic_eval_host: nothing to produce");
    ZUTIS_REQUIRE(!hist_host || gt, "zutis_semantic_eval_host: hist requested without gt");
    ZUTIS_REQUIRE(B > 0 && Q > 0 && D > 0 && h > 0 && w > 0 && H > 0 && W > 0, "zutis_semantic_eval_host: bad shape");
    const int gt_bytes = gt_dtype_bytes(gt_dtype);
    ZUTIS_REQUIRE(!gt || gt_bytes > 0, "zutis_semantic_eval_host: bad gt_dtype %d", gt_dtype);
    ZUTIS_CUDA(cudaSetDevice(device));
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    if (device >= 0 && device < 64) std::call_once(g_pool_once[device], keep_pool_memory, device);

    const long hw = (long)h * w, HW = (long)H * W;
    const int Qp = (Q + 3) & ~3;                       // pixel-major logits row, 16-byte aligned
    const long n2 = (long)Q * Q;
    // chunk so that one chunk's tokens are ~32 MB: big enough to run PCIe at full rate, small enough to overlap
    long chunk = (32L << 20) / (hw * D * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    const int nlanes = chunk < B ? 2 : 1;

    Lane lanes[2];
    float* d_text = nullptr;
    long long* d_hist = nullptr;
    void* d_ws = nullptr;
    const size_t ws_bytes = zutis_gemm_workspace_bytes(Q, hw, D, (int)chunk, gemm_flags);
    const size_t dws_bytes = zutis_decode_workspace_bytes((int)chunk, Q, h, w, H, W);
    int rc = ZUTIS_OK;
    auto guard = [&](int s) { if (rc == ZUTIS_OK && s != ZUTIS_OK) rc = s; return rc == ZUTIS_OK; };

    for (int l = 0; l < nlanes && rc == ZUTIS_OK; ++l) {
        Lane& L = lanes[l];
        if (!guard(check_cuda(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking), "cudaStreamCreate"))) break;
        guard(check_cuda(cudaMallocAsync((void**)&L.tokens, (size_t)chunk * hw * D * 4, L.stream), "cudaMallocAsync tokens"));
        guard(check_cuda(cudaMallocAsync((void**)&L.logits, (size_t)chunk * hw * Qp * 4, L.stream), "cudaMallocAsync logits"));
        if (gt) guard(check_cuda(cudaMallocAsync(&L.gt, (size_t)chunk * HW * gt_bytes, L.stream), "cudaMallocAsync gt"));
        if (labels_host) guard(check_cuda(cudaMallocAsync((void**)&L.labels, (size_t)chunk * HW * 2, L.stream), "cudaMallocAsync labels"));
        guard(check_cuda(cudaMallocAsync(&L.decode_ws, dws_bytes, L.stream), "cudaMallocAsync decode workspace"));
        if (hist_host) {
            guard(check_cuda(cudaMallocAsync((void**)&L.partial, (size_t)n2 * 4, L.stream), "cudaMallocAsync partial"));
            if (rc == ZUTIS_OK) guard(check_cuda(cudaMemsetAsync(L.partial, 0, (size_t)n2 * 4, L.stream), "cudaMemsetAsync"));
        }
    }
    cudaStream_t s0 = lanes[0].stream;
    if (rc == ZUTIS_OK) {
        guard(check_cuda(cudaMallocAsync((void**)&d_text, (size_t)Q * D * 4, s0), "cudaMallocAsync text"));
        if (ws_bytes) guard(check_cuda(cudaMallocAsync(&d_ws, ws_bytes * nlanes, s0), "cudaMallocAsync workspace"));
        if (hist_host) {
            guard(check_cuda(cudaMallocAsync((void**)&d_hist, (size_t)n2 * 8, s0), "cudaMallocAsync hist"));
            if (rc == ZUTIS_OK) guard(check_cuda(cudaMemsetAsync(d_hist, 0, (size_t)n2 * 8, s0), "cudaMemsetAsync hist"));
        }
        if (rc == ZUTIS_OK) guard(check_cuda(cudaMemcpyAsync(d_text, text, (size_t)Q * D * 4, cudaMemcpyHostToDevice, s0), "H2D text"));
        if (rc == ZUTIS_OK) guard(check_cuda(cudaStreamSynchronize(s0), "sync text"));
    }

    for (long b0 = 0, it = 0; b0 < B && rc == ZUTIS_OK; b0 += chunk, ++it) {
        Lane& L = lanes[it % nlanes];
        const int nb = (int)((B - b0 < chunk) ? (B - b0) : chunk);
        guard(check_cuda(cudaMemcpyAsync(L.tokens, tokens + (size_t)b0 * hw * D, (size_t)nb * hw * D * 4, cudaMemcpyHostToDevice, L.stream), "H2D tokens"));
        if (gt && rc == ZUTIS_OK)
            guard(check_cuda(cudaMemcpyAsync(L.gt, (const char*)gt + (size_t)b0 * HW * gt_bytes, (size_t)nb * HW * gt_bytes, cudaMemcpyHostToDevice, L.stream), "H2D gt"));
        if (rc != ZUTIS_OK) break;
        guard(zutis_gemm_logits(d_text, D, 0, L.tokens, D, hw * D, L.logits, 1, Qp, hw * Qp, Q, hw, D, nb, gemm_flags,
                                d_ws ? (char*)d_ws + ws_bytes * (it % nlanes) : nullptr, ws_bytes, L.stream));
        if (rc != ZUTIS_OK) break;
        guard(zutis_decode_score_ws(L.logits, hw * Qp, 1, (long)w * Qp, Qp, nb, Q, h, w, H, W, L.gt, gt_dtype, HW,
                                    L.labels, hist_host ? L.partial : nullptr, Q, ZUTIS_DECODE_AUTO, L.decode_ws, dws_bytes, L.stream));
        if (labels_host && rc == ZUTIS_OK)
            guard(check_cuda(cudaMemcpyAsync(labels_host + (size_t)b0 * HW, L.labels, (size_t)nb * HW * 2, cudaMemcpyDeviceToHost, L.stream), "D2H labels"));
    }
    for (int l = 0; l < nlanes; ++l)
        if (lanes[l].stream) guard(check_cuda(cudaStreamSynchronize(lanes[l].stream), "sync lane"));
    if (hist_host && rc == ZUTIS_OK) {
        // fold both lanes' int32 partials into one int64 matrix and bring it home
        for (int l = 0; l < nlanes && rc == ZUTIS_OK; ++l)
            guard(zutis_hist_merge(lanes[l].partial, 1, d_hist, n2, 0, s0));
        static thread_local long long* h_tmp = nullptr;
        static thread_local long h_tmp_n = 0;
        if (rc == ZUTIS_OK && h_tmp_n < n2) {
            if (h_tmp) cudaFreeHost(h_tmp);
            h_tmp = nullptr; h_tmp_n = 0;
            if (guard(check_cuda(cudaMallocHost((void**)&h_tmp, (size_t)n2 * 8), "cudaMallocHost"))) h_tmp_n = n2;
        }
        if (rc == ZUTIS_OK) guard(check_cuda(cudaMemcpyAsync(h_tmp, d_hist, (size_t)n2 * 8, cudaMemcpyDeviceToHost, s0), "D2H hist"));
        if (rc == ZUTIS_OK) guard(check_cuda(cudaStreamSynchronize(s0), "sync hist"));
        if (rc == ZUTIS_OK)
            for (long i = 0; i < n2; ++i) hist_host[i] += h_tmp[i];
    }
    // release (stream-ordered; the pool keeps the memory for the next call)
    for (int l = 0; l < nlanes; ++l) {
        Lane& L = lanes[l];
        if (!L.stream) continue;
        if (L.tokens) cudaFreeAsync(L.tokens, L.stream);
        if (L.logits) cudaFreeAsync(L.logits, L.stream);
        if (L.gt) cudaFreeAsync(L.gt, L.stream);
        if (L.labels) cudaFreeAsync(L.labels, L.stream);
        if (L.partial) cudaFreeAsync(L.partial, L.stream);
        if (L.decode_ws) cudaFreeAsync(L.decode_ws, L.stream);
    }
    if (s0) {
        if (d_text) cudaFreeAsync(d_text, s0);
        if (d_ws) cudaFreeAsync(d_ws, s0);
        if (d_hist) cudaFreeAsync(d_hist, s0);
    }
    for (int l = 0; l < nlanes; ++l)
        if (lanes[l].stream) { cudaStreamSynchronize(lanes[l].stream); cudaStreamDestroy(lanes[l].stream); }
    (void)cudaGetLastError();
    return rc;
}
