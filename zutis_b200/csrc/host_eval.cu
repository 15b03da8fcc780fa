// End-to-end entry with host buffers: copy in -> contraction -> fused decode+score -> merge -> copy out.
//
// This is the call a holder of host (numpy) arrays makes -- the reference hands host int64 ground truth
// to RunningScore.update (trainer.py:320,347) -- and the leg bench.py reports as `e2e`.  Images are
// processed in chunks on two streams so the host->device copy of chunk i+1 overlaps the kernels of
// chunk i.  Streams, events, device scratch and the pinned read-back buffer live in a per-device context that is
// created on first use and only ever grows: a call enqueues copies and kernels, nothing else (no stream creation, no
// allocation, no synchronisation before the final read-back), which is what matters for small batches.
//
// Ground truth crosses PCIe in the narrowest type that keeps the confusion matrix unchanged: the kernels ignore every
// label outside [0, n) (running_score.py:11), so an int64 / int32 label v becomes (0 <= v < n) ? v : sentinel in uint8
// (n <= 255, sentinel 255) or int16 (sentinel -1).  A few host threads do that chunk by chunk into a pinned staging
// buffer while the (8x larger) token copy of the same chunk is already on the wire; for the reference's int64 labels
// this removes 17 % of the bytes of a step, and the copy engine is what bounds the call.
#include "gemm.cuh"

#include <sched.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

using namespace zutis;

namespace {

struct Buffer {
    void* p = nullptr;
    size_t bytes = 0;
    // grow-only device buffer; returns false on allocation failure
    bool ensure(size_t need, bool zero = false) {
        if (need <= bytes) return true;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, need) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        if (zero && cudaMemset(p, 0, need) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        bytes = need;
        return true;
    }
};

struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    Buffer tokens, gt, logits, labels, partial, gemm_ws;
    Buffer decode_ws;                // run counter of the cell decode kernel: zeroed at allocation, left zero by every launch
};

struct HostContext {
    std::mutex mu;                   // one call at a time per device
    bool ready = false;
    Lane lanes[2];
    cudaEvent_t text_ready = nullptr;
    Buffer text, hist;
    long long* h_hist = nullptr;     // pinned
    long h_hist_n = 0;
    void* h_gt = nullptr;            // pinned staging of the narrowed ground truth
    size_t h_gt_bytes = 0;
};

// bytes per ground-truth pixel after narrowing for n classes (0: leave the caller's type alone)
int narrowed_gt_code(int gt_dtype, int n) {
    const int have = gt_dtype_bytes(gt_dtype);
    if (n <= 255 && have > 1) return ZUTIS_GT_U8;
    if (n <= 32767 && have > 2) return ZUTIS_GT_I16;
    return gt_dtype;
}

template <typename SRC, typename DST>
void narrow_range(const SRC* src, DST* dst, size_t count, int n, DST sentinel) {
    for (size_t i = 0; i < count; ++i) {
        const SRC v = src[i];
        dst[i] = (v >= 0 && v < (SRC)n) ? (DST)v : sentinel;
    }
}

void narrow_labels(const void* src, int src_code, void* dst, int dst_code, size_t first, size_t count, int n) {
    if (dst_code == ZUTIS_GT_U8) {
        uint8_t* d = (uint8_t*)dst + first;
        if (src_code == ZUTIS_GT_I64) narrow_range((const long long*)src + first, d, count, n, (uint8_t)255);
        else if (src_code == ZUTIS_GT_I32) narrow_range((const int32_t*)src + first, d, count, n, (uint8_t)255);
        else narrow_range((const int16_t*)src + first, d, count, n, (uint8_t)255);
    } else {
        int16_t* d = (int16_t*)dst + first;
        if (src_code == ZUTIS_GT_I64) narrow_range((const long long*)src + first, d, count, n, (int16_t)-1);
        else narrow_range((const int32_t*)src + first, d, count, n, (int16_t)-1);
    }
}

int usable_cpus();

// Host threads that narrow the labels of one call.  ZUTIS_HOST_THREADS overrides; otherwise the CPUs this process may use
// are shared out among the ranks of the node (LOCAL_WORLD_SIZE, set by torchrun), one is left to the enqueueing thread,
// and more than four do not help (the narrowing of a chunk only has to keep ahead of the 8x larger token copy).
int narrowing_threads() {
    if (const char* e = getenv("ZUTIS_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 64) return v;
    }
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    int n = usable_cpus() / ranks - 1;
    return n < 1 ? 1 : (n > 4 ? 4 : n);
}

int usable_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) != 0) return 1;
    const int n = CPU_COUNT(&set);
    return n > 0 ? n : 1;
}

HostContext g_ctx[64];

int init_context(HostContext& c) {
    if (c.ready) return ZUTIS_OK;
    for (int l = 0; l < 2; ++l) {
        ZUTIS_CUDA(cudaStreamCreateWithFlags(&c.lanes[l].stream, cudaStreamNonBlocking));
        ZUTIS_CUDA(cudaEventCreateWithFlags(&c.lanes[l].done, cudaEventDisableTiming));
    }
    ZUTIS_CUDA(cudaEventCreateWithFlags(&c.text_ready, cudaEventDisableTiming));
    c.ready = true;
    return ZUTIS_OK;
}

}  // namespace

extern "C" int zutis_semantic_eval_host(const float* text, const float* tokens, const void* gt, int gt_dtype,
                                        int B, int Q, int D, int h, int w, int H, int W,
                                        long long* hist_host, int16_t* labels_host, int gemm_flags, int device) {
    ZUTIS_REQUIRE(text && tokens, "zutis_semantic_eval_host: NULL input");
    ZUTIS_REQUIRE(hist_host || labels_host, "zutis_semantic_eval_host: nothing to produce");
    ZUTIS_REQUIRE(!hist_host || gt, "zutis_semantic_eval_host: hist requested without gt");
    ZUTIS_REQUIRE(B > 0 && Q > 0 && D > 0 && h > 0 && w > 0 && H > 0 && W > 0, "zutis_semantic_eval_host: bad shape");
    ZUTIS_REQUIRE(device >= 0 && device < 64, "zutis_semantic_eval_host: device %d out of range", device);
    const int src_gt_bytes = gt_dtype_bytes(gt_dtype);
    ZUTIS_REQUIRE(!gt || src_gt_bytes > 0, "zutis_semantic_eval_host: bad gt_dtype %d", gt_dtype);
    // ground truth is only read for the histogram; it crosses PCIe narrowed (see the header of this file)
    const int src_gt_dtype = gt_dtype;
    if (gt && hist_host) gt_dtype = narrowed_gt_code(gt_dtype, Q);
    const bool narrow = gt && gt_dtype != src_gt_dtype;
    const int gt_bytes = gt ? gt_dtype_bytes(gt_dtype) : 0;
    ZUTIS_CUDA(cudaSetDevice(device));
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    HostContext& c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);
    st = init_context(c);
    if (st != ZUTIS_OK) return st;

    const long hw = (long)h * w, HW = (long)H * W;
    const int Qp = (Q + 3) & ~3;                       // pixel-major logits row, 16-byte aligned
    const long n2 = (long)Q * Q;
    // chunk so that one chunk's tokens are ~32 MB: big enough to run PCIe at full rate, small enough to overlap
    long chunk = (32L << 20) / (hw * D * 4);
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    const int nlanes = chunk < B ? 2 : 1;
    const size_t ws_bytes = zutis_gemm_workspace_bytes(Q, hw, D, (int)chunk, gemm_flags);
    const size_t dws_bytes = zutis_decode_workspace_bytes((int)chunk, Q, h, w, H, W);

    // ---- scratch (grow-only; allocation happens on the first call of a shape, never in steady state)
    bool ok = c.text.ensure((size_t)Q * D * 4) && (!hist_host || c.hist.ensure((size_t)n2 * 8));
    for (int l = 0; l < nlanes && ok; ++l) {
        Lane& L = c.lanes[l];
        ok = L.tokens.ensure((size_t)chunk * hw * D * 4) && L.logits.ensure((size_t)chunk * hw * Qp * 4) &&
             (!gt || L.gt.ensure((size_t)chunk * HW * gt_bytes)) && (!labels_host || L.labels.ensure((size_t)chunk * HW * 2)) &&
             (!hist_host || L.partial.ensure((size_t)n2 * 4)) && (!ws_bytes || L.gemm_ws.ensure(ws_bytes)) &&
             L.decode_ws.ensure(dws_bytes, /*zero=*/true);
    }
    if (!ok) return fail(ZUTIS_ERR_CUDA, "zutis_semantic_eval_host: device allocation failed");
    if (hist_host && c.h_hist_n < n2) {
        if (c.h_hist) cudaFreeHost(c.h_hist);
        c.h_hist = nullptr; c.h_hist_n = 0;
        ZUTIS_CUDA(cudaMallocHost((void**)&c.h_hist, (size_t)n2 * 8));
        c.h_hist_n = n2;
    }
    if (narrow && c.h_gt_bytes < (size_t)B * HW * gt_bytes) {
        if (c.h_gt) cudaFreeHost(c.h_gt);
        c.h_gt = nullptr; c.h_gt_bytes = 0;
        ZUTIS_CUDA(cudaMallocHost(&c.h_gt, (size_t)B * HW * gt_bytes));
        c.h_gt_bytes = (size_t)B * HW * gt_bytes;
    }
    // narrowing runs ahead of the copies, chunk by chunk: worker t converts slice t of every chunk and bumps the chunk's
    // counter; the enqueueing thread waits for a chunk's counter before it hands that chunk to the copy engine
    const long n_chunks = (B + chunk - 1) / chunk;
    int n_workers = 0;
    std::vector<std::atomic<int>> narrowed(narrow ? n_chunks : 0);
    std::vector<std::thread> workers;
    if (narrow) {
        for (auto& a : narrowed) a.store(0, std::memory_order_relaxed);
        const size_t total = (size_t)B * HW;
        n_workers = total * src_gt_bytes < (4u << 20) ? 0 : narrowing_threads();
        void* staging = c.h_gt;
        auto work = [=, &narrowed](int t, int T) {
            for (long k = 0; k < n_chunks; ++k) {
                const size_t first = (size_t)k * chunk * HW;
                const size_t count = (size_t)((B - k * chunk < chunk) ? (B - k * chunk) : chunk) * HW;
                const size_t lo = first + count * t / T, hi = first + count * (t + 1) / T;
                narrow_labels(gt, src_gt_dtype, staging, gt_dtype, lo, hi - lo, Q);
                narrowed[k].fetch_add(1, std::memory_order_release);
            }
        };
        if (n_workers == 0) { work(0, 1); n_workers = 1; }          // small batch: done before the first copy, no threads
        else for (int t = 0; t < n_workers; ++t) workers.emplace_back(work, t, n_workers);
    }
    struct Joiner { std::vector<std::thread>& w; ~Joiner() { for (auto& t : w) if (t.joinable()) t.join(); } } joiner{workers};

    int rc = ZUTIS_OK;
    auto guard = [&](int s) { if (rc == ZUTIS_OK && s != ZUTIS_OK) rc = s; return rc == ZUTIS_OK; };
    cudaStream_t s0 = c.lanes[0].stream;
    float* d_text = (float*)c.text.p;
    guard(check_cuda(cudaMemcpyAsync(d_text, text, (size_t)Q * D * 4, cudaMemcpyHostToDevice, s0), "H2D text"));
    guard(check_cuda(cudaEventRecord(c.text_ready, s0), "cudaEventRecord"));
    if (hist_host) {
        guard(check_cuda(cudaMemsetAsync(c.hist.p, 0, (size_t)n2 * 8, s0), "cudaMemsetAsync hist"));
        for (int l = 0; l < nlanes; ++l) guard(check_cuda(cudaMemsetAsync(c.lanes[l].partial.p, 0, (size_t)n2 * 4, c.lanes[l].stream), "cudaMemsetAsync partial"));
    }
    for (int l = 1; l < nlanes; ++l) guard(check_cuda(cudaStreamWaitEvent(c.lanes[l].stream, c.text_ready, 0), "cudaStreamWaitEvent"));

    bool prepared[2] = {false, false};                 // the text operand's hi/lo split is made once per lane and call
    // everything of chunk `it` after its token copy: labels in, contraction, decode + score, labels out
    auto finish = [&](long it) {
        Lane& L = c.lanes[it % nlanes];
        const long b0 = it * chunk;
        const int nb = (int)((B - b0 < chunk) ? (B - b0) : chunk);
        if (gt && rc == ZUTIS_OK) {
            const char* src = (const char*)gt;
            if (narrow) {
                while (narrowed[it].load(std::memory_order_acquire) < n_workers) std::this_thread::yield();
                src = (const char*)c.h_gt;
            }
            guard(check_cuda(cudaMemcpyAsync(L.gt.p, src + (size_t)b0 * HW * gt_bytes, (size_t)nb * HW * gt_bytes, cudaMemcpyHostToDevice, L.stream), "H2D gt"));
        }
        if (rc != ZUTIS_OK) return;
        const bool tensor_core = (gemm_flags & ZUTIS_GEMM_PRECISION_MASK) != ZUTIS_GEMM_FP32_SIMT;
        const int fl = gemm_flags | ((tensor_core && prepared[it % nlanes] && nb == chunk) ? ZUTIS_GEMM_A_PREPARED : 0);
        guard(zutis_gemm_logits(d_text, D, 0, (const float*)L.tokens.p, D, hw * D, (float*)L.logits.p, 1, Qp, hw * Qp, Q, hw, D, nb, fl,
                                L.gemm_ws.p, ws_bytes, L.stream));
        if (rc != ZUTIS_OK) return;
        prepared[it % nlanes] = (nb == chunk);
        guard(zutis_decode_score_ws((const float*)L.logits.p, hw * Qp, 1, (long)w * Qp, Qp, nb, Q, h, w, H, W, L.gt.p, gt_dtype, HW,
                                    labels_host ? (int16_t*)L.labels.p : nullptr, hist_host ? (int32_t*)L.partial.p : nullptr, Q,
                                    ZUTIS_DECODE_AUTO | ZUTIS_DECODE_WORKSPACE_ZEROED, L.decode_ws.p, dws_bytes, L.stream));
        if (labels_host && rc == ZUTIS_OK)
            guard(check_cuda(cudaMemcpyAsync(labels_host + (size_t)b0 * HW, L.labels.p, (size_t)nb * HW * 2, cudaMemcpyDeviceToHost, L.stream), "D2H labels"));
    };
    // The token copy of chunk i goes to the copy engine BEFORE chunk i-1 is finished: if this thread has to wait for chunk
    // i-1's narrowed labels, the wire is busy with chunk i's tokens meanwhile.
    for (long it = 0; it < n_chunks && rc == ZUTIS_OK; ++it) {
        Lane& L = c.lanes[it % nlanes];
        const long b0 = it * chunk;
        const int nb = (int)((B - b0 < chunk) ? (B - b0) : chunk);
        guard(check_cuda(cudaMemcpyAsync(L.tokens.p, tokens + (size_t)b0 * hw * D, (size_t)nb * hw * D * 4, cudaMemcpyHostToDevice, L.stream), "H2D tokens"));
        if (it > 0 && rc == ZUTIS_OK) finish(it - 1);
    }
    if (rc == ZUTIS_OK) finish(n_chunks - 1);
    // lane 1 -> lane 0, then fold both lanes' int32 partials into one int64 matrix and bring it home
    for (int l = 1; l < nlanes && rc == ZUTIS_OK; ++l) {
        guard(check_cuda(cudaEventRecord(c.lanes[l].done, c.lanes[l].stream), "cudaEventRecord"));
        guard(check_cuda(cudaStreamWaitEvent(s0, c.lanes[l].done, 0), "cudaStreamWaitEvent"));
    }
    if (hist_host && rc == ZUTIS_OK) {
        for (int l = 0; l < nlanes && rc == ZUTIS_OK; ++l)
            guard(zutis_hist_merge((int32_t*)c.lanes[l].partial.p, 1, (long long*)c.hist.p, n2, 0, s0));
        if (rc == ZUTIS_OK) guard(check_cuda(cudaMemcpyAsync(c.h_hist, c.hist.p, (size_t)n2 * 8, cudaMemcpyDeviceToHost, s0), "D2H hist"));
    }
    // the one synchronisation of the call (also after an error: the context's buffers must be idle when we return)
    for (int l = 0; l < nlanes; ++l) {
        const cudaError_t e = cudaStreamSynchronize(c.lanes[l].stream);
        if (e != cudaSuccess) guard(check_cuda(e, "cudaStreamSynchronize"));
    }
    if (hist_host && rc == ZUTIS_OK)
        for (long i = 0; i < n2; ++i) hist_host[i] += c.h_hist[i];
    return rc;
}

// Bytes the call above moves host -> device for one batch (text + tokens + ground truth as it crosses PCIe).
extern "C" size_t zutis_semantic_eval_host_h2d_bytes(int gt_dtype, int want_hist, int B, int Q, int D, int h, int w, int H, int W) {
    const int code = want_hist ? narrowed_gt_code(gt_dtype, Q) : gt_dtype;
    const int gb = gt_dtype_bytes(code);
    return (size_t)Q * D * 4 + (size_t)B * h * w * D * 4 + (size_t)B * H * W * (gb > 0 ? gb : 0);
}
