// Confusion-matrix kernels that stand alone (sm_100a):
//   score_labels_kernel : RunningScore._fast_hist (utils/running_score.py:10-16) for labels that
//                         already exist on the device (RunningScore.update with CUDA tensors).
//   hist_merge_kernel   : folds int32 per-launch partial matrices into the persistent int64 matrix
//                         (the "+=" of running_score.py:20) and clears the partials for reuse.
#include "common.cuh"

#include <mutex>

namespace zutis {

__global__ void __launch_bounds__(256) score_labels_kernel(const void* gt, int gt_dtype, const void* pred, int pred_dtype,
                                                           long npix, int* hist, int n, int in_smem) {
    extern __shared__ int s_hist[];
    const int nn = n * n;
    if (in_smem) {
        for (int i = threadIdx.x; i < nn; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    int* target = in_smem ? s_hist : hist;
    const int lane = threadIdx.x & 31;
    const long nchunks = (npix + 31) >> 5;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long c = warp0; c < nchunks; c += nwarps) {
        const long i = c * 32 + lane;
        int key = -1;
        if (i < npix) {
            const long long g = load_label(gt, gt_dtype, (size_t)i);
            const long long q = load_label(pred, pred_dtype, (size_t)i);
            // The reference does not range-check predictions (they would alias into other bins);
            // predictions outside [0,n) are dropped here instead.
            if (g >= 0 && g < n && q >= 0 && q < n) key = (int)g * n + (int)q;
        }
        warp_hist_add(target, key);
    }
    if (in_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            const int v = s_hist[i];
            if (v) atomicAdd(hist + i, v);
        }
    }
}

__global__ void __launch_bounds__(256) hist_merge_kernel(int* partials, int n_partials, long long* hist, long n2, int clear) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
        long long acc = 0;
        for (int j = 0; j < n_partials; ++j) {
            const long o = (long)j * n2 + i;
            acc += (long long)partials[o];
            if (clear) partials[o] = 0;
        }
        if (acc) hist[i] += acc;
    }
}

}  // namespace zutis

using namespace zutis;

extern "C" int zutis_score_labels(const void* gt, int gt_dtype, const void* pred, int pred_dtype,
                                  long n_pixels, int32_t* hist_partial, int n_classes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ZUTIS_REQUIRE(gt && pred && hist_partial, "zutis_score_labels: NULL pointer");
    ZUTIS_REQUIRE(gt_dtype_bytes(gt_dtype) > 0 && gt_dtype_bytes(pred_dtype) > 0, "zutis_score_labels: bad dtype");
    ZUTIS_REQUIRE(n_pixels >= 0 && n_pixels < 2147483647L, "zutis_score_labels: n_pixels=%ld out of range", n_pixels);
    ZUTIS_REQUIRE(n_classes > 0 && n_classes <= 46340, "zutis_score_labels: bad n_classes %d", n_classes);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    if (n_pixels == 0) return ZUTIS_OK;
    const int nn = n_classes * n_classes;
    const int in_smem = nn * 4 <= 64 * 1024;
    const size_t smem = in_smem ? (size_t)nn * 4 : 0;
    if (smem > 48 * 1024)
        ZUTIS_CUDA(cudaFuncSetAttribute(score_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long blocks = (n_pixels + 2047) / 2048;
    const long cap = (long)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    score_labels_kernel<<<(unsigned)blocks, 256, smem, stream>>>(gt, gt_dtype, pred, pred_dtype, n_pixels, hist_partial, n_classes, in_smem);
    return check_launch("score_labels_kernel");
}

extern "C" int zutis_hist_merge(int32_t* partials, int n_partials, long long* hist_i64, long n2,
                                int clear_partials, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ZUTIS_REQUIRE(partials && hist_i64, "zutis_hist_merge: NULL pointer");
    ZUTIS_REQUIRE(n_partials > 0 && n2 > 0, "zutis_hist_merge: n_partials=%d n2=%ld", n_partials, n2);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    long blocks = (n2 + 255) / 256;
    const long cap = (long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    hist_merge_kernel<<<(unsigned)blocks, 256, 0, stream>>>(partials, n_partials, hist_i64, n2, clear_partials);
    return check_launch("hist_merge_kernel");
}

// ------------------------------------------------------------------------------------------------- multi-GPU sum
// One ncclAllReduce(sum, int64, n^2) over the caller's communicator.  The library does not link
// NCCL: the symbol is taken from the NCCL the host process has already loaded (torch's, or the host application's own), so the
// communicator and the library always belong to the same NCCL build.
#include <dlfcn.h>

namespace {
typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
NcclAllReduceFn resolve_nccl_allreduce() {
    static NcclAllReduceFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            if (sym) break;
            if (void* h = dlopen(name, RTLD_NOW | RTLD_NOLOAD)) sym = dlsym(h, "ncclAllReduce");   // already loaded (RTLD_LOCAL)
        }
        fn = reinterpret_cast<NcclAllReduceFn>(sym);
    });
    return fn;
}
}  // namespace

extern "C" int zutis_allreduce_hist(long long* hist_i64, long n2, void* nccl_comm, void* stream) {
    ZUTIS_REQUIRE(hist_i64 && nccl_comm, "zutis_allreduce_hist: NULL pointer");
    ZUTIS_REQUIRE(n2 > 0, "zutis_allreduce_hist: n2=%ld", n2);
    NcclAllReduceFn fn = resolve_nccl_allreduce();
    if (!fn) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_allreduce_hist: no NCCL in this process (the communicator's library must be loaded first)");
    const int nccl_int64 = 4, nccl_sum = 0;                      // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)
    const int rc = fn(hist_i64, hist_i64, (size_t)n2, nccl_int64, nccl_sum, nccl_comm, (cudaStream_t)stream);
    if (rc != 0) return fail(ZUTIS_ERR_CUDA, "zutis_allreduce_hist: ncclAllReduce returned ncclResult_t %d", rc);
    return ZUTIS_OK;
}
