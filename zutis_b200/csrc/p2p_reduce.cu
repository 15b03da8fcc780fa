// All-reduce of the confusion matrix over NVLink peer memory (sm_100a, one process per GPU of one node).
//
// The path's one collective is the sum of the per-GPU int64 Q x Q matrices.  For Q = 81 that is 52 KB,
// and a library all-reduce of that size is pure latency: 18-30 us through NCCL on 2-8 B200s, more than the kernels of an
// 8-image shard (bench.py `strong`).  Here every rank stages its matrix in a buffer its peers have mapped (CUDA IPC), and ONE
// kernel per rank does the whole exchange:
//     stage   local matrix -> local staging buffer (peers read it there)
//     arrive  store this call's epoch into flag[rank] of every peer (st.release.sys over NVLink), wait until every
//             peer's epoch has arrived in the local flags (ld.acquire.sys)
//     sum     out[i] = sum over ranks r = 0..world-1 of staging_r[i]   (coalesced peer loads; fixed order, integers)
// Two staging buffers are used in turns: a peer that has arrived for call e has finished reading call e-1, so restaging
// buffer e % 2 at call e is safe without a second handshake.  Eight CTAs per rank (one element per thread for Q = 81) that meet through monotonic device-memory counters.
// The epoch lives in device memory and is advanced by the kernel itself, so the launch can be captured in a CUDA graph and
// replayed.  Every rank must make the same sequence of calls (as with any collective).  Meant for small matrices; the
// caller keeps NCCL for large ones (Q = 920: 6.8 MB, where a ring moves a quarter of the bytes an all-read does).
#include "common.cuh"

#include <mutex>

using namespace zutis;

namespace {

constexpr int kMaxWorld = 16;
constexpr int kFlagWords = 2 * kMaxWorld;              // arrive[world] | leave[world]

struct P2PContext {
    int world = 0, rank = 0, device = 0;
    size_t data_bytes = 0;                             // staging area of each rank's block; the flags follow it
    char* local = nullptr;                             // this rank's block (cudaMalloc)
    char* peers[kMaxWorld] = {};                       // every rank's block as mapped here (peers[rank] == local)
    char** d_peers = nullptr;                          // device copy of peers[]
    unsigned* d_epoch = nullptr;
    bool connected = false;
};

std::mutex g_mu;
P2PContext* g_ctx[64] = {};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int kCtas = 8;                               // co-resident by a wide margin; they meet through device-memory counters

// signal every peer (flag[which][rank] = epoch in the peer's block) and wait for every peer's signal in the local block
__device__ __forceinline__ void flag_round(char* const* peers, size_t data_bytes, int rank, int world, unsigned epoch, int which) {
    if ((int)threadIdx.x < world) {
        unsigned* theirs = reinterpret_cast<unsigned*>(peers[threadIdx.x] + data_bytes) + which * kMaxWorld + rank;
        st_release_sys(theirs, epoch);
    }
}
__device__ __forceinline__ void flag_wait(char* const* peers, size_t data_bytes, int rank, int world, unsigned epoch, int which) {
    if ((int)threadIdx.x < world) {
        const unsigned* mine = reinterpret_cast<const unsigned*>(peers[rank] + data_bytes) + which * kMaxWorld + threadIdx.x;
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {}
    }
    __syncthreads();
}

// counters: [0] epoch of the last finished call, [1] CTAs that have staged (monotonic), [2] CTAs that have summed (monotonic)
__global__ void __launch_bounds__(1024) p2p_allreduce_kernel(char* const* __restrict__ peers, size_t data_bytes, int rank, int world,
                                                             long long* __restrict__ hist, int* __restrict__ partial, long n2,
                                                             long long* __restrict__ out, unsigned* counters) {
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(counters) + 1u;      // advanced by the last CTA, at the very end
    const long stride = (long)gridDim.x * blockDim.x, first = (long)blockIdx.x * blockDim.x + threadIdx.x;
    // Two staging buffers, taken in turns: a peer that has arrived for call e has finished reading call e-1, so when this
    // rank restages buffer e % 2 at call e, every peer is done with what call e-2 left there -- no second handshake.
    const size_t half = data_bytes / 2, mine_off = (epoch & 1u) ? half : 0;
    long long* staging = reinterpret_cast<long long*>(peers[rank] + mine_off);
    if (partial) {
        // the "+=" of running_score.py:20 for the launches since the last merge, folded into the staging pass
        for (long i = first; i < n2; i += stride) {
            const long long v = hist[i] + (long long)partial[i];
            hist[i] = v; partial[i] = 0; staging[i] = v;
        }
    } else {
        for (long i = first; i < n2; i += stride) staging[i] = hist[i];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(counters + 1, 1u);
    if (blockIdx.x == 0) {
        // the whole matrix is staged (every CTA has counted itself in) before any peer is told so
        if (threadIdx.x == 0)
            while ((int)(*reinterpret_cast<volatile unsigned*>(counters + 1) - epoch * gridDim.x) < 0) {}
        __syncthreads();
        __threadfence();
        flag_round(peers, data_bytes, rank, world, epoch, 0);        // st.release.sys: cumulative over what was fenced above
    }
    flag_wait(peers, data_bytes, rank, world, epoch, 0);
    for (long i = first; i < n2; i += stride) {
        long long sum = 0;
        for (int r = 0; r < world; ++r) sum += reinterpret_cast<const volatile long long*>(peers[r] + mine_off)[i];
        out[i] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(counters + 2, 1u) == epoch * gridDim.x - 1u)
        *reinterpret_cast<volatile unsigned*>(counters) = epoch;      // last CTA of this call
}

// frees what a context owns (not the peers' mappings) and the context itself
void release(P2PContext* c) {
    if (c->local) cudaFree(c->local);
    if (c->d_peers) cudaFree(c->d_peers);
    if (c->d_epoch) cudaFree(c->d_epoch);
    (void)cudaGetLastError();
    delete c;
}

}  // namespace

// Step 1 of 2: allocate this rank's block and export its IPC handle (64 bytes) for the peers.
extern "C" int zutis_p2p_create(int world, int rank, long max_n2, unsigned char* ipc_handle_out, int* ctx_out) {
    ZUTIS_REQUIRE(ipc_handle_out && ctx_out, "zutis_p2p_create: NULL pointer");
    ZUTIS_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "zutis_p2p_create: world=%d rank=%d", world, rank);
    ZUTIS_REQUIRE(max_n2 > 0 && max_n2 <= (1L << 22), "zutis_p2p_create: max_n2=%ld", max_n2);
    int st = current_device_ok();
    if (st != ZUTIS_OK) return st;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    P2PContext* c = new P2PContext();
    c->world = world; c->rank = rank;
    cudaGetDevice(&c->device);
    c->data_bytes = 2 * (((size_t)max_n2 * 8 + 255) & ~(size_t)255);      // two staging buffers
    const size_t total = c->data_bytes + kFlagWords * sizeof(unsigned);
    if (cudaMalloc(&c->local, total) != cudaSuccess || cudaMemset(c->local, 0, total) != cudaSuccess ||
        cudaMalloc(&c->d_peers, kMaxWorld * sizeof(char*)) != cudaSuccess || cudaMalloc(&c->d_epoch, 4 * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(c->d_epoch, 0, 4 * sizeof(unsigned)) != cudaSuccess) {
        const int rc = check_cuda(cudaGetLastError(), "zutis_p2p_create allocation");
        release(c);
        return rc != ZUTIS_OK ? rc : fail(ZUTIS_ERR_CUDA, "zutis_p2p_create: allocation failed");
    }
    cudaIpcMemHandle_t h;
    st = check_cuda(cudaIpcGetMemHandle(&h, c->local), "cudaIpcGetMemHandle");
    if (st != ZUTIS_OK) { release(c); return st; }
    memcpy(ipc_handle_out, &h, 64);
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < 64; ++i)
        if (!g_ctx[i]) { g_ctx[i] = c; *ctx_out = i; return ZUTIS_OK; }
    release(c);
    return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_p2p_create: too many contexts");
}

// Step 2 of 2: map every peer's block.  handles: world x 64 bytes, in rank order (this rank's own entry is ignored).
extern "C" int zutis_p2p_connect(int ctx, const unsigned char* handles) {
    ZUTIS_REQUIRE(ctx >= 0 && ctx < 64 && g_ctx[ctx] && handles, "zutis_p2p_connect: bad context");
    P2PContext* c = g_ctx[ctx];
    ZUTIS_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->peers[r] = c->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        int st = check_cuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
        if (st != ZUTIS_OK) return st;
        c->peers[r] = (char*)p;
    }
    ZUTIS_CUDA(cudaMemcpy(c->d_peers, c->peers, kMaxWorld * sizeof(char*), cudaMemcpyHostToDevice));
    c->connected = true;
    return ZUTIS_OK;
}

extern "C" int zutis_merge_allreduce_hist_p2p(int ctx, long long* hist, int32_t* partial, long n2, long long* out, void* stream);

// out[i] = sum over ranks of hist_r[i], i < n2; hist and out are this rank's device buffers (they may be the same buffer:
// the matrix is staged before it is summed).  Enqueued on `stream`; graph-capturable.
extern "C" int zutis_allreduce_hist_p2p(int ctx, const long long* hist, long n2, long long* out, void* stream) {
    return zutis_merge_allreduce_hist_p2p(ctx, const_cast<long long*>(hist), nullptr, n2, out, stream);
}

// The same with the pending int32 partial of this rank folded in first (hist += partial; partial = 0), i.e.
// zutis_hist_merge + zutis_allreduce_hist_p2p in one launch.  out must not be hist here.
extern "C" int zutis_merge_allreduce_hist_p2p(int ctx, long long* hist, int32_t* partial, long n2, long long* out, void* stream) {
    ZUTIS_REQUIRE(ctx >= 0 && ctx < 64 && g_ctx[ctx], "zutis_allreduce_hist_p2p: bad context");
    P2PContext* c = g_ctx[ctx];
    ZUTIS_REQUIRE(c->connected, "zutis_allreduce_hist_p2p: zutis_p2p_connect has not been called");
    ZUTIS_REQUIRE(hist && out && n2 > 0 && (size_t)n2 * 8 <= c->data_bytes / 2, "zutis_allreduce_hist_p2p: n2=%ld does not fit the context", n2);
    ZUTIS_REQUIRE(!partial || out != hist, "zutis_merge_allreduce_hist_p2p: out must not alias hist");
    p2p_allreduce_kernel<<<kCtas, 1024, 0, (cudaStream_t)stream>>>(c->d_peers, c->data_bytes, c->rank, c->world, hist, partial, n2, out, c->d_epoch);
    return check_launch("p2p_allreduce_kernel");
}

extern "C" int zutis_p2p_destroy(int ctx) {
    ZUTIS_REQUIRE(ctx >= 0 && ctx < 64 && g_ctx[ctx], "zutis_p2p_destroy: bad context");
    P2PContext* c = g_ctx[ctx];
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_ctx[ctx] = nullptr;
    }
    release(c);
    return ZUTIS_OK;
}
