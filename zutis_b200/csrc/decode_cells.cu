// Fused upsample -> argmax -> labels -> histogram decode by exact per-cell candidate pruning (sm_100a): the kernel behind
// ZUTIS_DECODE_AUTO / ZUTIS_DECODE_CELLS of zutis_decode_score for pixel-major logits and >= 4x up-sampling.
// Replaces F.interpolate + torch.argmax (networks/zutis.py:366-372) and RunningScore._fast_hist (utils/running_score.py:10-16).
// Same arithmetic as decode_score.cu (bit-exact with oracle/zutis_oracle.c); compiled with -fmad=false.
//
// A CELL is the set of output pixels that share their top-left tap (cy, cx); inside it every interpolant is
//     v_q = fma(ly0, fma(lx0, A_q, lx1*B_q), ly1 * fma(lx0, C_q, lx1*D_q))
// of the cell's four corner logits with non-negative weights, and every rounding in that expression is monotone in the
// taps.  A warp takes a RUN of kRunCells horizontally adjacent cells and works in two phases, in registers and
// warp-private shared memory (no block-level synchronisation after the prologue):
//
//   P  prune, lane = (cell of the run, 1 of 8 slices of the categories).  The run's 2 x 5 low-res pixels (all
//      categories) arrive in the warp's shared memory by ONE TMA box load (cp.async.bulk.tensor.4d, issued by one lane
//      while the previous run is still being evaluated); the lanes of a cell read the four corner rows as float4.
//      Pass 1 finds the cell's dominator k* = argmax_q min_corner L_q (the category with the best guaranteed value).
//      Pass 2 drops every category j with
//          K_c - L_j[c] >= m   at all four corners,   m = 2^-20 * max_c |K_c|   (K = corner values of k*),
//      and appends the indices of the others, in ascending category order, to the cell's survivor list; a short
//      gather step then copies each survivor's four corner values (A, C, B, D) next to it.  Exactness: replace j by j' with L_j'[c] = K_c - m (>= L_j[c]); by monotonicity fl(v_j) <= fl(v_j').
//      The exact interpolants of k* and j' differ by m * (sum of weights) >= m (1 - 2^-22), while the five roundings of
//      one interpolant (two products, three fmas, each relative 2^-24 on terms whose weighted magnitudes sum to
//      <= 4 max|K|) move it by <= 2^-22 max|K|, the pair by <= 2^-21 max|K| (1 + 2^-20) < m (1 - 2^-22).  So
//      fl(v_k*) > fl(v_j) at every pixel of the cell: j is never a maximum, first or otherwise.  Which category serves
//      as dominator only affects how much is pruned, never the result; k* itself has lead 0 and always survives.
//      A cell with a NaN or an infinity among its taps (detected by summing the differences) is left to the NaN-aware
//      brute-force path below, so torch.argmax's NaN ordering is reproduced.
//   E  evaluate, lane = two horizontally adjacent pixels of an 8x8 tile of the cell: walk the survivor list (broadcast
//      LDS.128 per survivor, packed fp32x2 interpolation) with a strict compare = first maximum in ascending category
//      order; store the labels as 32-bit pairs; count (ground truth, label) pairs in the CTA's shared-memory Q x Q
//      histogram with match.any-aggregated atomics.  Cells with a single survivor are labelled without evaluation.
//
// Work distribution: with a workspace, runs are handed out one at a time through a global counter (a run's cost follows
// its survivors, and SMs differ); each warp requests its run two iterations ahead, and the last CTA re-arms the counter.
// Without a workspace, cell rows go round-robin to CTAs and a shared-memory counter hands a CTA's runs to its warps.
// More than 128 categories (NI = 0): the taps are read from global memory instead (a run's taps would not fit).
#include "cells.cuh"

namespace zutis {

namespace {

constexpr int kRunCells = 4;       // cells per run = 32 lanes / 8 slices
constexpr int kSlices = 8;
constexpr int kBoxPixels = kRunCells + 1;
constexpr int kCellWarpsMax = 20;   // 640 threads: 96 registers per thread, no spills (80 registers spill; the spills go to L2 because shared memory leaves no L1)

}  // namespace

// Cell rows of an image are visited from the borders inwards (0, h-1, 1, h-2, ...): the clamped border rows are taller than
// 8 pixels and take the slower general path, and work handed out last should be short (the kernel ends when the last
// run ends).
__device__ __forceinline__ int border_first(int k, int h) { return (k & 1) ? h - 1 - (k >> 1) : (k >> 1); }

struct CellParams {
    const float* logits;
    long sb;
    int sy, sx;
    int B, Q, h, w, H, W;
    float scale_y, scale_x;
    const void* gt;
    long gt_sb;
    int16_t* labels;
    int* hist;
    int n, hist_in_smem;
    int cap;                 // survivor slots per cell
    int runs_per_row;
    unsigned n_rows;         // B * h cell rows
    FastDiv div_runs_per_row, div_h;
    unsigned* counter;       // [2] global run counter + finished-CTA count (zero before the launch, zero after); NULL: static rows
    int pair_ok;             // W even, labels 4-byte aligned, ground truth aligned for pair loads
    int tap_pitch_bytes;     // staged taps: bytes between low-res pixels of the box (= 4 * ceil4(Q))
    int tap_bytes;           // bytes of one staged box (2 rows x 5 pixels), rounded up to 128
    int off_ystart, off_xstart, off_ly, off_lx, off_lyp, off_warp, warp_bytes;    // byte offsets in dynamic shared memory
};

// NI: float4 iterations per lane in phase P (ceil(ceil(Q/4) / 8)), taps staged in shared memory by TMA;
// NI = 0: run-time count, taps from global memory (wide Q).
template <typename GT, int NI>
__global__ void __launch_bounds__(kCellWarpsMax * 32, 1) decode_cells_kernel(const __grid_constant__ CUtensorMap tap_map, const CellParams p) {
    constexpr bool STAGED = NI > 0;
    extern __shared__ __align__(128) float smem[];
    __shared__ unsigned s_next;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nn = p.n * p.n;
    char* smem_c = reinterpret_cast<char*>(smem);
    int* s_hist = reinterpret_cast<int*>(smem);                                   // [n*n] when the histogram fits
    // shared-memory tables, as 32-bit shared addresses kept in registers
    uint32_t a_hist = smem_addr_u32(smem_c);
#define a_ystart (a_hist + (uint32_t)p.off_ystart)                              /* [h+1] first output row of each cell row */
#define a_xstart (a_hist + (uint32_t)p.off_xstart)                              /* [w+1] */
#define a_ly (a_hist + (uint32_t)p.off_ly)                                      /* [H] float2 (ly0, ly1) */
#define a_lx (a_hist + (uint32_t)p.off_lx)                                      /* [W] float4 (lx0, lx0, lx1, lx1): packed operands */
#define a_lyp (a_hist + (uint32_t)p.off_lyp)                                    /* [H] float4 (ly0[Y], ly0[Y+1], ly1[Y], ly1[Y+1]) */
    // warp-private: [staged taps: 2 rows x 5 pixels x Qp] [survivors' corner values (A, C, B, D): kRunCells x (cap+1) float4]
    //               [their categories: kRunCells x (cap+1) u16] [mbarrier]
    const int cap = p.cap;
    uint32_t tap_base = a_hist + (uint32_t)p.off_warp + (uint32_t)warp * (uint32_t)p.warp_bytes;
#define val_base (tap_base + (STAGED ? (uint32_t)p.tap_bytes : 0u))
#define id_base (val_base + (uint32_t)(kRunCells * (cap + 1) * 16))
#define bar (tap_base + (uint32_t)p.warp_bytes - 8u)
#define sent_addr (tap_base + (uint32_t)p.warp_bytes - 32u)                     /* one survivor record of -inf: never a strict maximum */
    // only the two roots are pinned; everything else is root + kernel parameter (one IADD with a constant-bank operand)
    ZUTIS_KEEP(a_hist); ZUTIS_KEEP(tap_base);

    {
        int* ystart = reinterpret_cast<int*>(smem_c + p.off_ystart);
        int* xstart = reinterpret_cast<int*>(smem_c + p.off_xstart);
        float2* ly = reinterpret_cast<float2*>(smem_c + p.off_ly);
        float4* lx = reinterpret_cast<float4*>(smem_c + p.off_lx);
        for (int i = threadIdx.x; i < nn && p.hist_in_smem; i += blockDim.x) s_hist[i] = 0;
        for (int c = threadIdx.x; c <= p.h; c += blockDim.x) ystart[c] = first_dst_with_tap_ge(c, p.h, p.H, p.scale_y);
        for (int c = threadIdx.x; c <= p.w; c += blockDim.x) xstart[c] = first_dst_with_tap_ge(c, p.w, p.W, p.scale_x);
        for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) { const AxisTap t = axis_tap(Y, p.h, p.H, p.scale_y); ly[Y] = make_float2(t.l0, t.l1); }
        for (int X = threadIdx.x; X < p.W; X += blockDim.x) { const AxisTap t = axis_tap(X, p.w, p.W, p.scale_x); lx[X] = make_float4(t.l0, t.l0, t.l1, t.l1); }
        float4* lyp = reinterpret_cast<float4*>(smem_c + p.off_lyp);
        for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) {
            const AxisTap t0 = axis_tap(Y, p.h, p.H, p.scale_y), t1 = axis_tap(min(Y + 1, p.H - 1), p.h, p.H, p.scale_y);
            lyp[Y] = make_float4(t0.l0, t1.l0, t0.l1, t1.l1);
        }
        if (lane == 0) sts128(sent_addr, -INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (threadIdx.x == 0) { s_next = 0; if (STAGED) tma_prefetch_descriptor(&tap_map); }
        if (STAGED && lane == 0) { mbarrier_init(bar, 1); fence_mbarrier_init(); }
    }
    __syncthreads();
#define gt_base (reinterpret_cast<const GT*>(p.gt))

    const int cell = lane >> 3, slice = lane & 7;          // phase P role
    const int nf4 = (p.Q + 3) >> 2;                        // float4 chunks per low-res pixel
    const int iters = STAGED ? NI : (nf4 + kSlices - 1) / kSlices;
    const unsigned long long MINUS1 = pack2(-1.0f, -1.0f);
    const int r8 = lane >> 2, c2 = (lane & 3) * 2;         // phase E role: row and first column inside an 8x8 tile
#define my_id (id_base + (uint32_t)(cell * (cap + 1)) * 2u)

    // ---- work items: a run number, requested by lane 0 ahead of time, decoded by every lane when it is needed
    // With a global counter the first two runs of every warp are fixed (its global warp number, then that plus the number
    // of warps): 7000 warps asking one address at once would start the kernel with several microseconds of queueing.
    const unsigned n_warps_grid = gridDim.x * (blockDim.x >> 5);
    unsigned static_next = blockIdx.x * (blockDim.x >> 5) + (unsigned)warp;
    // (inline PTX: with atomicAdd the compiler aggregates over the warp and reads the result at once, which would put the
    // atomic's round trip right here instead of two phases later, where the value is first needed)
    auto request = [&]() -> unsigned {
        unsigned g = 0;
        if (p.counter) {
            if (static_next < 2u * n_warps_grid) { g = static_next; static_next += n_warps_grid; }
            else if (lane == 0) {
                asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(g) : "l"(p.counter) : "memory");
                g += 2u * n_warps_grid;
            }
        } else if (lane == 0) g = atomicAdd(&s_next, 1u);
        return g;
    };
    // (global cell row b*h + cy, first cell column) of run g; row = 0xffffffff: no more work
    auto decode = [&](unsigned g_lane0, unsigned& row, int& cxb) {
        const unsigned g = __shfl_sync(kFull, g_lane0, 0);
        const unsigned uk = fast_div(g, p.div_runs_per_row);
        cxb = (int)(g - uk * (unsigned)p.runs_per_row) * kRunCells;
        row = p.counter ? uk : blockIdx.x + uk * gridDim.x;
        if (row >= p.n_rows) row = 0xffffffffu;
    };
    auto issue_taps = [&](unsigned row, int cxb) {
        if (STAGED) {
            const unsigned b = fast_div(row, p.div_h);
            const int cy = border_first((int)(row - b * (unsigned)p.h), p.h);
            if (elect_one_lane()) {
                mbarrier_arrive_expect_tx(bar, (uint32_t)(2 * kBoxPixels) * (uint32_t)p.tap_pitch_bytes);
                tma_load_4d(tap_base, &tap_map, bar, 0, cxb, cy, (int)b);
            }
            __syncwarp();
        }
    };

    unsigned cur_row; int cur_cxb;
    decode(request(), cur_row, cur_cxb);
    if (cur_row != 0xffffffffu) issue_taps(cur_row, cur_cxb);
    unsigned pending = request();                          // the run after `cur`
    uint32_t phase = 0;

    while (cur_row != 0xffffffffu) {
        const int b = (int)fast_div(cur_row, p.div_h);
        const int cy = border_first((int)(cur_row - (unsigned)b * (unsigned)p.h), p.h), cx_begin = cur_cxb;
        const int cy1 = min(cy + 1, p.h - 1);
        const int maxslot = min(kRunCells, p.w - 1 - cx_begin);                // last pixel of the box that exists
        const uint32_t lower = (STAGED && cy1 != cy) ? (uint32_t)(kBoxPixels * p.tap_pitch_bytes) : 0u;

        // ------------------------------------------------------------------ P: prune (lane = cell, slice)
        // tap sources of this lane's cell: shared-memory addresses (staged) or global pointers, slice included
        uint32_t aA = 0, aB = 0, aC = 0, aD = 0;
        const float4 *gA = nullptr, *gB = nullptr, *gC = nullptr, *gD = nullptr;
        if (STAGED) {
            aA = tap_base + (uint32_t)(min(cell, maxslot) * p.tap_pitch_bytes) + (uint32_t)slice * 16u;
            aB = tap_base + (uint32_t)(min(cell + 1, maxslot) * p.tap_pitch_bytes) + (uint32_t)slice * 16u;
            aC = aA + lower; aD = aB + lower;
            mbarrier_wait_parity(bar, phase);
            phase ^= 1u;
        } else {
            const float* row0 = p.logits + (long)b * p.sb + (long)cy * p.sy;
            const float* row1 = p.logits + (long)b * p.sb + (long)cy1 * p.sy;
            const int cxl = min(cx_begin + cell, p.w - 1), cxl1 = min(cxl + 1, p.w - 1);
            gA = reinterpret_cast<const float4*>(row0 + (long)cxl * p.sx) + slice;
            gB = reinterpret_cast<const float4*>(row0 + (long)cxl1 * p.sx) + slice;
            gC = reinterpret_cast<const float4*>(row1 + (long)cxl * p.sx) + slice;
            gD = reinterpret_cast<const float4*>(row1 + (long)cxl1 * p.sx) + slice;
        }

        // pass 1: dominator k* = argmax_q min_corner L_q, tracked per chunk of 4 categories
        float best = -INFINITY;
        int bcode = slice;
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int c4 = it * kSlices + slice;
            if ((STAGED && it < NI - 1) || c4 < nf4) {
                float4 a, bq, c_, d;
                if (STAGED) { a = lds128(aA + it * 128); bq = lds128(aB + it * 128); c_ = lds128(aC + it * 128); d = lds128(aD + it * 128); }
                else { a = __ldg(gA + it * kSlices); bq = __ldg(gB + it * kSlices); c_ = __ldg(gC + it * kSlices); d = __ldg(gD + it * kSlices); }
                float m0 = fminf(min3(a.x, bq.x, c_.x), d.x), m1 = fminf(min3(a.y, bq.y, c_.y), d.y);
                float m2 = fminf(min3(a.z, bq.z, c_.z), d.z), m3 = fminf(min3(a.w, bq.w, c_.w), d.w);
                if (!STAGED || it == NI - 1) {                       // padding categories of the last chunk
                    const int q0 = c4 * 4;
                    if (q0 + 1 >= p.Q) m1 = -INFINITY;
                    if (q0 + 2 >= p.Q) m2 = -INFINITY;
                    if (q0 + 3 >= p.Q) m3 = -INFINITY;
                }
                const float gm = fmaxf(max3(m0, m1, m2), m3);
                if (gm > best) { best = gm; bcode = c4; }
            }
        }
#pragma unroll
        for (int o = 1; o < kSlices; o <<= 1) {
            const float ob = __shfl_xor_sync(kFull, best, o);
            const int oc = __shfl_xor_sync(kFull, bcode, o);
            if (ob > best || (ob == best && oc < bcode)) { best = ob; bcode = oc; }
        }
        // the category inside the winning chunk, and its four corner values K
        float Ka, Kb, Kc, Kd;
        {
            float4 a, bq, c_, d;
            if (STAGED) {
                const uint32_t o = (uint32_t)(bcode - slice) * 16u;
                a = lds128(aA + o); bq = lds128(aB + o); c_ = lds128(aC + o); d = lds128(aD + o);
            } else {
                a = __ldg(gA + (bcode - slice)); bq = __ldg(gB + (bcode - slice)); c_ = __ldg(gC + (bcode - slice)); d = __ldg(gD + (bcode - slice));
            }
            const float m0 = fminf(min3(a.x, bq.x, c_.x), d.x), m1 = fminf(min3(a.y, bq.y, c_.y), d.y);
            const float m2 = fminf(min3(a.z, bq.z, c_.z), d.z);
            const int j = m0 == best ? 0 : (m1 == best ? 1 : (m2 == best ? 2 : 3));
            Ka = j == 0 ? a.x : (j == 1 ? a.y : (j == 2 ? a.z : a.w));
            Kb = j == 0 ? bq.x : (j == 1 ? bq.y : (j == 2 ? bq.z : bq.w));
            Kc = j == 0 ? c_.x : (j == 1 ? c_.y : (j == 2 ? c_.z : c_.w));
            Kd = j == 0 ? d.x : (j == 1 ? d.y : (j == 2 ? d.z : d.w));
        }
        // 2^-20 * max|K|, never 0: k* must not dominate itself (its differences are exactly 0)
        float margin = fmaxf(fmaxf(fmaxf(fabsf(Ka), fabsf(Kb)), fmaxf(fabsf(Kc), fabsf(Kd))) * 9.5367431640625e-07f, 1e-37f);
        unsigned long long KA = pack2(Ka, Ka), KB = pack2(Kb, Kb), KC = pack2(Kc, Kc), KD = pack2(Kd, Kd);

        // pass 2: indices of the survivors, ascending, into the cell's list.  Run a second time, with a different
        // dominator, for cells whose list overflowed (see below).
        int n;                                              // survivors of this lane's cell (same in its 8 lanes)
#pragma unroll 1
        for (int attempt = 0;; ++attempt) {
        unsigned long long acc = pack2(0.f, 0.f);          // sum of all differences: non-finite iff a tap is NaN / inf (or overflow)
        n = 0;
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int c4 = it * kSlices + slice;
            const bool have = (STAGED && it < NI - 1) || c4 < nf4;
            const int q0 = c4 * 4;
            int k0 = 0, k1 = 0, k2 = 0, k3 = 0;
            if (have) {
                float4 a, bq, c_, d;
                if (STAGED) { a = lds128(aA + it * 128); bq = lds128(aB + it * 128); c_ = lds128(aC + it * 128); d = lds128(aD + it * 128); }
                else { a = __ldg(gA + it * kSlices); bq = __ldg(gB + it * kSlices); c_ = __ldg(gC + it * kSlices); d = __ldg(gD + it * kSlices); }
                // leads of k* over categories q0..q0+3 at the four corners (packed pairs along the category index)
                const unsigned long long dA0 = fma2(f4_lo(a), MINUS1, KA), dA1 = fma2(f4_hi(a), MINUS1, KA);
                const unsigned long long dB0 = fma2(f4_lo(bq), MINUS1, KB), dB1 = fma2(f4_hi(bq), MINUS1, KB);
                const unsigned long long dC0 = fma2(f4_lo(c_), MINUS1, KC), dC1 = fma2(f4_hi(c_), MINUS1, KC);
                const unsigned long long dD0 = fma2(f4_lo(d), MINUS1, KD), dD1 = fma2(f4_hi(d), MINUS1, KD);
                float a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2_, c3, d0, d1, d2, d3;
                unpack2(dA0, a0, a1); unpack2(dA1, a2, a3); unpack2(dB0, b0, b1); unpack2(dB1, b2, b3);
                unpack2(dC0, c0, c1); unpack2(dC1, c2_, c3); unpack2(dD0, d0, d1); unpack2(dD1, d2, d3);
                const float l0 = fminf(min3(a0, b0, c0), d0), l1 = fminf(min3(a1, b1, c1), d1);
                const float l2 = fminf(min3(a2, b2, c2_), d2), l3 = fminf(min3(a3, b3, c3), d3);
                const unsigned long long s0 = add2(add2(dA0, dB0), add2(dC0, dD0));
                const unsigned long long s1 = add2(add2(dA1, dB1), add2(dC1, dD1));
                k0 = !(l0 >= margin); k1 = !(l1 >= margin); k2 = !(l2 >= margin); k3 = !(l3 >= margin);
                if ((STAGED && it < NI - 1) || q0 + 3 < p.Q) acc = add2(acc, add2(s0, s1));
                else {
                    // padding categories (q >= Q inside the last chunk) hold arbitrary bits: they stay out of the sum and the list
                    float e0, e1, e2, e3;
                    unpack2(s0, e0, e1); unpack2(s1, e2, e3);
                    acc = add2(acc, pack2(e0, q0 + 1 < p.Q ? e1 : 0.f));
                    acc = add2(acc, pack2(q0 + 2 < p.Q ? e2 : 0.f, 0.f));
                    k1 = k1 && q0 + 1 < p.Q; k2 = k2 && q0 + 2 < p.Q; k3 = 0;
                }
            }
            const int mine = k0 + k1 + k2 + k3;
            // exclusive prefix over the 8 lanes of the cell
            int incl = mine;
#pragma unroll
            for (int o = 1; o < kSlices; o <<= 1) {
                const int up = __shfl_up_sync(kFull, incl, o, kSlices);
                if (slice >= o) incl += up;
            }
            int pos = n + incl - mine;
            n += __shfl_sync(kFull, incl, kSlices - 1, kSlices);
            // entries beyond the capacity are not stored; such a cell is evaluated by brute force (n > cap)
            sts_u16_if(k0 && pos < cap, my_id + (uint32_t)pos * 2u, q0); pos += k0;
            sts_u16_if(k1 && pos < cap, my_id + (uint32_t)pos * 2u, q0 + 1); pos += k1;
            sts_u16_if(k2 && pos < cap, my_id + (uint32_t)pos * 2u, q0 + 2); pos += k2;
            sts_u16_if(k3 && pos < cap, my_id + (uint32_t)pos * 2u, q0 + 3);
        }
        // non-finite anywhere in the cell: brute force with torch's NaN ordering
        bool finite;
        {
            float sa, sb_;
            unpack2(acc, sa, sb_);
            float tot = __fadd_rn(sa, sb_);
#pragma unroll
            for (int o = 1; o < kSlices; o <<= 1) tot = __fadd_rn(tot, __shfl_xor_sync(kFull, tot, o));
            finite = fabsf(tot) <= 3.402823466e38f;
            if (!finite) n = cap + 1;
        }
        // A list that overflows means k* does not dominate: typically a cell on the border of two regions, where one
        // category leads on the left corners and another on the right ones, and everything in between survives both.
        // Such a cell gets a second attempt with a VIRTUAL dominator, the average of the four corner champions
        //     V_c = (L_kA[c] + L_kB[c] + L_kC[c] + L_kD[c]) / 4,   k_X = argmax_q L_q[X],
        // whose interpolant is the average of four real interpolants and therefore never above their maximum: a
        // category that stays m' = 2^-19 * max|the 16 taps| below V at all four corners stays below that maximum at
        // every pixel.  (m' covers the roundings of the two interpolants, 2^-21 max|tap| as above, plus the
        // 0.75 * 2^-22 max|tap| of computing V in fp32.)  Each corner champion is within reach of V at its own
        // corner, so the list is never empty.
        const bool again = attempt == 0 && finite && n > cap;
        if (!__any_sync(kFull, again)) break;
        {
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int ix[4] = {0, 0, 0, 0};
            for (int it = 0; it < iters; ++it) {
                const int c4 = it * kSlices + slice;
                if (c4 < nf4) {
                    float4 t[4];
                    if (STAGED) { t[0] = lds128(aA + it * 128); t[1] = lds128(aB + it * 128); t[2] = lds128(aC + it * 128); t[3] = lds128(aD + it * 128); }
                    else { t[0] = __ldg(gA + it * kSlices); t[1] = __ldg(gB + it * kSlices); t[2] = __ldg(gC + it * kSlices); t[3] = __ldg(gD + it * kSlices); }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float e[4] = {t[c].x, t[c].y, t[c].z, t[c].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (c4 * 4 + k < p.Q && e[k] > mx[c]) { mx[c] = e[k]; ix[c] = c4 * 4 + k; }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int o = 1; o < kSlices; o <<= 1) {
                    const float om = __shfl_xor_sync(kFull, mx[c], o);
                    const int oi = __shfl_xor_sync(kFull, ix[c], o);
                    if (om > mx[c] || (om == mx[c] && oi < ix[c])) { mx[c] = om; ix[c] = oi; }
                }
            const uint32_t corner[4] = {aA - (uint32_t)slice * 16u, aB - (uint32_t)slice * 16u, aC - (uint32_t)slice * 16u, aD - (uint32_t)slice * 16u};
            const float* gcorner[4] = {reinterpret_cast<const float*>(gA - slice), reinterpret_cast<const float*>(gB - slice),
                                       reinterpret_cast<const float*>(gC - slice), reinterpret_cast<const float*>(gD - slice)};
            float V[4], big = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float t0, t1, t2, t3;
                if (STAGED) {
                    t0 = lds32(corner[c] + (uint32_t)ix[0] * 4u); t1 = lds32(corner[c] + (uint32_t)ix[1] * 4u);
                    t2 = lds32(corner[c] + (uint32_t)ix[2] * 4u); t3 = lds32(corner[c] + (uint32_t)ix[3] * 4u);
                } else {
                    t0 = __ldg(gcorner[c] + ix[0]); t1 = __ldg(gcorner[c] + ix[1]); t2 = __ldg(gcorner[c] + ix[2]); t3 = __ldg(gcorner[c] + ix[3]);
                }
                V[c] = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(t0, t1), __fadd_rn(t2, t3)));
                big = fmaxf(big, fmaxf(fmaxf(fabsf(t0), fabsf(t1)), fmaxf(fabsf(t2), fabsf(t3))));
            }
            if (again) {
                KA = pack2(V[0], V[0]); KB = pack2(V[1], V[1]); KC = pack2(V[2], V[2]); KD = pack2(V[3], V[3]);
                margin = fmaxf(big * 1.9073486328125e-06f, 1e-37f);
            }
        }
        __syncwarp();
        }
        __syncwarp();
        // ---- gather, lane = list entry of the run: survivor j of cell c -> (A, C, B, D) next to its index
        const int cells_here = min(kRunCells, p.w - cx_begin);
        const int n0 = __shfl_sync(kFull, n, 0), n1 = __shfl_sync(kFull, n, 8), n2 = __shfl_sync(kFull, n, 16), n3 = __shfl_sync(kFull, n, 24);
        {
            const int e0 = (n0 >= 2 && n0 <= cap) ? n0 : 0;
            const int e1 = (n1 >= 2 && n1 <= cap && cells_here > 1) ? n1 : 0;
            const int e2 = (n2 >= 2 && n2 <= cap && cells_here > 2) ? n2 : 0;
            const int e3 = (n3 >= 2 && n3 <= cap && cells_here > 3) ? n3 : 0;
            const int f0 = e0, f01 = f0 + e1, f012 = f01 + e2, total = f012 + e3;
            for (int e = lane; e < total; e += 32) {
                const int c = (e >= f0) + (e >= f01) + (e >= f012);
                const int j = e - (c == 0 ? 0 : (c == 1 ? f0 : (c == 2 ? f01 : f012)));
                const int cnt = c == 0 ? e0 : (c == 1 ? e1 : (c == 2 ? e2 : e3));
                const uint32_t slot = (uint32_t)(c * (cap + 1) + j);
                float A = -INFINITY, Bv = -INFINITY, Cv = -INFINITY, Dv = -INFINITY;
                if (j < cnt) {
                    const unsigned q = lds_u16(id_base + slot * 2u);
                    if (STAGED) {
                        const uint32_t tA = tap_base + (uint32_t)(min(c, maxslot) * p.tap_pitch_bytes) + q * 4u;
                        const uint32_t tB = tap_base + (uint32_t)(min(c + 1, maxslot) * p.tap_pitch_bytes) + q * 4u;
                        A = lds32(tA); Bv = lds32(tB); Cv = lds32(tA + lower); Dv = lds32(tB + lower);
                    } else {
                        const float* row0 = p.logits + (long)b * p.sb + (long)cy * p.sy;
                        const float* row1 = p.logits + (long)b * p.sb + (long)cy1 * p.sy;
                        const int cx = cx_begin + c, cx1 = min(cx + 1, p.w - 1);
                        A = __ldg(row0 + (long)cx * p.sx + q); Bv = __ldg(row0 + (long)cx1 * p.sx + q);
                        Cv = __ldg(row1 + (long)cx * p.sx + q); Dv = __ldg(row1 + (long)cx1 * p.sx + q);
                    }
                }
                sts128(val_base + slot * 16u, A, Cv, Bv, Dv);
            }
        }
        __syncwarp();
        // the taps are consumed: start the next run's load while this one is evaluated, and request the run after it
        unsigned nxt_row; int nxt_cxb;
        decode(pending, nxt_row, nxt_cxb);
        if (nxt_row != 0xffffffffu) {
            if (STAGED) fence_proxy_async_smem();
            issue_taps(nxt_row, nxt_cxb);
            pending = request();
            if (p.hist) {
                // pull the next run's ground truth towards L2: lane = (row of 8, 128-byte segment of 4)
                const unsigned nb = fast_div(nxt_row, p.div_h);
                const int ncy = border_first((int)(nxt_row - nb * (unsigned)p.h), p.h);
                const int nys = (int)lds_u32(a_ystart + (uint32_t)ncy * 4u);
                const int nxs = (int)lds_u32(a_xstart + (uint32_t)nxt_cxb * 4u);
                const int nxe = (int)lds_u32(a_xstart + (uint32_t)min(nxt_cxb + kRunCells, p.w) * 4u);
                const int seg = lane >> 3;
                if (seg * 128 < (nxe - nxs) * (int)sizeof(GT) && nys + (lane & 7) < p.H)
                    prefetch_l2(reinterpret_cast<const char*>(gt_base + (size_t)nb * p.gt_sb + (size_t)(nys + (lane & 7)) * p.W + nxs) + seg * 128);
            }
        }

        // ------------------------------------------------------------------ E: evaluate
        const int ys = (int)lds_u32(a_ystart + (uint32_t)cy * 4u), ye = (int)lds_u32(a_ystart + (uint32_t)cy * 4u + 4u);
        const unsigned img_px = (unsigned)b * (unsigned)(p.H * p.W);
        const GT* gt_img = gt_base + (size_t)b * p.gt_sb;
        // ---- the bottom-right 8x8 block of every cell of the run at once (the whole cell at 8x up-sampling; the clamped
        // border cells are taller / wider and leave an L-shaped rest to the per-cell path below): lane = (cell, column),
        // 8 rows per lane.  The four lists are walked together (a lane past the end of its list reads the -inf record),
        // so the per-run work is shared by the cells.
        unsigned fast_mask;
        {
            const int xs_m = (int)lds_u32(a_xstart + (uint32_t)min(cx_begin + cell, p.w) * 4u);
            const int xe_m = (int)lds_u32(a_xstart + (uint32_t)min(cx_begin + cell + 1, p.w) * 4u);
            const bool fast_m = (ye - ys >= 8) && cell < cells_here && (xe_m - xs_m >= 8) && n <= cap;
            fast_mask = __ballot_sync(kFull, fast_m);
            if (fast_mask) {
                const int yf = ye - 8;                                                             // first row of the block
                const int X = min(max(xe_m - 8, 0) + slice, p.W - 1);
                const unsigned px0 = (unsigned)yf * (unsigned)p.W + (unsigned)X;                 // row yf of this lane's column
                GT g[8];
                if (p.hist) {
                    const GT* gp = gt_img + px0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) g[r] = fast_m ? gp[(size_t)r * p.W] : (GT)-1;
                }
                const int nloop = (fast_m && n >= 2) ? n : 0;
                const int maxn = __reduce_max_sync(kFull, nloop);
                const float4 lx = lds128(a_lx + (uint32_t)X * 16u);
                const unsigned long long LX0 = pack2(lx.x, lx.y), LX1 = pack2(lx.z, lx.w);
                unsigned long long LY0[4], LY1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = lds128(a_lyp + (uint32_t)(yf + 2 * q) * 16u);
                    LY0[q] = pack2(t.x, t.y); LY1[q] = pack2(t.z, t.w);
                }
                float best[8];
                int win[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { best[r] = -INFINITY; win[r] = 0; }
                const uint32_t val_m = val_base + (uint32_t)(cell * (cap + 1)) * 16u;
                float4 v = lds128(nloop > 0 ? val_m : sent_addr);
#pragma unroll 1
                for (int j = 0; j < maxn; ++j) {
                    // the next record is requested before this one is used
                    const float4 nv = lds128(j + 1 < nloop ? val_m + (uint32_t)(j + 1) * 16u : sent_addr);
                    float t, u;
                    unpack2(fma2(LX0, pack2(v.x, v.y), mul2(LX1, pack2(v.z, v.w))), t, u);      // t = fma(lx0,A,lx1*B), u = fma(lx0,C,lx1*D)
                    const unsigned long long tt = pack2(t, t), uu = pack2(u, u);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v0, v1;
                        unpack2(fma2(LY0[q], tt, mul2(LY1[q], uu)), v0, v1);                    // rows 2q, 2q+1: fma(ly0, t, ly1*u)
                        if (v0 > best[2 * q]) { best[2 * q] = v0; win[2 * q] = j; }
                        if (v1 > best[2 * q + 1]) { best[2 * q + 1] = v1; win[2 * q + 1] = j; }
                    }
                    v = nv;
                }
                const uint32_t ids_m = id_base + (uint32_t)(cell * (cap + 1)) * 2u;
                int lab[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) lab[r] = (int)lds_u16(ids_m + (uint32_t)win[r] * 2u);
                if (p.labels && fast_m) {
                    int16_t* out = p.labels + (img_px + px0);
#pragma unroll
                    for (int r = 0; r < 8; ++r) out[(size_t)r * p.W] = (int16_t)lab[r];
                }
                if (p.hist) {
#pragma unroll
                    for (int r = 0; r < 8; r += 2) {
                        const int key0 = (fast_m && label_in_range<GT>(g[r], p.n)) ? (int)g[r] * p.n + lab[r] : -1;
                        const int key1 = (fast_m && label_in_range<GT>(g[r + 1], p.n)) ? (int)g[r + 1] * p.n + lab[r + 1] : -1;
                        // lanes with the same (key0, key1) elect one leader
                        const unsigned peers = __match_any_sync(kFull, ((unsigned long long)(unsigned)key0 << 32) | (unsigned)key1);
                        if (lane == __ffs(peers) - 1) {
                            const int c = __popc(peers);
                            if (p.hist_in_smem) {
                                if (key0 == key1) { if (key0 >= 0) red_shared_add(a_hist + (uint32_t)key0 * 4u, 2 * c); }
                                else { if (key0 >= 0) red_shared_add(a_hist + (uint32_t)key0 * 4u, c); if (key1 >= 0) red_shared_add(a_hist + (uint32_t)key1 * 4u, c); }
                            } else {
                                if (key0 == key1) { if (key0 >= 0) atomicAdd(p.hist + key0, 2 * c); }
                                else { if (key0 >= 0) atomicAdd(p.hist + key0, c); if (key1 >= 0) atomicAdd(p.hist + key1, c); }
                            }
                        }
                    }
                }
            }
        }
        // ---- the other cells of the run, one at a time: lane = pixel pair of an 8x8 tile of the cell
#pragma unroll 1
        for (int ci = 0; ci < cells_here; ++ci) {
            const int cx = cx_begin + ci;
            const int nc = __shfl_sync(kFull, n, ci * kSlices);
            const int xs = (int)lds_u32(a_xstart + (uint32_t)cx * 4u), xe = (int)lds_u32(a_xstart + (uint32_t)cx * 4u + 4u);
            // the block [yb, ye) x [xb, xe) is done already (empty if this cell had none); an 8x8 cell is finished
            const bool had_block = (fast_mask & (1u << (ci * kSlices))) != 0;
            if (had_block && ye - ys == 8 && xe - xs == 8) continue;
            const int yb = had_block ? ye - 8 : ye, xb = had_block ? xe - 8 : xe;
            const uint32_t val = val_base + (uint32_t)(ci * (cap + 1)) * 16u;
            const uint32_t ids = id_base + (uint32_t)(ci * (cap + 1)) * 2u;
            const bool vec = p.pair_ok && ((xs & 1) == 0);
            // ---- general case: any cell shape, odd alignment, list overflow, non-finite taps
            for (int ty = ys; ty < ye; ty += 8) {
                const int Y = ty + r8;
                const bool oky = Y < ye;
                const float2 ly = lds_f2(a_ly + (uint32_t)min(Y, p.H - 1) * 8u);
                for (int tx = xs; tx < xe; tx += 8) {
                    if (ty >= yb && tx >= xb) continue;                                      // tile inside the finished block
                    const int X0 = tx + c2;
                    const bool in_rows = Y >= yb;
                    const bool ok0 = oky && X0 < xe && !(in_rows && X0 >= xb), ok1 = oky && X0 + 1 < xe && !(in_rows && X0 + 1 >= xb);
                    // on the vector path a pair is stored / loaded only as a whole; the odd last column of a cell goes scalar
                    const bool vec_here = vec && (ok0 == ok1);
                    const unsigned px = (unsigned)Y * (unsigned)p.W + (unsigned)X0;
                    GT g0 = (GT)0, g1 = (GT)0;
                    if (p.hist) {
                        const GT* gp = gt_img + px;
                        if (vec_here) {
                            if (ok0) { const typename Pair<GT>::type v = *reinterpret_cast<const typename Pair<GT>::type*>(gp); g0 = (GT)v.x; g1 = (GT)v.y; }
                        } else {
                            if (ok0) g0 = gp[0];
                            if (ok1) g1 = gp[1];
                        }
                    }
                    int i0 = 0, i1 = 0;
                    const float4 la = lds128(a_lx + (uint32_t)min(X0, p.W - 1) * 16u), lb = lds128(a_lx + (uint32_t)min(X0 + 1, p.W - 1) * 16u);
                    if (nc == 1) {
                        i0 = i1 = (int)lds_u16(ids);
                    } else if (nc <= cap) {
                        const unsigned long long LA0 = pack2(la.x, la.y), LA1 = pack2(la.z, la.w);
                        const unsigned long long LB0 = pack2(lb.x, lb.y), LB1 = pack2(lb.z, lb.w);
                        float best0 = -INFINITY, best1 = -INFINITY;
                        const uint32_t last = val + (uint32_t)nc * 16u;
                        uint32_t w0 = val, w1 = val;
                        for (uint32_t at = val; at < last; at += 16u) {
                            const float4 v = lds128(at);
                            const unsigned long long ac = pack2(v.x, v.y), bd = pack2(v.z, v.w);
                            float t0, u0, t1, u1;
                            unpack2(fma2(LA0, ac, mul2(LA1, bd)), t0, u0);
                            unpack2(fma2(LB0, ac, mul2(LB1, bd)), t1, u1);
                            const float v0 = __fmaf_rn(ly.x, t0, __fmul_rn(ly.y, u0));
                            const float v1 = __fmaf_rn(ly.x, t1, __fmul_rn(ly.y, u1));
                            if (v0 > best0) { best0 = v0; w0 = at; }
                            if (v1 > best1) { best1 = v1; w1 = at; }
                        }
                        i0 = (int)lds_u16(ids + ((w0 - val) >> 3)); i1 = (int)lds_u16(ids + ((w1 - val) >> 3));
                    } else {
                        // list overflow or a non-finite tap: every category, taps from global memory, torch's NaN ordering
                        const int cx1 = min(cx + 1, p.w - 1);
                        const float* row0 = p.logits + (long)b * p.sb + (long)cy * p.sy;
                        const float* row1 = p.logits + (long)b * p.sb + (long)cy1 * p.sy;
                        const float* pA = row0 + (long)cx * p.sx;
                        const float* pB = row0 + (long)cx1 * p.sx;
                        const float* pC = row1 + (long)cx * p.sx;
                        const float* pD = row1 + (long)cx1 * p.sx;
                        float best0 = -INFINITY, best1 = -INFINITY;
                        for (int q = 0; q < p.Q; ++q) {
                            const float A = __ldg(pA + q), Bv = __ldg(pB + q), Cv = __ldg(pC + q), Dv = __ldg(pD + q);
                            const float v0 = __fmaf_rn(ly.x, lerp_w(la.x, A, la.z, Bv), __fmul_rn(ly.y, lerp_w(la.x, Cv, la.z, Dv)));
                            const float v1 = __fmaf_rn(ly.x, lerp_w(lb.x, A, lb.z, Bv), __fmul_rn(ly.y, lerp_w(lb.x, Cv, lb.z, Dv)));
                            if (q == 0 || better_nan_aware(v0, best0)) { best0 = v0; i0 = q; }
                            if (q == 0 || better_nan_aware(v1, best1)) { best1 = v1; i1 = q; }
                        }
                    }
                    if (p.labels) {
                        int16_t* out = p.labels + (img_px + px);
                        if (vec_here) {
                            if (ok0) *reinterpret_cast<uint32_t*>(out) = (uint32_t)i0 | ((uint32_t)i1 << 16);
                        } else {
                            if (ok0) out[0] = (int16_t)i0;
                            if (ok1) out[1] = (int16_t)i1;
                        }
                    }
                    if (p.hist) {
                        const int key0 = (ok0 && label_in_range<GT>(g0, p.n)) ? (int)g0 * p.n + i0 : -1;
                        const int key1 = (ok1 && label_in_range<GT>(g1, p.n)) ? (int)g1 * p.n + i1 : -1;
                        // lanes with the same (key0, key1) elect one leader
                        const unsigned peers = __match_any_sync(kFull, ((unsigned long long)(unsigned)key0 << 32) | (unsigned)key1);
                        if (lane == __ffs(peers) - 1) {
                            const int c = __popc(peers);
                            if (p.hist_in_smem) {
                                if (key0 == key1) { if (key0 >= 0) atomicAdd(s_hist + key0, 2 * c); }
                                else { if (key0 >= 0) atomicAdd(s_hist + key0, c); if (key1 >= 0) atomicAdd(s_hist + key1, c); }
                            } else {
                                if (key0 == key1) { if (key0 >= 0) atomicAdd(p.hist + key0, 2 * c); }
                                else { if (key0 >= 0) atomicAdd(p.hist + key0, c); if (key1 >= 0) atomicAdd(p.hist + key1, c); }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        cur_row = nxt_row; cur_cxb = nxt_cxb;
    }
    __syncthreads();
    if (p.hist && p.hist_in_smem) flush_shared_hist(s_hist, p.hist, nn);
    // re-arm the global run counter: the last CTA to get here sets both words back to zero
    if (p.counter && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.counter + 1, 1u) == gridDim.x - 1) { p.counter[0] = 0; p.counter[1] = 0; __threadfence(); }
    }
#undef a_ystart
#undef a_xstart
#undef a_ly
#undef a_lx
#undef a_lyp
#undef sent_addr
#undef val_base
#undef id_base
#undef bar
#undef my_id
#undef gt_base
}

namespace {
typedef void (*CellsKernel)(const CUtensorMap, const CellParams);
template <typename GT>
CellsKernel cells_kernel_for_q(int Q) {
    const int ni = ((Q + 3) / 4 + kSlices - 1) / kSlices;
    switch (ni) {
        case 1: return decode_cells_kernel<GT, 1>;
        case 2: return decode_cells_kernel<GT, 2>;
        case 3: return decode_cells_kernel<GT, 3>;
        case 4: return decode_cells_kernel<GT, 4>;
        default: return decode_cells_kernel<GT, 0>;
    }
}
CellsKernel cells_kernel_for(int gt_dtype, int Q) {
    switch (gt_dtype) {
        case ZUTIS_GT_U8: return cells_kernel_for_q<uint8_t>(Q);
        case ZUTIS_GT_I16: return cells_kernel_for_q<int16_t>(Q);
        case ZUTIS_GT_I32: return cells_kernel_for_q<int32_t>(Q);
        default: return cells_kernel_for_q<long long>(Q);
    }
}
}  // namespace

// Host side.  *launched = false with ZUTIS_OK means "shape not taken, use another kernel" (only when not forced).
int launch_decode_cells(const DecodeParams& d, bool forced, int label_dtype, unsigned* counter, int sms, cudaStream_t stream, bool* launched) {
    *launched = false;
    CellParams p;
    p.logits = d.logits; p.sb = d.sb; p.sy = (int)d.sy; p.sx = (int)d.sx;
    p.B = d.B; p.Q = d.Q; p.h = d.h; p.w = d.w; p.H = d.H; p.W = d.W;
    p.scale_y = d.scale_y; p.scale_x = d.scale_x;
    p.gt = d.gt; p.gt_sb = d.gt_sb; p.labels = d.labels; p.hist = d.hist; p.n = d.n;
    const int nn = p.n * p.n;
    p.hist_in_smem = (d.hist != nullptr) && (nn * 4 <= 64 * 1024);
    const int gt_bytes = d.gt ? gt_dtype_bytes(d.gt_dtype) : 1;
    p.pair_ok = (d.W % 2 == 0) && (!d.labels || (reinterpret_cast<uintptr_t>(d.labels) & 3) == 0) &&
                (!d.hist || ((reinterpret_cast<uintptr_t>(d.gt) % (2 * gt_bytes)) == 0 && d.gt_sb % 2 == 0));
    const int Qp = (d.Q + 3) & ~3;
    const bool staged = d.Q <= 128;
    // survivor slots per cell and warps per CTA: narrow Q keeps up to 32 warps resident, wide Q trades warps for longer lists
    int warps = staged ? kCellWarpsMax : 16;
    p.cap = staged ? 40 : 96;
    if (p.cap > Qp) p.cap = Qp;
    p.tap_pitch_bytes = Qp * 4;
    p.tap_bytes = staged ? ((2 * kBoxPixels * p.tap_pitch_bytes + 127) & ~127) : 0;
    p.warp_bytes = (p.tap_bytes + kRunCells * (p.cap + 1) * 18 + 32 + 127) & ~127;       // + the -inf record and the barrier
    p.off_ystart = (p.hist_in_smem ? ((nn + 3) & ~3) : 0) * 4;
    p.off_xstart = p.off_ystart + ((d.h + 1 + 3) & ~3) * 4;
    p.off_ly = p.off_xstart + ((d.w + 1 + 3) & ~3) * 4;
    p.off_lx = p.off_ly + ((d.H + 1) & ~1) * 8;             // 16-byte aligned: float4 per output column
    p.off_lyp = p.off_lx + d.W * 16;
    p.off_warp = (p.off_lyp + d.H * 16 + 127) & ~127;
    const size_t smem_max = 226 * 1024;
    size_t smem = (size_t)p.off_warp + (size_t)warps * p.warp_bytes;
    while (smem > smem_max && warps > 8) { warps -= 2; smem = (size_t)p.off_warp + (size_t)warps * p.warp_bytes; }
    if (smem > smem_max) {
        if (forced) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: cell kernel does not fit this shape (smem=%zu)", smem);
        return ZUTIS_OK;
    }
    p.runs_per_row = (d.w + kRunCells - 1) / kRunCells;
    const long n_rows = (long)d.B * d.h;
    if (n_rows * p.runs_per_row >= 2147483647L / 2) {
        if (forced) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: too many cells for the cell kernel");
        return ZUTIS_OK;
    }
    p.n_rows = (unsigned)n_rows;
    p.div_runs_per_row = make_fast_div((unsigned)p.runs_per_row);
    p.div_h = make_fast_div((unsigned)d.h);
    p.counter = counter;

    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (staged) {
        // logits as a 4-D tensor {category, x, y, image}; one box = all categories of 5 x 2 low-res pixels.  Pixels and
        // rows beyond the image are zero-filled and never read (the lanes clamp their tap indices like the reference).
        TensorMapEncodeTiledFn fn = tensor_map_encode_fn();
        if (!fn) return fail(ZUTIS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        cuuint64_t dims[4] = {(cuuint64_t)Qp, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.sx * 4, (cuuint64_t)d.sy * 4, (cuuint64_t)(d.B > 1 ? d.sb : (long)d.h * d.sy) * 4};
        cuuint32_t box[4] = {(cuuint32_t)Qp, (cuuint32_t)kBoxPixels, 2, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d.logits), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            if (forced) return fail(ZUTIS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return ZUTIS_OK;
        }
    }
    CellsKernel k = cells_kernel_for(label_dtype, d.Q);
    ZUTIS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long blocks = n_rows < sms ? n_rows : sms;
    k<<<(unsigned)blocks, warps * 32, smem, stream>>>(map, p);
    const int st = check_launch("decode_cells_kernel");
    if (st != ZUTIS_OK) return st;
    *launched = true;
    return ZUTIS_OK;
}

}  // namespace zutis
