// Exact candidate pruning for the fused upsample -> argmax -> labels -> histogram decode (sm_100a): the champion pass
// and the warp-per-cell kernel behind ZUTIS_DECODE_AUTO / ZUTIS_DECODE_PRUNED of zutis_decode_score_ws.
// Same arithmetic as decode_score.cu (bit-exact with oracle/zutis_oracle.c); compiled with -fmad=false.
#include "decode.cuh"

namespace zutis {

// ------------------------------------------------------------------------------ pruned decode
// Exact candidate pruning per low-res CELL (the output pixels that share their top-left tap (cy, cx)).
//
// Every interpolant is  v_q = fma(ly0, fma(lx0, A_q, lx1*B_q), ly1 * fma(lx0, C_q, lx1*D_q))  of the cell's four
// corner logits with non-negative weights, and every rounding in that expression is monotone in the taps.  Hence if
// A_k >= A_j, B_k >= B_j, C_k >= C_j and D_k >= D_j then v_k >= v_j at EVERY pixel of the cell, in floating point.
// Category j can therefore never be the first maximum anywhere in the cell when some k dominates it that way and
//   * k < j (a tie still goes to k), or
//   * k > j and the dominance holds with a margin m = 2^-20 * M, M = max|logit of the image|: the exact difference
//     of the two interpolants is then >= m * (1 - 2^-22) (the weights sum to >= 1 - 2^-23, the margin test itself
//     rounds once), while the roundings of one interpolant (two products and three fmas, each relative 2^-24 on terms
//     whose weighted magnitudes sum to <= 4 M) move it by <= 2^-22 M, the pair by <= 2^-21 M < m: v_k > v_j strictly.
// Only the four corner champions (first maxima of the corner pixels, found once per low-res pixel by champion_kernel)
// are tried as dominators.  On model-like logits 6 of 81 categories survive per cell (22 of 920); the survivors'
// corner values are compacted into shared memory and the warp (lane = pixel of an 8x8 tile, 2 pixels per lane) walks
// only those, in ascending category order with a strict compare = torch.argmax's first maximum.
// A cell whose four corners share one champion that leads every category in front of it by the margin is labelled
// without reading a tap (the categories behind the champion can at best tie): most cells of a real segmentation map.
// A warp walks a run of kCellRun horizontally adjacent cells (runs are handed out through an atomic counter); the next
// cell's champions are requested one cell ahead, the corner taps (NQ = ceil(Q/32) values per lane and corner; NQ = 0:
// wide Q, read inside the loop) are re-read per cell -- the left pair hits L1 -- because carrying them in registers
// cost more than the loads.  The ground truth of a run is pulled towards L2 at its start and the first tile's labels
// are requested before the pruning work, so that their latency is hidden.
// Images with a non-finite logit (NaN ordering) or without spatial coherence (pruning would not pay) are left to the
// tiled kernel; both kernels derive the same image split from champion_kernel's per-image counters.
constexpr int kPrunedWarps = 24;     // one CTA per SM: one shared-memory histogram to flush per SM instead of three
constexpr int kCellRun = 4;

// Per low-res pixel: first-max category and its lead over the categories in front of it; per image: the number of
// horizontally adjacent pixels that share their champion, a non-finite flag and max |logit|.  One block per (image, low-res row); 8 lanes per pixel read the pixel's
// categories as float4 (category index contiguous, 16-byte aligned pixels), 4 pixels per warp at a time.
__global__ void __launch_bounds__(256) champion_kernel(const float* __restrict__ logits, long sb, long sy, long sx, int B, int Q,
                                                       int h, int w, int* __restrict__ champ, float* __restrict__ lead,
                                                       int* __restrict__ stats) {
    extern __shared__ int s_row[];                            // [w]
    __shared__ int s_agree[8];
    const int b = blockIdx.x / h, y = blockIdx.x % h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, grp = lane >> 3;
    const float* row = logits + (long)b * sb + (long)y * sy;
    const int chunks = (Q + 3) >> 2;
    bool bad = false;
    float row_amax = 0.f;
    for (int x0 = warp * 4; x0 < w; x0 += 32) {
        const int x = x0 + grp;
        float best = -INFINITY, amax = 0.f;
        int idx = 0x7fffffff;
        if (x < w) {
            const float4* v = reinterpret_cast<const float4*>(row + (long)x * sx);
            for (int c = sub; c < chunks; c += 8) {
                const float4 f = __ldg(v + c);
                const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = c * 4 + j;
                    if (q < Q) {
                        const float af = fabsf(e[j]);
                        bad = bad || !(af <= 3.402823466e38f);
                        amax = fmaxf(amax, af);
                        if (e[j] > best || idx == 0x7fffffff) { best = e[j]; idx = q; }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
        }
        row_amax = fmaxf(row_amax, amax);
        // second sweep (L1 hits): the best of the categories in front of the champion
        float prev = -INFINITY;
        if (x < w) {
            const float4* v = reinterpret_cast<const float4*>(row + (long)x * sx);
            for (int c = sub; c < chunks; c += 8) {
                const float4 f = __ldg(v + c);
                const float e[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c * 4 + j < idx) prev = fmaxf(prev, e[j]);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) prev = fmaxf(prev, __shfl_xor_sync(0xffffffffu, prev, o));
        if (sub == 0 && x < w) {
            champ[((long)b * h + y) * w + x] = idx;
            lead[((long)b * h + y) * w + x] = __fsub_rn(best, prev);
            s_row[x] = idx;
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    int amax_bits = __float_as_int(row_amax);
    for (int o = 16; o > 0; o >>= 1) amax_bits = max(amax_bits, __shfl_xor_sync(0xffffffffu, amax_bits, o));
    if (lane == 0) {
        if (bad) atomicOr(stats + B + b, 1);
        atomicMax(stats + 2 * B + b, amax_bits);
    }
    __syncthreads();
    int agree = 0;
    for (int x = threadIdx.x; x + 1 < w; x += blockDim.x) agree += (s_row[x] == s_row[x + 1]);
    for (int o = 16; o > 0; o >>= 1) agree += __shfl_xor_sync(0xffffffffu, agree, o);
    if (lane == 0) s_agree[warp] = agree;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += s_agree[k];
        if (t) atomicAdd(stats + b, t);
    }
}

template <typename GT, int NQ>
__global__ void __launch_bounds__(kPrunedWarps * 32, 1) decode_pruned_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nn = p.n * p.n;
    char* smem_c = reinterpret_cast<char*>(smem);
    int* s_hist = reinterpret_cast<int*>(smem);                                   // [n*n] when the histogram fits
    int* s_ystart = reinterpret_cast<int*>(smem_c + p.off_ystart);                // [h+1] first output row of each cell row
    int* s_xstart = reinterpret_cast<int*>(smem_c + p.off_xstart);                // [w+1]
    float2* s_ly = reinterpret_cast<float2*>(smem_c + p.off_ly);                  // [H] (ly0, ly1)
    float2* s_lx = reinterpret_cast<float2*>(smem_c + p.off_lx);                  // [W] (lx0, lx1)
    int* s_img = reinterpret_cast<int*>(smem_c + p.off_img);                      // [B] images of this launch
    // per-warp areas behind the tables: survivors' corner values (A, C, B, D), their categories, the champions' corner values
    char* warp_area = smem_c + p.off_warp;
    float4* s_val = reinterpret_cast<float4*>(warp_area) + warp * p.cap;
    int* s_list = reinterpret_cast<int*>(warp_area + (size_t)kPrunedWarps * p.cap * 16) + warp * p.cap;
    float* s_champ = reinterpret_cast<float*>(warp_area + (size_t)kPrunedWarps * p.cap * 20) + warp * 16;
    __shared__ int s_nimg;

    build_image_list(p, 1, s_img, &s_nimg);
    __syncthreads();
    if (s_nimg == 0) return;

    for (int i = threadIdx.x; i < nn && p.hist_in_smem; i += blockDim.x) s_hist[i] = 0;
    for (int c = threadIdx.x; c <= p.h; c += blockDim.x) s_ystart[c] = first_dst_with_tap_ge(c, p.h, p.H, p.scale_y);
    for (int c = threadIdx.x; c <= p.w; c += blockDim.x) s_xstart[c] = first_dst_with_tap_ge(c, p.w, p.W, p.scale_x);
    for (int Y = threadIdx.x; Y < p.H; Y += blockDim.x) { const AxisTap t = axis_tap(Y, p.h, p.H, p.scale_y); s_ly[Y] = make_float2(t.l0, t.l1); }
    for (int X = threadIdx.x; X < p.W; X += blockDim.x) { const AxisTap t = axis_tap(X, p.w, p.W, p.scale_x); s_lx[X] = make_float2(t.l0, t.l1); }
    __syncthreads();
    int* hist = p.hist ? (p.hist_in_smem ? s_hist : p.hist) : nullptr;
    const GT* gt_base = reinterpret_cast<const GT*>(p.gt);

    const int sx = (int)p.sx, sy = (int)p.sy;
    const unsigned runs_per_row = (unsigned)(p.w + kCellRun - 1) / kCellRun;
    const unsigned runs_per_image = runs_per_row * (unsigned)p.h;
    const unsigned total = (unsigned)s_nimg * runs_per_image;
    // runs are handed out dynamically (their cost follows the number of survivors): one atomic per run of kCellRun cells
    unsigned* work_counter = reinterpret_cast<unsigned*>(p.img_stats + 3 * p.B);
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const unsigned slot = item / runs_per_image;
        const unsigned rr = item - slot * runs_per_image;
        const int cy = (int)(rr / runs_per_row);
        const int cx_begin = (int)(rr - (unsigned)cy * runs_per_row) * kCellRun;
        const int cx_end = min(cx_begin + kCellRun, p.w);
        const int b = s_img[slot];
        const int ys = s_ystart[cy], ye = s_ystart[cy + 1];
        if (ys >= ye) continue;
        const int cy1 = min(cy + 1, p.h - 1);
        const float* row0 = p.logits + (long)b * p.sb + cy * sy;      // taps of the cells' upper corners
        const float* row1 = p.logits + (long)b * p.sb + cy1 * sy;     //                    lower corners
        const int* ch0 = p.champ + ((size_t)b * p.h + cy) * p.w;
        const int* ch1 = p.champ + ((size_t)b * p.h + cy1) * p.w;

        // 2^-20 * max|logit of the image|, never 0: a champion must not dominate itself (its differences are exactly 0)
        const float margin = fmaxf(__int_as_float(p.img_stats[2 * p.B + b]) * 9.5367431640625e-07f, 1e-37f);
        // per-lane bases: lane = (row lane/8 [+4], column lane%8) of an 8x8 pixel tile; category lane (+32, +64, ..) of a tap
        const size_t lane_px = (size_t)(ys + (lane >> 3)) * p.W + (lane & 7);
        int16_t* lbl_lane = p.labels ? p.labels + (size_t)b * p.H * p.W + lane_px : nullptr;
        const GT* gt_lane = gt_base + (size_t)b * p.gt_sb + lane_px;
        const float* row0_lane = row0 + lane;
        const float* row1_lane = row1 + lane;
        const int W4 = 4 * p.W;
        if (hist) {
            // pull the run's ground truth towards L2 now (the whole-cell shortcut has nothing to hide its latency behind):
            // lane = (row, 128-byte segment) of the run's pixel rectangle
            const int xs_run = s_xstart[cx_begin], span = (s_xstart[cx_end] - xs_run) * (int)sizeof(GT);
            const int segs = (span + 127) >> 7, r = lane / max(segs, 1), sgm = lane - r * max(segs, 1);
            if (segs > 0 && ys + r < ye && r < 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(gt_base + (size_t)b * p.gt_sb + (size_t)(ys + r) * p.W + xs_run) + sgm * 128));
        }

        // champions of the current cell's four corners; the next cell's right corners are requested one cell ahead
        int hA = __ldg(ch0 + cx_begin), hC = __ldg(ch1 + cx_begin);
        int hB = __ldg(ch0 + min(cx_begin + 1, p.w - 1)), hD = __ldg(ch1 + min(cx_begin + 1, p.w - 1));

        for (int cx = cx_begin; cx < cx_end; ++cx) {
            const int cx1 = min(cx + 1, p.w - 1);
            const int xs = s_xstart[cx], xe = s_xstart[cx + 1];
            const float* pA = row0 + cx * sx;
            const float* pB = row0 + cx1 * sx;
            const float* pC = row1 + cx * sx;
            const float* pD = row1 + cx1 * sx;
            const int cx2 = min(cx + 2, p.w - 1);
            const int hBn = __ldg(ch0 + cx2), hDn = __ldg(ch1 + cx2);
            // One champion at all four corners that leads every category in front of it by the margin at each of them wins
            // the whole cell (the categories behind it can at best tie, and ties go to the smaller index): no taps, no
            // survivor list, no evaluation.  Real segmentation maps are mostly such cells.
            if (hA == hB && hA == hC && hA == hD) {
                // the leads live lead_delta elements behind the champions (one constant instead of two more row pointers)
                const float* ld0 = reinterpret_cast<const float*>(ch0) + p.lead_delta;
                const float* ld1 = reinterpret_cast<const float*>(ch1) + p.lead_delta;
                const float l = fminf(fminf(__ldg(ld0 + cx), __ldg(ld0 + cx1)), fminf(__ldg(ld1 + cx), __ldg(ld1 + cx1)));
                if (l >= margin) {
                    for (int ty = ys; ty < ye; ty += 8) {
                        for (int tx = xs; tx < xe; tx += 8) {
                            const bool okx = tx + (lane & 7) < xe, ok0 = okx && ty + (lane >> 3) < ye, ok1 = okx && ty + (lane >> 3) + 4 < ye;
                            const int tile_off = (ty - ys) * p.W + tx;
                            if (lbl_lane) {
                                if (ok0) lbl_lane[tile_off] = (int16_t)hA;
                                if (ok1) lbl_lane[tile_off + W4] = (int16_t)hA;
                            }
                            if (hist) {
                                const int c0 = ok0 ? class_of<GT>(gt_lane[tile_off], p.n) : -1;
                                const int c1 = ok1 ? class_of<GT>(gt_lane[tile_off + W4], p.n) : -1;
                                warp_hist_add(hist, c0 >= 0 ? c0 * p.n + hA : -1);
                                warp_hist_add(hist, c1 >= 0 ? c1 * p.n + hA : -1);
                            }
                        }
                    }
                    hA = hB; hC = hD; hB = hBn; hD = hDn;
                    continue;
                }
            }
            // the four corner taps of this lane's categories (the left pair was the previous cell's right pair: L1 hits)
            float la_[NQ > 0 ? NQ : 1], lc_[NQ > 0 ? NQ : 1], rb_[NQ > 0 ? NQ : 1], rd_[NQ > 0 ? NQ : 1];
            if (NQ > 0) {
#pragma unroll
                for (int it = 0; it < NQ; ++it) {
                    const bool qv = it * 32 + lane < p.Q;
                    la_[it] = qv ? __ldg(row0_lane + cx * sx + it * 32) : 0.f;
                    lc_[it] = qv ? __ldg(row1_lane + cx * sx + it * 32) : 0.f;
                    rb_[it] = qv ? __ldg(row0_lane + cx1 * sx + it * 32) : 0.f;
                    rd_[it] = qv ? __ldg(row1_lane + cx1 * sx + it * 32) : 0.f;
                }
            }
            // ground truth of the first tile, requested now, used after the evaluation
            const int Xf = xs + (lane & 7), Yf0 = ys + (lane >> 3), Yf1 = Yf0 + 4;
            const bool okxf = Xf < xe, okf0 = okxf && Yf0 < ye, okf1 = okxf && Yf1 < ye;
            GT g0 = (GT)0, g1 = (GT)0;
            if (hist) {
                if (okf0) g0 = gt_lane[xs];
                if (okf1) g1 = gt_lane[xs + W4];
            }
            if (xs < xe) {
                const int kk[4] = {hA, hB, hC, hD};
                const bool use[4] = {true, kk[1] != kk[0], kk[2] != kk[0] && kk[2] != kk[1], kk[3] != kk[0] && kk[3] != kk[1] && kk[3] != kk[2]};

                int n = 0;
                // the champions' values at the four corners: lane 4c+r fetches corner r of champion c
                __syncwarp();
                if (lane < 16) {
                    const int c = lane >> 2, r = lane & 3;
                    const int k = c == 0 ? kk[0] : c == 1 ? kk[1] : c == 2 ? kk[2] : kk[3];
                    const float* pr = r == 0 ? pA : r == 1 ? pB : r == 2 ? pC : pD;
                    s_champ[lane] = __ldg(pr + k);
                }
                __syncwarp();
                const unsigned long long MINUS1 = pack2(-1.0f, -1.0f);

                // survivors, ascending category order
                const int iters = NQ > 0 ? NQ : (p.Q + 31) >> 5;
#pragma unroll
                for (int it = 0; it < iters; ++it) {
                    const int q0 = it * 32, q = q0 + lane;
                    const bool valid = q < p.Q;
                    float a, bq, c_, d;
                    if (NQ > 0) { a = la_[it]; bq = rb_[it]; c_ = lc_[it]; d = rd_[it]; }
                    else {
                        a = bq = c_ = d = 0.f;
                        if (valid) { a = __ldg(pA + q); bq = __ldg(pB + q); c_ = __ldg(pC + q); d = __ldg(pD + q); }
                    }
                    bool dom = false;
                    const unsigned long long ac = pack2(a, c_), bd = pack2(bq, d);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (use[c]) {                              // warp-uniform
                            // smallest of the champion's four leads over q: RN(champion - q) is >= 0 exactly when champion >= q.
                            // A champion with a smaller index may tie; one with a larger index (or q itself) must lead by the margin.
                            const float4 t4 = reinterpret_cast<const float4*>(s_champ)[c];    // champion c at (A, B, C, D): broadcast read
                            float d0, d1, d2, d3;
                            unpack2(fma2(ac, MINUS1, pack2(t4.x, t4.z)), d0, d1);
                            unpack2(fma2(bd, MINUS1, pack2(t4.y, t4.w)), d2, d3);
                            const float lead = fminf(fminf(d0, d1), fminf(d2, d3));
                            dom = dom || (lead >= (kk[c] < q ? 0.f : margin));
                        }
                    }
                    const bool keep = valid && !dom;
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const int pos = n + __popc(bal & ((1u << lane) - 1u));
                        if (pos < p.cap) { s_val[pos] = make_float4(a, c_, bq, d); s_list[pos] = q; }
                    }
                    n += __popc(bal);
                }
                __syncwarp();

                // evaluate the survivors on 8x8 pixel tiles of the cell: lane = (row lane/8 and +4, column lane%8)
                for (int ty = ys; ty < ye; ty += 8) {
                    for (int tx = xs; tx < xe; tx += 8) {
                        const int X = tx + (lane & 7), Y0 = ty + (lane >> 3), Y1 = Y0 + 4;
                        const bool okx = X < xe, ok0 = okx && Y0 < ye, ok1 = okx && Y1 < ye;
                        const float2 lx = s_lx[min(X, p.W - 1)], la = s_ly[min(Y0, p.H - 1)], lb = s_ly[min(Y1, p.H - 1)];
                        float best0 = -INFINITY, best1 = -INFINITY;
                        int i0 = 0, i1 = 0;
                        if (n <= p.cap) {
                            const unsigned long long LX0 = pack2(lx.x, lx.x), LX1 = pack2(lx.y, lx.y);
                            // walk the survivor list by shared-memory address (the winner is remembered as an address too)
                            const uint32_t first = (uint32_t)__cvta_generic_to_shared(s_val), last = first + (uint32_t)n * 16u;
                            uint32_t w0 = first, w1 = first;
#pragma unroll 2
                            for (uint32_t at = first; at < last; at += 16u) {
                                float4 v;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(at));
                                float tt, uu;
                                unpack2(fma2(LX0, pack2(v.x, v.y), mul2(LX1, pack2(v.z, v.w))), tt, uu);   // t = fma(lx0,A,lx1*B), u = fma(lx0,C,lx1*D)
                                const float v0 = __fmaf_rn(la.x, tt, __fmul_rn(la.y, uu));
                                const float v1 = __fmaf_rn(lb.x, tt, __fmul_rn(lb.y, uu));
                                if (v0 > best0) { best0 = v0; w0 = at; }
                                if (v1 > best1) { best1 = v1; w1 = at; }
                            }
                            i0 = s_list[(w0 - first) >> 4]; i1 = s_list[(w1 - first) >> 4];
                        } else {
                            // more survivors than slots (incoherent cell of a wide-Q image): every category, taps from global memory
                            for (int q = 0; q < p.Q; ++q) {
                                const float tt = lerp_w(lx.x, __ldg(pA + q), lx.y, __ldg(pB + q));
                                const float uu = lerp_w(lx.x, __ldg(pC + q), lx.y, __ldg(pD + q));
                                const float v0 = __fmaf_rn(la.x, tt, __fmul_rn(la.y, uu));
                                const float v1 = __fmaf_rn(lb.x, tt, __fmul_rn(lb.y, uu));
                                if (v0 > best0) { best0 = v0; i0 = q; }
                                if (v1 > best1) { best1 = v1; i1 = q; }
                            }
                        }
                        const int tile_off = (ty - ys) * p.W + tx;    // relative to the lane's base pixel
                        if (lbl_lane) {
                            if (ok0) lbl_lane[tile_off] = (int16_t)i0;
                            if (ok1) lbl_lane[tile_off + W4] = (int16_t)i1;
                        }
                        if (hist) {
                            if (ty != ys || tx != xs) {           // later tiles of a large cell: the ground truth was not requested ahead
                                g0 = ok0 ? gt_lane[tile_off] : (GT)0;
                                g1 = ok1 ? gt_lane[tile_off + W4] : (GT)0;
                            }
                            const int c0 = ok0 ? class_of<GT>(g0, p.n) : -1, c1 = ok1 ? class_of<GT>(g1, p.n) : -1;
                            warp_hist_add(hist, c0 >= 0 ? c0 * p.n + i0 : -1);
                            warp_hist_add(hist, c1 >= 0 ? c1 * p.n + i1 : -1);
                        }
                    }
                }
            }
            // slide right
            hA = hB; hC = hD; hB = hBn; hD = hDn;
        }
    }
    if (p.hist && p.hist_in_smem) {
        __syncthreads();
        flush_shared_hist(s_hist, p.hist, nn);
    }
}

namespace {
typedef void (*PrunedKernel)(const DecodeParams);
template <typename GT>
PrunedKernel pruned_kernel_for_q(int Q) {
    const int nq = (Q + 31) / 32;
    switch (nq) {
        case 1: return decode_pruned_kernel<GT, 1>;
        case 2: return decode_pruned_kernel<GT, 2>;
        case 3: return decode_pruned_kernel<GT, 3>;
        case 4: return decode_pruned_kernel<GT, 4>;
        default: return decode_pruned_kernel<GT, 0>;
    }
}
PrunedKernel pruned_kernel_for(int gt_dtype, int Q) {
    switch (gt_dtype) {
        case ZUTIS_GT_U8: return pruned_kernel_for_q<uint8_t>(Q);
        case ZUTIS_GT_I16: return pruned_kernel_for_q<int16_t>(Q);
        case ZUTIS_GT_I32: return pruned_kernel_for_q<int32_t>(Q);
        default: return pruned_kernel_for_q<long long>(Q);
    }
}
}  // namespace

int launch_decode_pruned(DecodeParams& p, bool forced, bool champions_ready, int label_dtype, void* workspace, int sms,
                         cudaStream_t stream, bool* launched) {
    *launched = false;
    const int B = p.B, Q = p.Q, h = p.h, w = p.w, H = p.H, W = p.W;
    const int nn = p.n * p.n;
    p.cap = Q <= 128 ? ((Q + 3) & ~3) : 256;
    p.off_ystart = (p.hist_in_smem ? ((nn + 3) & ~3) : 0) * 4;
    p.off_xstart = p.off_ystart + ((h + 1 + 3) & ~3) * 4;
    p.off_ly = p.off_xstart + ((w + 1 + 3) & ~3) * 4;
    p.off_lx = p.off_ly + ((H + 1) & ~1) * 8;
    p.off_img = p.off_lx + ((W + 1) & ~1) * 8;
    p.off_warp = p.off_img + ((B + 3) & ~3) * 4;              // every offset is a multiple of 16 bytes
    const size_t psmem = (size_t)p.off_warp + (size_t)kPrunedWarps * p.cap * 20 + (size_t)kPrunedWarps * 64;
    if (psmem > 200 * 1024) {
        if (forced) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: pruned kernel does not fit this shape (smem=%zu)", psmem);
        return ZUTIS_OK;                                      // AUTO: the tiled kernel alone
    }
    int* champ = decode_ws_champ(workspace);
    float* lead = decode_ws_lead(workspace, B, (long)h * w);
    int* stats = decode_ws_stats(workspace, B, (long)h * w);
    int st = ZUTIS_OK;
    if (!champions_ready) {
        ZUTIS_CUDA(cudaMemsetAsync(stats, 0, decode_ws_counter_bytes(B), stream));
        champion_kernel<<<(unsigned)(B * h), 256, (size_t)w * 4, stream>>>(p.logits, p.sb, p.sy, p.sx, B, Q, h, w, champ, lead, stats);
        st = check_launch("champion_kernel");
        if (st != ZUTIS_OK) return st;
    }
    p.champ = champ; p.lead = lead; p.lead_delta = lead - reinterpret_cast<const float*>(champ); p.img_stats = stats;
    // AUTO: an image is worth pruning when >= 10 % of its horizontally adjacent low-res pixels share their champion
    // (model outputs: ~30 %; i.i.d. noise: 1 %); forced: every finite image goes through the pruned kernel
    p.agree_min = forced ? 0 : (int)(((long)h * (w - 1) + 9) / 10);
    p.select = 2;
    PrunedKernel pk = pruned_kernel_for(label_dtype, Q);
    ZUTIS_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    int per_sm = 1;
    ZUTIS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk, kPrunedWarps * 32, psmem));
    if (per_sm < 1) return fail(ZUTIS_ERR_UNSUPPORTED, "zutis_decode_score: pruned kernel does not fit (smem %zu)", psmem);
    long blocks = ((long)B * h * ((w + kCellRun - 1) / kCellRun) + kPrunedWarps - 1) / kPrunedWarps;
    if (blocks > (long)sms * per_sm) blocks = (long)sms * per_sm;
    pk<<<(unsigned)blocks, kPrunedWarps * 32, psmem, stream>>>(p);
    st = check_launch("decode_pruned_kernel");
    if (st != ZUTIS_OK) return st;
    p.select = 1;
    *launched = true;
    return ZUTIS_OK;
}

}  // namespace zutis
