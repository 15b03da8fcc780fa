/*
 * zutis_b200.h -- C ABI of libzutis_b200.so: the B200 (sm_100a) implementation of ZUTIS's
 * dense mask-decode + scoring hot path (NoelShin/zutis).
 *
 * The reference has no FFI layer of its own: its boundary for this path is the Python call
 * surface of networks/zutis.py (ZUTIS.predict, ZUTIS.get_mask_proposals), utils/running_score.py
 * (RunningScore) and utils/iou.py (compute_iou).  Each entry point below names the reference
 * expression it replaces; zutis_b200/{decode,running_score,iou}.py re-create the reference
 * call surface on top of these functions through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *   - Every function returns a zutis_status (0 = OK).  zutis_last_error_string() describes the
 *     last failure on the calling thread.  No exceptions, no device-side asserts.
 *   - All device pointers belong to the caller (PyTorch allocates them).  Kernels never
 *     allocate or free; nothing is cached between calls except immutable per-device facts.
 *   - Work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL is the
 *     legacy default stream).  Functions whose name ends in _host synchronise internally.
 *   - Strides are in ELEMENTS, not bytes.
 */
#ifndef ZUTIS_B200_H_
#define ZUTIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZUTIS_API __attribute__((visibility("default")))
#else
#define ZUTIS_API
#endif

typedef enum {
    ZUTIS_OK = 0,
    ZUTIS_ERR_BAD_ARG = 1,      /* null pointer, non-positive size, inconsistent shapes */
    ZUTIS_ERR_UNSUPPORTED = 2,  /* legal request this build has no kernel for */
    ZUTIS_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed (see error string) */
    ZUTIS_ERR_WORKSPACE = 4,    /* caller-provided workspace too small */
    ZUTIS_ERR_NO_DEVICE = 5     /* no sm_100 device: there is no CPU fallback */
} zutis_status;

/* ground-truth label storage types accepted by the scoring kernels */
typedef enum { ZUTIS_GT_U8 = 0, ZUTIS_GT_I16 = 1, ZUTIS_GT_I32 = 2, ZUTIS_GT_I64 = 3 } zutis_gt_dtype;

/* decode kernel selection */
typedef enum {
    ZUTIS_DECODE_AUTO = 0,     /* fastest exact kernel the shape allows */
    ZUTIS_DECODE_GENERIC = 1,  /* one thread per output pixel, any scale/stride, NaN-exact */
    ZUTIS_DECODE_TILED = 2,    /* warp-tile kernel: low-res taps staged in shared memory */
    ZUTIS_DECODE_CELLS = 3,    /* exact per-cell candidate pruning: taps by TMA, one dominator per cell, survivor lists */
    /* or-ed into the mode of zutis_decode_score_ws: the first 16 bytes of the workspace are zero (a fresh cudaMemset,
     * or the workspace was last used by zutis_decode_score_ws, which leaves them zero), so no memset is enqueued */
    ZUTIS_DECODE_WORKSPACE_ZEROED = 0x100
} zutis_decode_mode;

/* GEMM flags (bitwise or) */
enum {
    ZUTIS_GEMM_FP32_SIMT = 0,      /* fp32 FFMA kernel (exact fp32 products, fp32 accumulate) */
    ZUTIS_GEMM_TF32X3 = 1,         /* tcgen05 kind::tf32, 3-term error-compensated split (fp32-grade) */
    ZUTIS_GEMM_TF32 = 2,           /* tcgen05 kind::tf32, single pass (reduced precision: meets the 2e-2 logit bar only) */
    ZUTIS_GEMM_PRECISION_MASK = 3,
    ZUTIS_GEMM_SIGMOID = 16,       /* fused sigmoid epilogue (zutis.py:209) */
    ZUTIS_GEMM_A_PREPARED = 32     /* tcgen05 paths: `workspace` still holds the hi/lo split of this same A from an earlier
                                      call with identical M, K, batch and flags (text embeddings are constant per model,
                                      zutis.py:36-38), so the split kernel is skipped */
};

ZUTIS_API const char* zutis_last_error_string(void);
ZUTIS_API int zutis_abi_version(void);

/* ZUTIS_OK iff `device` is an sm_100 part this library has kernels for. */
ZUTIS_API int zutis_device_check(int device);

/* ---------------------------------------------------------------------------------------------
 * (1) Query x patch contraction.
 * Replaces torch.einsum("nc,bchw->bnhw", text, tokens)        networks/zutis.py:361-365, trainer.py:162-166
 *      and torch.einsum("bqc,bhwc->bqhw" | "bdqc,bhwc->bdqhw") networks/zutis.py:184-186, :196-198 (+ sigmoid :209)
 *
 *   out[b][n][p] = act( sum_k A[b][n][k] * Bm[b][p][k] )      n < M (queries/categories), p < N (pixels)
 *
 * A  : M rows of K floats, row stride lda, batch stride strideA (0 => one A shared by the batch,
 *      the semantic path's text embeddings).
 * Bm : N rows (pixels, channel-last) of K floats, row stride ldb, batch stride strideB.
 * C  : element (b,n,p) is written at C[b*strideC + n*stride_cn + p*stride_cp]; (stride_cn=1,
 *      stride_cp>=M) is the pixel-major layout the decode kernels read fastest, (stride_cn=N,
 *      stride_cp=1) is the reference's [B,Q,h,w].
 * workspace: device scratch of at least zutis_gemm_workspace_bytes(...) bytes (may be 0 / NULL).
 * ------------------------------------------------------------------------------------------- */
ZUTIS_API size_t zutis_gemm_workspace_bytes(int M, long N, int K, int batch, int flags);
ZUTIS_API int zutis_gemm_logits(const float* A, long lda, long strideA,
                                const float* Bm, long ldb, long strideB,
                                float* C, long stride_cn, long stride_cp, long strideC,
                                int M, long N, int K, int batch, int flags,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (2)+(3)+(4) Fused bilinear upsample -> per-pixel argmax -> int16 labels -> confusion histogram.
 * Replaces F.interpolate(size, "bilinear") + torch.argmax(dim=1)   networks/zutis.py:366-372, trainer.py:169-173
 *      and RunningScore._fast_hist / update                        utils/running_score.py:10-20
 * Full-resolution logits are never written.
 *
 * logits : element (b,q,y,x) at logits[b*sb + q*sq + y*sy + x*sx], fp32, low resolution h x w.
 * H, W   : output size; H == h && W == w is the reference's size=None / copy branch (plain argmax).
 * gt     : ground truth [B,H,W] of `gt_dtype`, image stride gt_sb elements (row stride W); may be
 *          NULL when hist_partial is NULL.  A pixel counts iff 0 <= gt < n_classes (running_score.py:12).
 * labels : optional int16 [B,H,W] output (first-max rule, NaN counts as maximum: torch.argmax).
 * hist_partial : optional int32 [n_classes*n_classes], rows = gt, cols = prediction, ACCUMULATED into
 *          (atomics); fold it into the int64 matrix with zutis_hist_merge.  Requires Q <= n_classes and
 *          B*H*W < 2^31.
 * ------------------------------------------------------------------------------------------- */
ZUTIS_API int zutis_decode_score(const float* logits, long sb, long sq, long sy, long sx,
                                 int B, int Q, int h, int w, int H, int W,
                                 const void* gt, int gt_dtype, long gt_sb,
                                 int16_t* labels, int32_t* hist_partial, int n_classes,
                                 int mode, void* stream);

/* Same contract with a caller-provided device workspace (>= zutis_decode_workspace_bytes, 8-byte aligned).  The
 * per-cell pruning kernel (csrc/decode_cells.cu; ZUTIS_DECODE_AUTO picks it for pixel-major logits -- sq == 1,
 * 16-byte aligned pixels -- and >= 4x up-sampling) then hands its work out through a global counter kept there, which
 * balances the SMs; without a workspace it distributes cell rows statically.  Results are identical either way and
 * identical to every other mode.  The kernel leaves the workspace zeroed; pass ZUTIS_DECODE_WORKSPACE_ZEROED when it
 * is known to be zero and the call enqueues no memset.  workspace == NULL behaves like zutis_decode_score. */
ZUTIS_API size_t zutis_decode_workspace_bytes(int B, int Q, int h, int w, int H, int W);
ZUTIS_API int zutis_decode_score_ws(const float* logits, long sb, long sq, long sy, long sx,
                                    int B, int Q, int h, int w, int H, int W,
                                    const void* gt, int gt_dtype, long gt_sb,
                                    int16_t* labels, int32_t* hist_partial, int n_classes,
                                    int mode, void* workspace, size_t workspace_bytes, void* stream);

/* Histogram only, for labels that already exist (RunningScore.update with device tensors).
 * pred: int16/int32/int64/u8 labels (pred_dtype uses zutis_gt_dtype), same element count as gt. */
ZUTIS_API int zutis_score_labels(const void* gt, int gt_dtype, const void* pred, int pred_dtype,
                                 long n_pixels, int32_t* hist_partial, int n_classes, void* stream);

/* Small global merge: hist_i64[i] += sum_j partials[j*n2 + i]; partials are cleared when
 * clear_partials != 0 so they can be reused by the next launch. */
ZUTIS_API int zutis_hist_merge(int32_t* partials, int n_partials, long long* hist_i64, long n2,
                               int clear_partials, void* stream);

/* Multi-GPU sum of the per-GPU matrices: ncclAllReduce(sum, int64, n2) in place on `stream` over
 * the caller's communicator (an ncclComm_t passed as void*).  The library does not link NCCL; it takes ncclAllReduce from
 * the NCCL already loaded in the process, i.e. the one the communicator belongs to.  Integer sums make the result
 * bit-identical for any number of GPUs.  (Python callers without a raw communicator use RunningScore.all_reduce(), which
 * goes through torch.distributed.) */
ZUTIS_API int zutis_allreduce_hist(long long* hist_i64, long n2, void* nccl_comm, void* stream);

/* The same sum over NVLink peer memory, for small matrices (csrc/p2p_reduce.cu): every rank stages its matrix in a buffer
 * its peers have mapped through CUDA IPC, and one kernel per rank signals, waits, sums all ranks' matrices with peer loads
 * and signals again.  52 KB (Q = 81): ~8-10 us against 18-30 us for the library all-reduce.  One process per GPU of one node.
 *   zutis_p2p_create   allocates this rank's block for matrices of up to max_n2 elements and returns its 64-byte IPC handle;
 *   zutis_p2p_connect  takes all ranks' handles (world x 64 bytes, rank order; exchange them however the host likes);
 *   zutis_allreduce_hist_p2p  out[i] = sum_r hist_r[i] on `stream` (out may be hist); every rank must make the same calls;
 *                      the epoch of the handshake lives on the device, so the launch can be captured in a CUDA graph;
 *   zutis_p2p_destroy  unmaps and frees. */
ZUTIS_API int zutis_p2p_create(int world, int rank, long max_n2, unsigned char* ipc_handle_out, int* ctx_out);
ZUTIS_API int zutis_p2p_connect(int ctx, const unsigned char* handles);
ZUTIS_API int zutis_allreduce_hist_p2p(int ctx, const long long* hist, long n2, long long* out, void* stream);
/* ... with this rank's pending int32 partial folded in first (hist += partial; partial = 0), i.e. zutis_hist_merge and the
 * all-reduce in one launch; out must not be hist. */
ZUTIS_API int zutis_merge_allreduce_hist_p2p(int ctx, long long* hist, int32_t* partial, long n2, long long* out, void* stream);
ZUTIS_API int zutis_p2p_destroy(int ctx);

/* ---------------------------------------------------------------------------------------------
 * return_logits=True (networks/zutis.py:369-370): the one mode where full-resolution fp32 logits
 * are materialised on request.  out is [B,Q,H,W] contiguous.
 * ------------------------------------------------------------------------------------------- */
ZUTIS_API int zutis_upsample_bilinear(const float* in, long sb, long sq, long sy, long sx,
                                      int B, int Q, int h, int w, int H, int W, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Instance path: interp(probabilities) > threshold                networks/zutis.py:422-425 (:390 when H==h, W==w)
 * mask_bits : uint32 [B,Q,H,words] with words = (W+31)/32, bit (x & 31) of word x>>5 is pixel x.
 * areas     : optional int32 [B,Q], number of set pixels per mask, ACCUMULATED into.
 * ------------------------------------------------------------------------------------------- */
ZUTIS_API int zutis_decode_threshold(const float* probs, long sb, long sq, long sy, long sx,
                                     int B, int Q, int h, int w, int H, int W, float threshold,
                                     uint32_t* mask_bits, int32_t* areas, void* stream);

/* Expand bit-packed masks to one byte per pixel (bool [n_masks,H,W]) for legacy consumers. */
ZUTIS_API int zutis_unpack_mask_bits(const uint32_t* mask_bits, long n_masks, int H, int W,
                                     uint8_t* out_bytes, void* stream);

/* compute_iou on bit-packed masks (utils/iou.py:6-38 inside NMS, networks/zutis.py:258-259):
 * inter[i*M+j] = popcount(mask_i & mask_j) for all pairs of the M masks of one image; the diagonal
 * holds the areas, so union = area_i + area_j - inter.  words_per_mask = H * ((W+31)/32). */
ZUTIS_API int zutis_pairwise_mask_intersections(const uint32_t* mask_bits, int M, long words_per_mask,
                                                int32_t* inter, void* stream);

/* Greedy per-category hard NMS on the device (networks/zutis.py:245-278, nms_type="hard"), driven by the intersection
 * counts of zutis_pairwise_mask_intersections for B images ([B,M,M], diagonal = areas).
 * categories [B,M] int32 (0 = background, skipped), scores [B,M] fp32 (hard NMS leaves them unchanged).
 * pick_rank[b,i] = round in which mask i was chosen for its category (emission order inside the category), -1 when it
 * was suppressed (IoU with a chosen mask > iou_threshold, float64 like utils/iou.py) or dropped (score <= score_floor
 * after the first pick).  tie[b] != 0: two candidates of a category had equal scores at a pick -- the reference's choice
 * then follows numpy's unstable argsort and the caller should replay that image with the reference loop. */
ZUTIS_API int zutis_instance_nms_hard(const int32_t* inter, const int32_t* categories, const float* scores, int B, int M,
                                      double iou_threshold, float score_floor, int32_t* pick_rank, int32_t* tie, void* stream);

/* COCO run-length encoding + bounding boxes of bit-packed masks, on the device.
 * Replaces pycocotools.mask.encode(np.asfortranarray(m)) (networks/zutis.py:290; cocoapi rleEncode: the mask flattened
 * column by column, alternating run lengths starting with a possibly empty run of zeros) and
 * torchvision.ops.masks_to_boxes (networks/zutis.py:294) for the masks that survive NMS.
 * mask_bits: row-packed masks, mask i at mask_bits + id(i) * mask_stride_words, id(i) = mask_ids ? mask_ids[i] : i,
 * each [H][(W+31)/32] words, bit x%32 of word [y][x/32] = pixel (y, x).
 * Count pass (runs == NULL): n_runs[i] = number of runs, boxes[4i..] = {xmin, ymin, xmax, ymax} (inclusive pixel
 * indices, -1 x4 for an empty mask; boxes may be NULL).  Write pass (runs != NULL): the n_runs[i] run lengths of mask
 * i go to runs + run_offsets[i]; the caller derives run_offsets from the count pass.  H*W < 2^31; masks whose
 * column-packed copy exceeds 200 KB of shared memory (about 1400 x 1024) return ZUTIS_ERR_UNSUPPORTED. */
ZUTIS_API int zutis_mask_rle(const uint32_t* mask_bits, long mask_stride_words, const int32_t* mask_ids, int n_masks,
                             int H, int W, const int64_t* run_offsets, uint32_t* runs, int32_t* n_runs, int32_t* boxes,
                             void* stream);

/* cocoapi rleToString on the device: the `counts` bytes of pycocotools.mask.encode (networks/zutis.py:290) from the run
 * lengths zutis_mask_rle wrote.  Mask i's string has string_lengths[i] bytes at strings + string_offsets[i]; the
 * strings are packed back to back in completion order behind *cursor (device counter, zeroed by the caller; its final
 * value is the number of bytes needed -- nothing is written past `capacity`, 7 bytes per run always suffice). */
ZUTIS_API int zutis_rle_to_string(const uint32_t* runs, const int64_t* run_offsets, const int32_t* n_runs, int n_masks,
                                  uint8_t* strings, int64_t capacity, uint64_t* cursor, int64_t* string_offsets,
                                  int32_t* string_lengths, void* stream);

/* Low-resolution instance statistics                              networks/zutis.py:390-406
 * probs [B,Q,h,w] (strides as above), tokens [B,hw,D] channel-last contiguous.
 * sizes[b,q] = #(p > thr); psum[b,q] = sum of in-mask p; mean_tokens[b,q,:] = sum of in-mask tokens / (size+1e-7). */
ZUTIS_API int zutis_instance_lowres_stats(const float* probs, long sb, long sq, long sy, long sx,
                                          const float* tokens, int B, int Q, int h, int w, int D,
                                          float threshold, int32_t* sizes, float* psum, float* mean_tokens,
                                          void* stream);

/* The same statistics with the masked average computed as a [Q, hw] x [hw, D] contraction per image on the tensor cores
 * (0/1 mask operand, tokens through the 3xTF32 split: fp32-grade sums) instead of the reference's [B,Q,h,w,D] broadcast.
 * workspace: >= zutis_instance_stats_workspace_bytes, 256-byte aligned; NULL / too small / a shape the contraction
 * kernel does not take runs zutis_instance_lowres_stats's kernel instead (same results up to summation order). */
ZUTIS_API size_t zutis_instance_stats_workspace_bytes(int B, int Q, int h, int w, int D);
ZUTIS_API int zutis_instance_lowres_stats_ws(const float* probs, long sb, long sq, long sy, long sx,
                                             const float* tokens, int B, int Q, int h, int w, int D,
                                             float threshold, int32_t* sizes, float* psum, float* mean_tokens,
                                             void* workspace, size_t workspace_bytes, void* stream);

/* Category decision per query                                      networks/zutis.py:409-420
 * mean_tokens [n_rows,D] (n_rows = B*Q), text [n_categories,D] unit rows:
 * prob[n] = sigmoid(temperature * <text[n], t/(|t|+1e-7)>); category = first argmax, max_prob = max. */
ZUTIS_API int zutis_instance_categories(const float* mean_tokens, long n_rows, const float* text, int n_categories, int D,
                                        float temperature, int32_t* category, float* max_prob, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SURVEY section 8(f) N1 -- the step right before the contraction: ZUTIS.image_to_text_space, ViT branch with
 * channel_last=True                                                networks/zutis.py:319-322
 *     y = einsum("bhwn,nc->bhwc", tokens, proj)   -> zutis_gemm_logits with A = proj^T (pixel-major output)
 *     y = F.layer_norm(y, y.shape[1:])            -> statistics over ALL h*w*c elements of an image, eps 1e-5, no affine
 *     y = y / (y.norm(dim=-1, keepdim=True) + 1e-7)
 * This entry does the last two lines in place on x [B, pixels, D] (contiguous): per-image moments in a fixed-order
 * two-stage reduction (deterministic), then one warp per pixel normalises its D channels.
 * workspace: at least zutis_image_norm_workspace_bytes(B, pixels, D) bytes of device scratch.
 * ------------------------------------------------------------------------------------------- */
ZUTIS_API size_t zutis_image_norm_workspace_bytes(int B, long pixels, int D);
ZUTIS_API int zutis_image_layernorm_l2norm(float* x, int B, long pixels, int D, int layer_norm, float ln_eps, float l2_eps,
                                           void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * End-to-end entry with HOST buffers (what a caller holding numpy arrays uses; also bench.py's
 * `e2e` leg): text [Q,D], tokens [B,h,w,D] fp32 and gt [B,H,W] live in (ideally pinned) host memory;
 * the call copies them in, runs contraction -> fused decode+score -> merge, and returns the int64
 * confusion matrix (ADDED into hist_host) and optionally the int16 labels.  Synchronous.
 * ------------------------------------------------------------------------------------------- */
ZUTIS_API int zutis_semantic_eval_host(const float* text, const float* tokens, const void* gt, int gt_dtype,
                                       int B, int Q, int D, int h, int w, int H, int W,
                                       long long* hist_host, int16_t* labels_host, int gemm_flags, int device);
/* Bytes one call of zutis_semantic_eval_host moves host -> device.  Integer ground truth crosses PCIe in the narrowest
 * type that leaves the confusion matrix unchanged (labels outside [0, Q) are ignored by running_score.py:11 and all map
 * to one out-of-range sentinel): uint8 for Q <= 255, else int16; host threads narrow it while the tokens are copied. */
ZUTIS_API size_t zutis_semantic_eval_host_h2d_bytes(int gt_dtype, int want_hist, int B, int Q, int D, int h, int w, int H, int W);

#ifdef __cplusplus
}
#endif
#endif /* ZUTIS_B200_H_ */
