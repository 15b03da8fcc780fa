"""Tensor-level wrappers over the C ABI: torch owns memory and streams, the kernels do the work.

Every function here takes CUDA tensors, passes raw pointers + element strides to
libzutis_b200.so on the caller's current stream, and returns CUDA tensors.  Nothing in this
module computes with torch ops.
"""
from __future__ import annotations

from typing import Optional, Tuple

import ctypes as C

import numpy as np
import torch

from . import _ffi as F

_GT_CODES = {torch.uint8: F.GT_U8, torch.int16: F.GT_I16, torch.int32: F.GT_I32, torch.int64: F.GT_I64}
_PRECISIONS = {"fp32": F.GEMM_FP32_SIMT, "tf32x3": F.GEMM_TF32X3, "tf32": F.GEMM_TF32}

# Contraction precision used when the caller does not choose.  "auto" = the tcgen05 kernel with the
# 3-term error-compensated TF32 split (fp32-grade, needed for the 99.99 % label bar) whenever the
# shape qualifies (K % 32 == 0, aligned K-contiguous rows), else the fp32 FFMA kernel.  Both are
# sm_100a CUDA kernels of this library; neither is a library or CPU fallback.
DEFAULT_PRECISION = "auto"


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (zutis_b200 has no CPU path)")


def _fp32_4d(t: torch.Tensor, name: str) -> torch.Tensor:
    """The kernels read raw fp32: half-precision inputs (autocast) are widened, anything else is refused."""
    if t.dim() != 4:
        raise ValueError(f"{name} must be 4-D, got {tuple(t.shape)}")
    if not t.is_floating_point():
        raise TypeError(f"{name} must be a floating-point tensor, got {t.dtype}")
    return t if t.dtype == torch.float32 else t.float()


def size_pair(size) -> Optional[Tuple[int, int]]:
    """(H, W) from (int,int), a torch.Size slice or a pair of 1-element tensors (trainer.py:322-325)."""
    if size is None:
        return None
    H, W = size
    return int(H), int(W)


def gemm_flags(precision: Optional[str], sigmoid: bool = False) -> int:
    p = DEFAULT_PRECISION if precision is None else precision
    if p == "auto":
        p = "tf32x3"
    if p not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {p!r}")
    return _PRECISIONS[p] | (F.GEMM_SIGMOID if sigmoid else 0)


class DecodeWorkspace:
    """The 16 bytes of device scratch behind ``zutis_decode_score_ws``: the cell decode kernel's global run counter.

    One buffer per (device, stream), since two launches that overlap must not share a counter.  A buffer is zero-filled
    when it is created and every launch leaves it zero, so calls through it enqueue no memset."""

    def __init__(self):
        self.bufs: dict = {}

    def ensure(self, nbytes: int, device) -> torch.Tensor:
        key = (str(device), _stream())
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = self.bufs[key] = torch.empty(max(nbytes, 256), device=device, dtype=torch.uint8)
            buf[:256].zero_()
        return buf


_default_workspace = None           # for callers that do not bring one


def contraction(a: torch.Tensor, feats: torch.Tensor, *, precision: Optional[str] = None, sigmoid: bool = False,
                pixel_major: bool = True, a_cache: Optional[dict] = None) -> torch.Tensor:
    """out[b,n,y,x] = act(sum_c a[(b,)n,c] * feats[b,y,x,c])   (zutis.py:361-365, :184-186 + :209).

    a: [M,C] shared by the batch (text embeddings) or [B,M,C] per image (queries).
    feats: [B,h,w,C] channel-last.  Returns a [B,M,h,w] tensor; with ``pixel_major`` its memory
    is [B,h,w,Mp] (category index contiguous, Mp = M rounded up to 4) -- the layout the fused
    decode kernel streams -- otherwise it is a contiguous [B,M,h,w].
    ``a_cache``: a dict owned by the caller; for a batch-shared ``a`` (text embeddings, constant per
    model) the tensor-core kernel's prepared operand is kept there and reused by later calls.
    """
    _need_cuda(a, "a"); _need_cuda(feats, "feats")
    if feats.dim() != 4:
        raise ValueError(f"feats must be [B,h,w,C], got {tuple(feats.shape)}")
    a = a.float(); feats = feats.float()
    if feats.stride(-1) != 1 or feats.stride(1) != feats.shape[2] * feats.stride(2) or feats.stride(0) != feats.shape[1] * feats.stride(1):
        feats = feats.contiguous()
    if a.stride(-1) != 1:
        a = a.contiguous()
    B, h, w, Cc = feats.shape
    shared = a.dim() == 2
    if not shared and (a.dim() != 3 or a.shape[0] != B):
        raise ValueError(f"a must be [M,C] or [B,M,C] with B={B}, got {tuple(a.shape)}")
    if not shared and a.stride(0) != a.shape[1] * a.stride(1):
        a = a.contiguous()
    M = a.shape[-2]
    if a.shape[-1] != Cc:
        raise ValueError(f"contraction width mismatch: {a.shape[-1]} vs {Cc}")
    N = h * w
    flags = gemm_flags(precision, sigmoid)
    if pixel_major:
        Mp = (M + 3) & ~3
        # the padding columns [M, Mp) are never exposed (`out` below) and every kernel that streams the buffer
        # (decode_score, decode_threshold) ignores them, so they are left as the allocator returned them
        buf = torch.empty((B, h, w, Mp), device=feats.device, dtype=torch.float32)
        s_cn, s_cp, s_c = 1, Mp, N * Mp
        out = buf[..., :M].permute(0, 3, 1, 2)
    else:
        buf = torch.empty((B, M, h, w), device=feats.device, dtype=torch.float32)
        s_cn, s_cp, s_c = N, 1, M * N
        out = buf
    def launch(fl: int) -> None:
        ws_bytes = F.lib().zutis_gemm_workspace_bytes(M, N, Cc, 1 if shared else B, fl)
        ws = None
        cache_key = None
        if ws_bytes and shared and a_cache is not None:
            # The prepared operand belongs to THIS tensor object at THIS version.  The entry keeps a reference to the
            # tensor and is compared with `is`: a data pointer would be recycled by the caching allocator when
            # update_text_embeddings (zutis.py:333-338) replaces the embeddings by a tensor of the same size.
            cache_key = (M, Cc, fl & F.GEMM_TF32X3 | fl & F.GEMM_TF32, str(feats.device))
            hit = a_cache.get("entry")
            if hit is not None and hit[0] is a and hit[1] == a._version and hit[2] == cache_key:
                ws = hit[3]
                fl |= F.GEMM_A_PREPARED
            else:
                a_cache.clear()
                ws = torch.empty(ws_bytes, device=feats.device, dtype=torch.uint8)
        elif ws_bytes:
            ws = torch.empty(ws_bytes, device=feats.device, dtype=torch.uint8)
        with torch.cuda.device(feats.device):
            F.call("zutis_gemm_logits", a.data_ptr(), a.stride(-2), 0 if shared else a.stride(0),
                   feats.data_ptr(), feats.stride(2), feats.stride(0),
                   buf.data_ptr(), s_cn, s_cp, s_c, M, N, Cc, B, fl,
                   ws.data_ptr() if ws is not None else None, ws_bytes, _stream())
        if cache_key is not None and not (fl & F.GEMM_A_PREPARED):
            a_cache["entry"] = (a, a._version, cache_key, ws)    # only once the launch that prepared it has succeeded

    if (precision or DEFAULT_PRECISION) == "auto":
        try:
            launch(gemm_flags("tf32x3", sigmoid))
        except F.ZutisUnsupported:
            launch(gemm_flags("fp32", sigmoid))
    else:
        launch(flags)
    return out


def decode_score(logits: torch.Tensor, size=None, *, gt: Optional[torch.Tensor] = None,
                 hist_partial: Optional[torch.Tensor] = None, n_classes: Optional[int] = None,
                 want_labels: bool = True, mode: int = F.DECODE_AUTO,
                 workspace: Optional[DecodeWorkspace] = None) -> Optional[torch.Tensor]:
    """Fused upsample -> argmax -> (labels, confusion counts)   (zutis.py:366-372 + running_score.py:10-16).

    logits [B,Q,h,w] fp32 with any strides; gt [B,H,W] integer CUDA tensor (or None);
    hist_partial int32 [n_classes*n_classes] accumulated into.  Returns int16 labels [B,H,W] or None.
    """
    _need_cuda(logits, "logits")
    if logits.dim() != 4 or logits.dtype != torch.float32:
        raise ValueError(f"logits must be fp32 [B,Q,h,w], got {logits.dtype} {tuple(logits.shape)}")
    B, Q, h, w = logits.shape
    hw = size_pair(size)
    H, W = hw if hw is not None else (h, w)
    labels = torch.empty((B, H, W), device=logits.device, dtype=torch.int16) if want_labels else None
    gt_ptr, gt_code, gt_sb = None, F.GT_I64, H * W
    if hist_partial is not None:
        if gt is None:
            raise ValueError("hist_partial needs gt")
        _need_cuda(gt, "gt"); _need_cuda(hist_partial, "hist_partial")
        if gt.dtype not in _GT_CODES:
            raise TypeError(f"gt dtype {gt.dtype} not supported (uint8/int16/int32/int64)")
        if tuple(gt.shape) != (B, H, W):
            raise ValueError(f"gt must be [B,H,W]={B, H, W}, got {tuple(gt.shape)}")
        if gt.stride(2) != 1 or gt.stride(1) != W:
            gt = gt.contiguous()
        gt_ptr, gt_code, gt_sb = gt.data_ptr(), _GT_CODES[gt.dtype], gt.stride(0) if B > 1 else H * W
        n_classes = Q if n_classes is None else n_classes
        if hist_partial.dtype != torch.int32 or hist_partial.numel() != n_classes * n_classes or not hist_partial.is_contiguous():
            raise ValueError("hist_partial must be a contiguous int32 tensor with n_classes^2 elements")
    with torch.cuda.device(logits.device):
        # the cell kernel's run counter (dynamic work distribution); the library ignores it when another kernel runs
        ws_bytes = F.lib().zutis_decode_workspace_bytes(B, Q, h, w, H, W)
        if workspace is None:
            global _default_workspace
            if _default_workspace is None:
                _default_workspace = DecodeWorkspace()
            workspace = _default_workspace
        ws = workspace.ensure(ws_bytes, logits.device)
        mode |= F.DECODE_WORKSPACE_ZEROED                    # created zero-filled, left zero by every launch
        F.call("zutis_decode_score_ws", logits.data_ptr(), logits.stride(0), logits.stride(1), logits.stride(2), logits.stride(3),
               B, Q, h, w, H, W, gt_ptr, gt_code, gt_sb,
               labels.data_ptr() if labels is not None else None,
               hist_partial.data_ptr() if hist_partial is not None else None,
               n_classes if hist_partial is not None else 0, mode, ws.data_ptr(), ws_bytes, _stream())
    return labels


def score_labels(gt: torch.Tensor, pred: torch.Tensor, hist_partial: torch.Tensor, n_classes: int) -> None:
    """RunningScore._fast_hist on device labels (any supported integer dtypes, same element count)."""
    _need_cuda(gt, "gt"); _need_cuda(pred, "pred"); _need_cuda(hist_partial, "hist_partial")
    if gt.dtype not in _GT_CODES or pred.dtype not in _GT_CODES:
        raise TypeError(f"label dtypes must be uint8/int16/int32/int64, got {gt.dtype}, {pred.dtype}")
    gt = gt.contiguous(); pred = pred.contiguous()
    if gt.numel() != pred.numel():
        raise ValueError(f"gt and pred differ in size: {gt.numel()} vs {pred.numel()}")
    with torch.cuda.device(gt.device):
        F.call("zutis_score_labels", gt.data_ptr(), _GT_CODES[gt.dtype], pred.data_ptr(), _GT_CODES[pred.dtype],
               gt.numel(), hist_partial.data_ptr(), n_classes, _stream())


def hist_merge(partials: torch.Tensor, hist_i64: torch.Tensor, clear: bool = True) -> None:
    n2 = hist_i64.numel()
    with torch.cuda.device(hist_i64.device):
        F.call("zutis_hist_merge", partials.data_ptr(), partials.numel() // n2, hist_i64.data_ptr(), n2, int(clear), _stream())


def upsample_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    """F.interpolate(x, size, mode="bilinear") materialised (return_logits=True, zutis.py:366-370)."""
    _need_cuda(x, "x")
    x = _fp32_4d(x, "x")
    B, Q, h, w = x.shape
    H, W = size_pair(size)
    out = torch.empty((B, Q, H, W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        F.call("zutis_upsample_bilinear", x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), x.stride(3),
               B, Q, h, w, H, W, out.data_ptr(), _stream())
    return out


def decode_threshold(probs: torch.Tensor, size=None, threshold: float = 0.5, want_areas: bool = True):
    """interp(probs) > threshold as bit-packed masks uint32-in-int32 [B,Q,H,words] (+ int32 areas [B,Q])."""
    _need_cuda(probs, "probs")
    probs = _fp32_4d(probs, "probs")
    B, Q, h, w = probs.shape
    hw = size_pair(size)
    H, W = hw if hw is not None else (h, w)
    words = (W + 31) // 32
    bits = torch.empty((B, Q, H, words), device=probs.device, dtype=torch.int32)
    areas = torch.zeros((B, Q), device=probs.device, dtype=torch.int32) if want_areas else None
    with torch.cuda.device(probs.device):
        F.call("zutis_decode_threshold", probs.data_ptr(), probs.stride(0), probs.stride(1), probs.stride(2), probs.stride(3),
               B, Q, h, w, H, W, float(threshold), bits.data_ptr(), areas.data_ptr() if areas is not None else None, _stream())
    return bits, areas


def unpack_mask_bits(bits: torch.Tensor, W: int) -> torch.Tensor:
    """bit-packed [...,H,words] -> bool [...,H,W]."""
    H = bits.shape[-2]
    n = bits.numel() // (H * bits.shape[-1])
    out = torch.empty(bits.shape[:-2] + (H, W), device=bits.device, dtype=torch.uint8)
    if n:
        with torch.cuda.device(bits.device):
            F.call("zutis_unpack_mask_bits", bits.contiguous().data_ptr(), n, H, W, out.data_ptr(), _stream())
    return out.view(torch.bool)


def pairwise_mask_intersections(bits: torch.Tensor) -> torch.Tensor:
    """bits [M,H,words] -> int32 [M,M] with popcount(m_i & m_j) (diagonal = areas)."""
    M = bits.shape[0]
    words = bits.numel() // M
    inter = torch.empty((M, M), device=bits.device, dtype=torch.int32)
    with torch.cuda.device(bits.device):
        F.call("zutis_pairwise_mask_intersections", bits.contiguous().data_ptr(), M, words, inter.data_ptr(), _stream())
    return inter


def instance_nms_hard(inter: torch.Tensor, categories: torch.Tensor, scores: torch.Tensor, iou_threshold: float = 0.3,
                      score_floor: float = 0.001):
    """Greedy per-category hard NMS (zutis.py:245-278) for a batch: inter int32 [B,M,M] (diagonal = areas), categories
    int32 [B,M], scores fp32 [B,M], all on the device.  Returns (pick_rank int32 [B,M], tie int32 [B])."""
    _need_cuda(inter, "inter")
    B, M = categories.shape
    pick = torch.empty((B, M), device=inter.device, dtype=torch.int32)
    tie = torch.empty((B,), device=inter.device, dtype=torch.int32)
    with torch.cuda.device(inter.device):
        F.call("zutis_instance_nms_hard", inter.contiguous().data_ptr(), categories.contiguous().data_ptr(), scores.contiguous().data_ptr(),
               B, M, float(iou_threshold), float(score_floor), pick.data_ptr(), tie.data_ptr(), _stream())
    return pick, tie


def _mask_rle_device(bits: torch.Tensor, W: int, mask_ids: Optional[torch.Tensor]):
    """Both passes of zutis_mask_rle; everything but the run counts stays on the device."""
    _need_cuda(bits, "bits")
    H, words = bits.shape[-2], bits.shape[-1]
    if words != (W + 31) // 32:
        raise ValueError(f"bits has {words} words per row, W={W} needs {(W + 31) // 32}")
    bits = bits.contiguous()
    n_all = bits.numel() // (H * words)
    ids = None
    n = n_all
    if mask_ids is not None:
        ids_host = torch.as_tensor(mask_ids, dtype=torch.int64).cpu()
        n = ids_host.numel()
        if n and (int(ids_host.min()) < 0 or int(ids_host.max()) >= n_all):
            raise ValueError("mask_ids out of range")
        ids = ids_host.to(device=bits.device, dtype=torch.int32)
    if n == 0:
        return None
    n_runs = torch.empty(n, device=bits.device, dtype=torch.int32)
    boxes = torch.empty((n, 4), device=bits.device, dtype=torch.int32)
    with torch.cuda.device(bits.device):
        F.call("zutis_mask_rle", bits.data_ptr(), H * words, ids.data_ptr() if ids is not None else None, n, H, W,
               None, None, n_runs.data_ptr(), boxes.data_ptr(), _stream())
        counts = n_runs.cpu().numpy().astype(np.int64)
        offsets = np.cumsum(counts) - counts
        off_dev = torch.from_numpy(offsets).to(bits.device)
        runs = torch.empty(int(counts.sum()), device=bits.device, dtype=torch.int32)
        F.call("zutis_mask_rle", bits.data_ptr(), H * words, ids.data_ptr() if ids is not None else None, n, H, W,
               off_dev.data_ptr(), runs.data_ptr(), n_runs.data_ptr(), None, _stream())
    return counts, n_runs, off_dev, runs, boxes


def mask_rle(bits: torch.Tensor, W: int, mask_ids: Optional[torch.Tensor] = None):
    """COCO run lengths and boxes of bit-packed masks, computed on the device   (zutis.py:290, :294).

    bits: int32 [..., H, words]; ``mask_ids`` (optional) selects and orders masks of the flattened leading dims.
    Returns host arrays ``(n_runs int64 [n], runs uint32 [sum n_runs], boxes int32 [n,4])``: mask i owns
    ``runs[offsets[i] : offsets[i] + n_runs[i]]`` with ``offsets = cumsum(n_runs) - n_runs`` -- the column-major run
    lengths ``pycocotools.mask.encode`` compresses -- and ``boxes[i] = (xmin, ymin, xmax, ymax)`` (``-1`` if empty)."""
    r = _mask_rle_device(bits, W, mask_ids)
    if r is None:
        return np.zeros(0, np.int64), np.zeros(0, np.uint32), np.zeros((0, 4), np.int32)
    counts, _, _, runs, boxes = r
    return counts, runs.cpu().numpy().view(np.uint32), boxes.cpu().numpy()


def mask_rle_strings(bits: torch.Tensor, W: int, mask_ids: Optional[torch.Tensor] = None):
    """``pycocotools.mask.encode(np.asfortranarray(m))["counts"]`` and ``masks_to_boxes`` for bit-packed device masks
    (zutis.py:290, :294): run lengths, their compressed strings and the boxes are all produced on the device; the host
    receives the run counts, then the packed strings -- no boolean mask is copied.  Returns ``(List[bytes], int32 [n,4])``."""
    r = _mask_rle_device(bits, W, mask_ids)
    if r is None:
        return [], np.zeros((0, 4), np.int32)
    counts, n_runs, off_dev, runs, boxes = r
    n = counts.size
    capacity = 7 * int(counts.sum())
    strings = torch.empty(capacity, device=bits.device, dtype=torch.uint8)
    meta = torch.zeros(1 + n, device=bits.device, dtype=torch.int64)          # [cursor | string offsets]
    lengths = torch.empty(n, device=bits.device, dtype=torch.int32)
    with torch.cuda.device(bits.device):
        F.call("zutis_rle_to_string", runs.data_ptr(), off_dev.data_ptr(), n_runs.data_ptr(), n, strings.data_ptr(), capacity,
               meta.data_ptr(), meta.data_ptr() + 8, lengths.data_ptr(), _stream())
    meta_h = meta.cpu().numpy()
    used = int(meta_h[0])
    if used > capacity:
        raise F.ZutisError(F.ERR_WORKSPACE, f"zutis_rle_to_string needed {used} bytes, {capacity} were provided")
    buf = strings[:used].cpu().numpy().tobytes()
    lens = lengths.cpu().numpy()
    return [buf[o:o + l] for o, l in zip(meta_h[1:].tolist(), lens.tolist())], boxes.cpu().numpy()


def instance_lowres_stats(probs: torch.Tensor, tokens: Optional[torch.Tensor], threshold: float = 0.5):
    """sizes int32 [B,Q], psum fp32 [B,Q], mean_tokens fp32 [B,Q,D]   (zutis.py:390-406)."""
    _need_cuda(probs, "probs")
    probs = _fp32_4d(probs, "probs")
    B, Q, h, w = probs.shape
    if tokens is not None:
        _need_cuda(tokens, "tokens")
        if tokens.dim() != 4 or tuple(tokens.shape[:3]) != (B, h, w):
            raise ValueError(f"tokens must be [B,h,w,D] = [{B},{h},{w},D], got {tuple(tokens.shape)}")
    sizes = torch.empty((B, Q), device=probs.device, dtype=torch.int32)
    psum = torch.empty((B, Q), device=probs.device, dtype=torch.float32)
    mean = None
    D = 0
    if tokens is not None:
        tokens = tokens.float().contiguous()
        D = tokens.shape[-1]
        mean = torch.empty((B, Q, D), device=probs.device, dtype=torch.float32)
    with torch.cuda.device(probs.device):
        # the masked average runs on the tensor cores (mask x tokens contraction); scratch: mask matrix, transposed tokens, operand split
        ws_bytes = F.lib().zutis_instance_stats_workspace_bytes(B, Q, h, w, D) if tokens is not None else 0
        ws = torch.empty(ws_bytes, device=probs.device, dtype=torch.uint8) if ws_bytes else None
        F.call("zutis_instance_lowres_stats_ws", probs.data_ptr(), probs.stride(0), probs.stride(1), probs.stride(2), probs.stride(3),
               tokens.data_ptr() if tokens is not None else None, B, Q, h, w, D, float(threshold),
               sizes.data_ptr(), psum.data_ptr(), mean.data_ptr() if mean is not None else None,
               ws.data_ptr() if ws is not None else None, ws_bytes, _stream())
    return sizes, psum, mean


def instance_categories(mean_tokens: torch.Tensor, text: torch.Tensor, temperature: float = 5.0):
    """category int32 [B,Q], max_prob fp32 [B,Q]   (zutis.py:409-420)."""
    B, Q, D = mean_tokens.shape
    text = text.float().contiguous()
    cat = torch.empty((B, Q), device=mean_tokens.device, dtype=torch.int32)
    prob = torch.empty((B, Q), device=mean_tokens.device, dtype=torch.float32)
    with torch.cuda.device(mean_tokens.device):
        F.call("zutis_instance_categories", mean_tokens.contiguous().data_ptr(), B * Q, text.data_ptr(), text.shape[0], D,
               float(temperature), cat.data_ptr(), prob.data_ptr(), _stream())
    return cat, prob


def image_layernorm_l2norm_(x: torch.Tensor, layer_norm: bool = True, ln_eps: float = 1e-5, l2_eps: float = 1e-7) -> torch.Tensor:
    """In place on x [B,h,w,D] (contiguous fp32): joint layer norm over (h,w,D) then per-pixel L2 norm (zutis.py:321-322)."""
    _need_cuda(x, "x")
    if x.dim() != 4 or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("x must be a contiguous fp32 [B,h,w,D] tensor")
    B, h, w, D = x.shape
    ws_bytes = F.lib().zutis_image_norm_workspace_bytes(B, h * w, D)
    ws = torch.empty(max(ws_bytes, 8), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        F.call("zutis_image_layernorm_l2norm", x.data_ptr(), B, h * w, D, int(layer_norm), float(ln_eps), float(l2_eps),
               ws.data_ptr(), ws_bytes, _stream())
    return x
