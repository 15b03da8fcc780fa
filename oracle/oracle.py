"""CPU oracle for the ZUTIS mask-decode + scoring path -- TEST INFRASTRUCTURE ONLY.

Two independent restatements of the reference arithmetic live here:

* ``c_*``  -- thin ctypes wrappers over oracle/zutis_oracle.c (plain C, explicit fmaf).
* ``torch_*`` -- the reference's own expressions restated on torch-CPU / numpy ops, i.e.
  the same third-party kernels (ATen einsum / upsample_bilinear2d / argmax, np.bincount)
  the reference calls.  This is also what bench.py times as the CPU baseline
  (``cpu_baseline.kind == "port"``): /root/reference does not exist on the GPU box.

Reference lines restated (NoelShin/zutis): networks/zutis.py:355-372 (semantic decode),
:177-209 (mask proposals), :374-427 (instance decode), :211-299 (hard NMS),
utils/running_score.py:5-50, utils/iou.py:6-38.

Parity pinning: the reference has no tests/golden vectors for this path ("parity
unpinned" upstream).  Both restatements are pinned against outputs of the reference code
itself, captured by tests/golden/make_golden.py into tests/golden/*.npz.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib() -> C.CDLL:
    """Load (building on first use) oracle/libzutis_oracle.so."""
    global _LIB
    if _LIB is None:
        from . import build_oracle
        path = build_oracle.build()
        L = C.CDLL(path)
        L.zo_max_threads.restype = C.c_int
        L.zo_fast_hist.restype = C.c_long
        L.zo_bilinear_f32.restype = C.c_int
        L.zo_decode_semantic.restype = C.c_int
        L.zo_decode_threshold.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _size_pair(size) -> Optional[Tuple[int, int]]:
    """Accept (int,int), torch.Size slices or a pair of 1-element tensors (trainer.py:322-325)."""
    if size is None:
        return None
    H, W = size
    return int(H), int(W)


# --------------------------------------------------------------------------- C restatement
def c_axis_table(n_in: int, n_out: int):
    i0 = np.empty(n_out, np.int32); i1 = np.empty(n_out, np.int32)
    l0 = np.empty(n_out, np.float32); l1 = np.empty(n_out, np.float32)
    lib().zo_axis_table(C.c_int(n_in), C.c_int(n_out), _p(i0, C.c_int32), _p(i1, C.c_int32),
                        _p(l0, C.c_float), _p(l1, C.c_float))
    return i0, i1, l0, l1


def c_bilinear(x: np.ndarray, size) -> np.ndarray:
    """x [..., h, w] fp32 -> [..., H, W] (ATen bilinear, align_corners=False)."""
    H, W = _size_pair(size)
    x = np.ascontiguousarray(x, np.float32)
    h, w = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2], dtype=np.int64))
    out = np.empty(x.shape[:-2] + (H, W), np.float32)
    rc = lib().zo_bilinear_f32(_p(x, C.c_float), C.c_long(planes), C.c_int(h), C.c_int(w),
                               _p(out, C.c_float), C.c_int(H), C.c_int(W))
    assert rc == 0
    return out


def c_logits(text: np.ndarray, tokens: np.ndarray, pixel_major: bool = False) -> np.ndarray:
    """text [Q,D], tokens [B,h,w,D] -> [B,Q,h,w] (or [B,h,w,Q] if pixel_major)."""
    text = np.ascontiguousarray(text, np.float32)
    tokens = np.ascontiguousarray(tokens, np.float32)
    B, h, w, D = tokens.shape
    Q = text.shape[0]
    if pixel_major:
        out = np.empty((B, h, w, Q), np.float32)
        lib().zo_logits(_p(text, C.c_float), _p(tokens, C.c_float), C.c_long(B * h * w), C.c_int(Q),
                        C.c_int(D), _p(out, C.c_float), C.c_int(1))
        return out
    out = np.empty((B, Q, h, w), np.float32)
    for b in range(B):
        lib().zo_logits(_p(text, C.c_float), _p(tokens[b], C.c_float), C.c_long(h * w), C.c_int(Q),
                        C.c_int(D), _p(out[b], C.c_float), C.c_int(0))
    return out


def c_decode_semantic(logits: np.ndarray, size=None) -> np.ndarray:
    """logits [B,Q,h,w] -> int64 labels [B,H,W] (or [B,h,w] when size is None)."""
    logits = np.ascontiguousarray(logits, np.float32)
    B, Q, h, w = logits.shape
    hw = _size_pair(size)
    H, W = hw if hw is not None else (0, 0)
    out = np.empty((B, H, W) if hw is not None else (B, h, w), np.int64)
    rc = lib().zo_decode_semantic(_p(logits, C.c_float), C.c_int(B), C.c_int(Q), C.c_int(h), C.c_int(w),
                                  C.c_int(H), C.c_int(W), _p(out, C.c_int64))
    assert rc == 0
    return out


def c_decode_threshold(probs: np.ndarray, size=None, threshold: float = 0.5) -> np.ndarray:
    """probs [B,Q,h,w] -> bool masks [B,Q,H,W]: interp(p) > threshold (zutis.py:424-425)."""
    probs = np.ascontiguousarray(probs, np.float32)
    B, Q, h, w = probs.shape
    hw = _size_pair(size)
    H, W = hw if hw is not None else (0, 0)
    out = np.empty((B, Q, H, W) if hw is not None else (B, Q, h, w), np.uint8)
    rc = lib().zo_decode_threshold(_p(probs, C.c_float), C.c_int(B), C.c_int(Q), C.c_int(h), C.c_int(w),
                                   C.c_int(H), C.c_int(W), C.c_float(threshold), _p(out, C.c_uint8))
    assert rc == 0
    return out.astype(bool)


def c_fast_hist(gt: np.ndarray, pred: np.ndarray, n: int, hist: Optional[np.ndarray] = None) -> np.ndarray:
    """int64 [n,n] confusion counts (running_score.py:10-16); accumulates into ``hist``."""
    g = np.ascontiguousarray(gt, np.int64).ravel()
    p = np.ascontiguousarray(pred, np.int64).ravel()
    assert g.size == p.size
    if hist is None:
        hist = np.zeros((n, n), np.int64)
    bad = lib().zo_fast_hist(_p(g, C.c_int64), _p(p, C.c_int64), C.c_long(g.size), C.c_int(n),
                             _p(hist, C.c_int64))
    assert bad == 0, f"{bad} predictions outside [0,{n})"
    return hist


def c_mask_iou(pred: np.ndarray, gt: np.ndarray, threshold: Optional[float] = None, eps: float = 1e-7) -> float:
    """compute_iou (iou.py:6-38), numpy branch, as a python float (float64 division)."""
    assert pred.shape == gt.shape and pred.ndim == 2
    p = np.ascontiguousarray(pred, np.float32).ravel()
    g = np.ascontiguousarray(gt, np.float32).ravel()
    inter = C.c_int64(0); uni = C.c_int64(0)
    lib().zo_mask_iou_counts(_p(p, C.c_float), _p(g, C.c_float), C.c_long(p.size),
                             C.c_int(threshold is not None), C.c_float(threshold or 0.0),
                             C.byref(inter), C.byref(uni))
    return inter.value / (uni.value + eps)


# ---------------------------------------------------------------- scoring (numpy, float64)
def scores_from_hist(hist: np.ndarray) -> Tuple[Dict[str, float], Dict[int, float]]:
    """RunningScore.get_scores (running_score.py:22-47) on a [n,n] count matrix.

    Same float64 operations in the same order, so results are bit-identical to the
    reference for the same counts (NaN for absent classes, skipped by nanmean).
    """
    m = np.asarray(hist, dtype=np.float64)
    n = m.shape[0]
    tp = np.diag(m)
    gt_count = m.sum(axis=1)
    pred_count = m.sum(axis=0)
    total = m.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        pixel_acc = tp.sum() / total
        mean_acc = np.nanmean(tp / gt_count)
        iou = tp / (gt_count + pred_count - tp)
        mean_iou = np.nanmean(iou)
        freq = gt_count / total
    seen = freq > 0
    fw_acc = (freq[seen] * iou[seen]).sum()
    summary = {"Pixel Acc": pixel_acc, "Mean Acc": mean_acc, "FreqW Acc": fw_acc, "Mean IoU": mean_iou}
    return summary, dict(zip(range(n), iou))


class OracleRunningScore:
    """running_score.py:5-50 restated: float64 [n,n] matrix, per-image np.bincount."""

    def __init__(self, n_classes: int):
        self.n_classes = n_classes
        self.confusion_matrix = np.zeros((n_classes, n_classes))

    def update(self, label_trues: Iterable[np.ndarray], label_preds: Iterable[np.ndarray]) -> None:
        n = self.n_classes
        for t, p in zip(label_trues, label_preds):
            t = np.asarray(t).ravel(); p = np.asarray(p).ravel()
            keep = (t >= 0) & (t < n)
            self.confusion_matrix += np.bincount(n * t[keep].astype(int) + p[keep], minlength=n * n).reshape(n, n)

    def get_scores(self):
        return scores_from_hist(self.confusion_matrix)

    def reset(self) -> None:
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))


# ----------------------------------------------------- torch-CPU restatement ("the port")
def torch_semantic_predict(text, tokens, size=None, return_logits: bool = False):
    """zutis.py:355-372 on torch CPU ops: einsum -> F.interpolate(bilinear) -> argmax."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        lo = torch.einsum("nc,bchw->bnhw", text, tokens.permute(0, 3, 1, 2))
        if size is not None:
            lo = F.interpolate(lo, size=size, mode="bilinear")
        if return_logits:
            return lo
        return torch.argmax(lo, dim=1).cpu().numpy()


def torch_lowres_logits(text, tokens):
    import torch
    with torch.no_grad():
        return torch.einsum("nc,bchw->bnhw", text, tokens.permute(0, 3, 1, 2))


def torch_mask_proposals(queries, feats):
    """zutis.py:184-186/:196-198 + :209: sigmoid(queries . feats), 3-D or 4-D queries."""
    import torch
    with torch.no_grad():
        if queries.dim() == 3:
            return torch.sigmoid(torch.einsum("bqc,bhwc->bqhw", queries, feats))
        return torch.sigmoid(torch.einsum("bdqc,bhwc->bdqhw", queries, feats))


def torch_instance_lowres(text, mask_proposals, tokens, threshold: float = 0.5, temperature: float = 5):
    """zutis.py:377-420: confidence [B,Q] and category id [B,Q] from LOW-res masks."""
    import torch
    with torch.no_grad():
        mp = mask_proposals[:, -1] if mask_proposals.dim() == 5 else mask_proposals
        binary = mp > threshold
        sizes = binary.sum(dim=(-2, -1))
        conf = (mp * binary).sum(dim=(-2, -1)) / (sizes + 1e-7)
        avg = (tokens[:, None] * binary[..., None]).sum(dim=(-3, -2)) / (sizes.unsqueeze(-1) + 1e-7)
        avg = avg / (avg.norm(dim=-1, keepdim=True) + 1e-7)
        cat_prob = torch.sigmoid(torch.einsum("nc,bqc->bqn", text, avg) * temperature)
        cat = torch.argmax(cat_prob, dim=-1).cpu().numpy()
        conf = (conf * cat_prob.max(dim=-1).values).cpu().numpy()
        return conf, cat, sizes.cpu().numpy()


def torch_instance_masks(mask_proposals, size=None, threshold: float = 0.5) -> np.ndarray:
    """zutis.py:422-427: interp(probabilities) > threshold -> host bool [B,Q,H,W]."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        mp = mask_proposals[:, -1] if mask_proposals.dim() == 5 else mask_proposals
        if size is not None:
            mp = F.interpolate(mp, size=size, mode="bilinear")
        return (mp > threshold).cpu().numpy()


def torch_image_to_text_space(tokens, proj, layer_norm: bool = True):
    """zutis.py:319-322 (ViT, channel_last=True): projection, joint layer norm over (h,w,c), per-pixel L2 norm."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        y = torch.einsum("bhwn,nc->bhwc", tokens, proj)
        if layer_norm:
            y = F.layer_norm(y, normalized_shape=y.shape[1:])
        return y / (y.norm(dim=-1, keepdim=True) + 1e-7)


def hard_nms(masks: np.ndarray, scores: np.ndarray, cats: np.ndarray,
             nms_threshold: float = 0.3, score_floor: float = 0.001, nms_type: str = "hard",
             sigma: float = 0.5) -> List[Tuple[int, int, float]]:
    """zutis.py:225-282 -> list of (category, query index, score) kept; nms_type "hard" (the default of predict),
    "linear" (score *= 1 - iou above the threshold) or "gaussian" (score *= exp(-iou^2 / sigma)), :261-266.

    Per category (0 = background skipped, iterated in ``set`` order like the reference):
    repeatedly keep the best-scoring candidate, drop candidates whose IoU with it exceeds
    the threshold, keep the rest if their score exceeds the floor; empty masks are dropped.
    """
    kept: List[Tuple[int, int, float]] = []
    for cat in set(cats):
        if cat == 0:
            continue
        cand = list(np.nonzero(cats == cat)[0])
        cand_scores = [scores[i] for i in cand]
        chosen: List[Tuple[int, float]] = []
        while len(cand) > 0:
            order = np.argsort(np.array(cand_scores))
            cand = [cand[i] for i in order]
            cand_scores = [cand_scores[i] for i in order]
            best, best_score = cand[-1], cand_scores[-1]
            chosen.append((best, best_score))
            nxt, nxt_scores = [], []
            for i, s in zip(cand[:-1], cand_scores[:-1]):
                inter = np.logical_and(masks[i], masks[best]).sum()
                union = np.logical_or(masks[i], masks[best]).sum()
                iou = inter / (union + 1e-7)
                if nms_type == "hard":
                    weight = 0 if iou > nms_threshold else 1
                elif nms_type == "linear":
                    weight = (1 - iou) if iou > nms_threshold else 1
                else:
                    weight = np.exp(-(iou * iou) / sigma)
                s = s * weight
                if s > score_floor:
                    nxt.append(i); nxt_scores.append(s)
            cand, cand_scores = nxt, nxt_scores
        for i, s in chosen:
            if masks[i].sum() == 0:
                continue
            kept.append((int(cat), int(i), float(s)))
    return kept


# ----------------------------------------------------------------------------------------------- COCO RLE + boxes
# The reference formats every kept instance mask with pycocotools.mask.encode(np.asfortranarray(m))
# (networks/zutis.py:290) and torchvision.ops.masks_to_boxes (:294).  pycocotools is a third-party dependency
# that is NOT vendored under /root/reference and is not installed here; the reference does not pin its version
# (README.md:83: "conda install -c conda-forge pycocotools").  What follows restates the published algorithm of
# cocoapi common/maskApi.c (rleEncode, rleToString, rleFrString), which has not changed since pycocotools 2.0.
# Parity of the compressed string is therefore UNPINNED against the real library; the run lengths are pinned
# against a direct numpy walk of the Fortran-order mask, the string against hand-derived vectors and a
# decode round trip (tests/test_oracle_golden.py).

def rle_counts(mask: np.ndarray) -> List[int]:
    """cocoapi rleEncode: run lengths of the mask flattened column by column, starting with a run of zeros
    (which is 0 long when the first pixel is set)."""
    flat = np.asarray(mask).astype(bool).ravel(order="F")
    counts: List[int] = []
    prev, run = False, 0
    for v in flat.tolist():
        if v != prev:
            counts.append(run)
            run, prev = 0, v
        run += 1
    counts.append(run)
    return counts


def rle_counts_numpy(mask: np.ndarray) -> np.ndarray:
    """Same run lengths, vectorised (for full-size masks)."""
    flat = np.asarray(mask).astype(np.uint8).ravel(order="F")
    change = np.flatnonzero(np.diff(flat)) + 1
    bounds = np.concatenate(([0], change, [flat.size]))
    runs = np.diff(bounds)
    if flat.size and flat[0] == 1:
        runs = np.concatenate(([0], runs))
    return runs.astype(np.int64)


def rle_to_string(counts) -> bytes:
    """cocoapi rleToString: each count (from the fourth on: its difference to the count two places back) is
    written as little-endian groups of 5 bits, bit 5 = "more groups follow", offset by 48 into printable ASCII;
    negative differences are sign-extended (a group with bit 4 set ends the number when the rest is -1)."""
    out = bytearray()
    counts = [int(c) for c in counts]
    for i, x in enumerate(counts):
        if i > 2:
            x -= counts[i - 2]
        more = True
        while more:
            c = x & 0x1F
            x >>= 5                                  # arithmetic shift: Python ints behave like C's signed long here
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def rle_from_string(s: bytes) -> List[int]:
    """cocoapi rleFrString (the inverse), used for round-trip checks."""
    counts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def mask_to_box(mask: np.ndarray) -> List[float]:
    """torchvision.ops.masks_to_boxes for one mask (networks/zutis.py:294): [xmin, ymin, xmax, ymax] of the set
    pixels as floats, inclusive indices."""
    ys, xs = np.nonzero(np.asarray(mask))
    return [float(xs.min()), float(ys.min()), float(xs.max()), float(ys.max())]
