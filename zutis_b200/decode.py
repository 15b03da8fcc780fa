"""Drop-in for the decode half of the reference's ``networks/zutis.py``.

``predict`` and ``get_mask_proposals`` keep the reference signatures and return types
(networks/zutis.py:340-353 and :177-182) and are written to be bound onto the reference model::

    from zutis_b200 import install; install(ZUTIS)      # ZUTIS.predict now runs on libzutis_b200

or used through ``ZutisDecoder`` (an object holding only ``text_embeddings``).  ``forward`` and everything it
calls are untouched by default, so training keeps its autograd graph (see ``install``).  What changes is what runs underneath:

  semantic (zutis.py:355-372)  einsum -> tcgen05 / FFMA contraction kernel (pixel-major logits);
                               F.interpolate + argmax -> one fused kernel that never writes the
                               [B,Q,H,W] logits; the int16 device labels are widened to the int64
                               numpy array the reference returns only at the very end.
  instance (zutis.py:374-470)  low-res statistics, category decision, full-resolution thresholding
                               into bit-packed masks and the pairwise intersections needed by NMS
                               are kernels; the greedy NMS bookkeeping and the COCO-style dicts
                               stay on the host like in the reference.

Additive fast path (does not exist in the reference): ``decode_and_score`` feeds a
``zutis_b200.RunningScore`` straight from the fused kernel, so neither labels nor logits leave
the device.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .running_score import RunningScore


def _rle_strings(n_runs: np.ndarray, runs: np.ndarray) -> List[bytes]:
    """COCO compressed RLE strings for a batch of masks from their run lengths (``ops.mask_rle``).

    This is cocoapi's ``rleToString`` (what ``pycocotools.mask.encode`` returns as ``counts``, zutis.py:290),
    vectorised over every run of every mask at once: from the fourth run of a mask on, the value written is the
    difference to the run two places back; each value becomes little-endian groups of 5 bits with bit 5 = "more
    groups follow", +48; a group with bit 4 set ends a number once the remaining (sign-extended) value is -1."""
    n_runs = np.asarray(n_runs, dtype=np.int64)
    total = int(n_runs.sum())
    if total == 0:
        return [b"" for _ in n_runs]
    offsets = np.cumsum(n_runs) - n_runs
    x = runs.astype(np.int64)
    idx = np.arange(total, dtype=np.int64) - np.repeat(offsets, n_runs)      # index of the run inside its mask
    late = idx > 2
    x[late] -= runs.astype(np.int64)[np.flatnonzero(late) - 2]
    chars = np.zeros((total, 7), dtype=np.uint8)                             # |x| < 2^31 needs at most 7 groups
    alive = np.ones(total, dtype=bool)
    length = np.zeros(total, dtype=np.int64)
    for k in range(7):
        c = x & 0x1F
        x >>= 5
        more = np.where((c & 0x10) != 0, x != -1, x != 0)
        c = np.where(more, c | 0x20, c) + 48
        chars[:, k] = np.where(alive, c, 0)
        length += alive
        alive &= more
        if not alive.any():
            break
    flat = chars[np.arange(7)[None, :] < length[:, None]]                    # row-major: run order, then group order
    ends = np.cumsum(np.add.reduceat(length, offsets))
    starts = ends - np.add.reduceat(length, offsets)
    buf = flat.tobytes()
    return [buf[a:b] for a, b in zip(starts.tolist(), ends.tolist())]


def _prediction(rle_counts: bytes, box: np.ndarray, size: Tuple[int, int], score: float, label_id: int, image_id,
                label_id_to_category) -> dict:
    out = {
        "category_id": label_id,
        "segmentation": {"size": [int(size[0]), int(size[1])], "counts": rle_counts},   # pycocotools.mask.encode's dict
        "score": score,
        "image_id": image_id,
        "image_size": (int(size[0]), int(size[1])),
        "bbox": [float(v) for v in box],                                                 # masks_to_boxes(...).tolist()
    }
    if label_id_to_category is not None:
        out["pred_class"] = label_id_to_category[label_id]
    return out


def _nms_keep(cats: np.ndarray, scores: np.ndarray, inter: np.ndarray, nms_type: str, nms_threshold: float = 0.3,
              sigma: float = 0.5, floor: float = 0.001) -> List[Tuple[np.integer, int, np.floating]]:
    """Greedy per-category NMS of zutis.py:225-282 driven by pairwise intersection counts.

    ``inter[i,j] = |m_i & m_j|`` (diagonal = areas) comes from the popcount kernel; the IoU is
    ``inter / (area_i + area_j - inter + 1e-7)`` in float64 exactly as utils/iou.py:31-33 computes it.
    Returns (category, query index, score) in the reference's emission order.
    """
    assert nms_type in ["hard", "linear", "gaussian"]
    inter = np.asarray(inter).tolist()                       # Python ints: exact, and much cheaper than numpy scalars below
    area = [row[i] for i, row in enumerate(inter)]
    kept = []
    for cat in set(cats):
        if cat == 0:                      # background category
            continue
        cand = list(np.nonzero(cats == cat)[0])
        cand_scores = list(scores[cand])
        chosen = []
        while len(cand) > 0:
            order = np.argsort(np.array(cand_scores))
            cand = [cand[k] for k in order]
            cand_scores = [cand_scores[k] for k in order]
            top, top_score = cand[-1], cand_scores[-1]
            chosen.append((top, top_score))
            survivors, survivor_scores = [], []
            for i, s in zip(cand[:-1], cand_scores[:-1]):
                both = inter[i][top]
                iou = both / ((area[i] + area[top] - both) + 1e-7)            # float64, as utils/iou.py:31-33
                if nms_type == "hard":
                    weight = 0 if iou > nms_threshold else 1
                elif nms_type == "linear":
                    weight = (1 - np.float64(iou)) if iou > nms_threshold else 1
                else:
                    weight = np.exp(-(np.float64(iou) * iou) / sigma)
                s = s * weight
                if s > floor:
                    survivors.append(i); survivor_scores.append(s)
            cand, cand_scores = survivors, survivor_scores
        for i, s in chosen:
            if area[i] == 0:
                continue
            kept.append((cat, int(i), s))
    return kept


def _ordered_picks(cats: np.ndarray, scores: np.ndarray, pick_rank: np.ndarray, areas: np.ndarray):
    """Device NMS result (ops.instance_nms_hard) in the reference's emission order: categories as ``set(cats)`` iterates
    them (zutis.py:232), and inside a category the masks in the order they were picked; empty masks are skipped (:280)."""
    kept = []
    chosen = np.nonzero(pick_rank >= 0)[0]
    for cat in set(cats):
        if cat == 0:
            continue
        mine = chosen[cats[chosen] == cat]
        for i in mine[np.argsort(pick_rank[mine], kind="stable")]:
            if areas[i] == 0:
                continue
            kept.append((cat, int(i), scores[i]))
    return kept


@torch.no_grad()
def predict(
        self,
        dict_outputs: dict,
        mask_type: str,
        threshold: float = 0.5,  # threshold for binarising an instance mask
        image_ids: Optional[List[int]] = None,  # for COCO-style format
        size: Optional[Tuple[int, int]] = None,  # (H, W) format
        label_id_to_category: Optional[Dict[int, str]] = None,  # for instance segmentation
        new_label_id_to_old_label_id: Optional[Dict[int, int]] = None,  # for coco labels
        temperature: float = 5,
        nms_type: str = "hard",
        return_logits: bool = False,
        precision: Optional[str] = None,
):
    """Same contract as ``ZUTIS.predict`` (networks/zutis.py:340-470); ``precision`` is additive."""
    assert mask_type in ["semantic", "instance"]
    text = self.text_embeddings
    if mask_type == "semantic":
        tokens = dict_outputs["patch_tokens"]                      # b x h x w x n_dims
        ws = self.__dict__.setdefault("_zutis_b200_decode_ws", ops.DecodeWorkspace())   # run counter of the cell decode kernel
        lowres = ops.contraction(text.to(tokens.device), tokens, precision=precision,
                                 a_cache=self.__dict__.setdefault("_zutis_b200_text_cache", {}))       # b x n x h x w
        if return_logits:
            # the one mode in which full-resolution logits are materialised, on request (zutis.py:369-370)
            return ops.upsample_bilinear(lowres, size) if size is not None else lowres
        labels = ops.decode_score(lowres, size, workspace=ws)      # int16 on the device
        return labels.cpu().numpy().astype(np.int64)

    mask_proposals: torch.Tensor = dict_outputs["mask_proposals"]
    if len(mask_proposals.shape) == 5:
        mask_proposals = mask_proposals[:, -1, ...]                # last decoder layer only (zutis.py:379-382)
    lo, hi = torch.aminmax(mask_proposals)
    assert 0 <= lo <= 1
    assert 0 <= hi <= 1
    tokens = dict_outputs["patch_tokens"]
    B, Q, h, w = mask_proposals.shape

    sizes, psum, mean_tokens = ops.instance_lowres_stats(mask_proposals, tokens, threshold)
    cat_dev, prob_dev = ops.instance_categories(mean_tokens, text.to(tokens.device), temperature)
    bits, _ = ops.decode_threshold(mask_proposals, size, threshold, want_areas=False)      # [B,Q,H,words]
    H, W = ops.size_pair(size) if size is not None else (h, w)

    # (sum of in-mask p) / (size + 1e-7) * max category probability, in fp32 like the reference (:396, :420)
    sizes_h = sizes.cpu().numpy()
    confidence = (psum.cpu().numpy() / (sizes_h.astype(np.float32) + np.float32(1e-7))) * prob_dev.cpu().numpy()
    category_ids = cat_dev.cpu().numpy().astype(np.int64)

    if image_ids is None:
        image_ids = [0 for _ in range(B)]
    # pairwise intersection counts of every image, enqueued back to back; they stay on the device unless a host replay needs them
    inter_dev = torch.stack([ops.pairwise_mask_intersections(bits[b]) for b in range(B)])
    n_images = min(B, len(image_ids))                               # the reference zips range(B) with image_ids
    kept: List[Tuple[int, int, int, float]] = []                   # (image, query, category, score), reference order
    if nms_type == "hard":
        # greedy per-category suppression on the device; only areas, pick order and the tie flags come back
        pick_dev, tie_dev = ops.instance_nms_hard(inter_dev, cat_dev.to(torch.int32), torch.from_numpy(confidence).to(inter_dev.device))
        areas_all = torch.diagonal(inter_dev, dim1=1, dim2=2).cpu().numpy()
        pick_all, tie_all = pick_dev.cpu().numpy(), tie_dev.cpu().numpy()
        for b in range(n_images):
            cats_b, conf_b = category_ids[b], confidence[b]
            if tie_all[b]:
                # equal scores inside a category: the reference's pick follows numpy's argsort, replay its loop
                keep = _nms_keep(cats_b, conf_b, inter_dev[b].cpu().numpy(), nms_type)
            else:
                keep = _ordered_picks(cats_b, conf_b, pick_all[b], areas_all[b])
            kept.extend((b, q, c, s) for c, q, s in keep)
    else:
        inter_all = inter_dev.cpu().numpy()
        for b in range(n_images):
            cats_b, conf_b = category_ids[b], confidence[b]
            if nms_type is None:
                areas = inter_all[b].diagonal()
                keep = [(c, q, s) for q, (s, c) in enumerate(zip(conf_b, cats_b)) if areas[q] != 0 and c != 0]
            else:
                keep = _nms_keep(cats_b, conf_b, inter_all[b], nms_type)
            kept.extend((b, q, c, s) for c, q, s in keep)
    predictions: List[dict] = list()
    if not kept:
        return predictions
    # run lengths + boxes of the kept masks straight from the bit-packed device masks (no boolean mask leaves the GPU)
    ids = torch.as_tensor([b * Q + q for b, q, _, _ in kept], dtype=torch.int32)
    strings, boxes = ops.mask_rle_strings(bits, W, mask_ids=ids)
    for (b, q, c, s), counts, box in zip(kept, strings, boxes):
        label_id = new_label_id_to_old_label_id[c.item()] if new_label_id_to_old_label_id is not None else c.item()
        predictions.append(_prediction(counts, box, (H, W), s.item(), label_id, image_ids[b], label_id_to_category))
    return predictions


def get_mask_proposals(self, queries: torch.Tensor, patch_tokens: torch.Tensor, return_binary_masks: bool = True,
                       precision: Optional[str] = None):
    """Same contract as ``ZUTIS.get_mask_proposals`` (networks/zutis.py:177-209).

    queries [B,Q,C] or [B,L,Q,C]; patch_tokens [B,h,w,C] -> sigmoid(queries . tokens) of shape
    [B,(L,)Q,h,w]; the sigmoid is fused into the contraction kernel's epilogue.  With
    ``return_binary_masks`` (no caller in the reference) the raw products and one-hot argmax masks
    are returned instead, as the reference does.
    """
    if len(queries.shape) not in (3, 4):
        raise ValueError(f"{len(queries.shape)} not in [3, 4]")
    B = queries.shape[0]
    h, w = patch_tokens.shape[1:3]
    lead = tuple(queries.shape[1:-1])                                 # (Q,) or (L,Q)
    flat = queries.reshape(B, -1, queries.shape[-1])
    if return_binary_masks:
        raw = ops.contraction(flat, patch_tokens, precision=precision, pixel_major=False)
        raw = raw.view(B, *lead, h, w)
        n_queries = lead[-1]
        per_group = raw.reshape(-1, n_queries, h, w)
        winners = ops.decode_score(per_group, None).view(B, *lead[:-1], h, w)
        ids = torch.arange(n_queries, device=raw.device, dtype=winners.dtype).view(*([1] * len(lead[:-1])), n_queries, 1, 1)
        one_hot = winners.unsqueeze(-3) == ids.unsqueeze(0)
        return raw, one_hot
    out = ops.contraction(flat, patch_tokens, precision=precision, sigmoid=True, pixel_major=False)
    return out.view(B, *lead, h, w)


def image_to_text_space(self, patch_tokens: torch.Tensor, proj: torch.Tensor, channel_last: bool, layer_norm: bool = True,
                        precision: Optional[str] = None) -> torch.Tensor:
    """Same contract as ``ZUTIS.image_to_text_space`` (networks/zutis.py:301-331) for the branch ``forward`` uses with
    a ViT encoder: ``channel_last=True`` -- projection ``einsum("bhwn,nc->bhwc")`` on the contraction kernel
    (A = proj^T, pixel-major output = the [B,h,w,C] tensor itself), joint layer norm over (h,w,c) and per-pixel L2
    normalisation in one fused pass.  SURVEY section 8(f) N1: the step right before the semantic contraction."""
    if not channel_last or "RN" in getattr(self, "clip_arch", "ViT"):
        raise NotImplementedError("zutis_b200.image_to_text_space covers the ViT / channel_last=True branch (zutis.py:319-322)")
    cache = self.__dict__.setdefault("_zutis_b200_proj_cache", {})
    key = (proj.data_ptr(), proj._version, tuple(proj.shape))
    if cache.get("key") != key:
        cache.clear(); cache["key"] = key
        cache["projT"] = proj.detach().float().t().contiguous()            # [C, N]: rows are K-contiguous
        cache["ws"] = {}
    y = ops.contraction(cache["projT"], patch_tokens.detach(), precision=precision, a_cache=cache["ws"])   # view [B,C,h,w] over [B,h,w,Cp]
    B, Cc, h, w = y.shape
    buf = y.permute(0, 2, 3, 1)                                           # [B,h,w,C]
    if not buf.is_contiguous():                                           # C not a multiple of 4: drop the padding columns
        buf = buf.contiguous()
    return ops.image_layernorm_l2norm_(buf, layer_norm=layer_norm)


def decode_and_score(text: torch.Tensor, patch_tokens: torch.Tensor, label_trues, size, meter: RunningScore,
                     want_labels: bool = False, precision: Optional[str] = None):
    """Fused semantic evaluation step: contraction -> (upsample, argmax, confusion counts) on the device.

    Equivalent to ``meter.update(label_trues, predict(..., "semantic", size=size))``
    (trainer.py:331-347) without materialising logits or labels on the host.
    """
    ws = meter.__dict__.setdefault("_decode_ws", ops.DecodeWorkspace())
    cache = meter.__dict__.setdefault("_text_cache", {})
    lowres = ops.contraction(text, patch_tokens, precision=precision, a_cache=cache)
    return meter.update_from_logits(lowres, label_trues, size=size, want_labels=want_labels, workspace=ws)


class StreamingScorer:
    """The evaluation loop of trainer.py:331-347 as a two-stream pipeline on the device.

    ``submit`` runs the contraction on the caller's stream and the fused decode + scoring on a stream of its own, so the
    caller's stream is free for the next batch (backbone forward, next contraction) while this one is decoded.  Both
    kernels fill every SM, so what overlaps is one kernel's launch ramp and tail with the other's body: measured 8 % per
    step on cfg2 (bench.py, ``single_stream`` against the headline).  ``join`` / ``get_scores`` order the caller's
    stream after all scoring.  Results are identical to calling ``decode_and_score`` batch by batch."""

    def __init__(self, text_embeddings: torch.Tensor, meter: RunningScore, size, precision: Optional[str] = None):
        if not text_embeddings.is_cuda:
            raise TypeError("text_embeddings must live on a CUDA device (zutis_b200 has no CPU path)")
        self.text = text_embeddings.detach().float().contiguous()
        self.meter, self.size, self.precision = meter, size, precision
        self._stream = torch.cuda.Stream(device=self.text.device)
        self._cache: dict = {}
        self._ws = ops.DecodeWorkspace()

    def submit(self, patch_tokens: torch.Tensor, label_trues) -> None:
        main = torch.cuda.current_stream(self.text.device)
        lowres = ops.contraction(self.text, patch_tokens, precision=self.precision, a_cache=self._cache)
        gt = self.meter._as_device_labels(label_trues)
        self._stream.wait_event(main.record_event())
        with torch.cuda.stream(self._stream):
            self.meter.update_from_logits(lowres, gt, size=self.size, want_labels=False, workspace=self._ws)
        # both tensors were allocated on the caller's stream and are last read on the scoring stream
        lowres.record_stream(self._stream)
        gt.record_stream(self._stream)

    def join(self) -> None:
        torch.cuda.current_stream(self.text.device).wait_stream(self._stream)

    def get_scores(self):
        self.join()
        return self.meter.get_scores()


class ZutisDecoder:
    """The decode half of the model as a standalone object (only ``text_embeddings`` is needed)."""

    def __init__(self, text_embeddings: torch.Tensor):
        if not text_embeddings.is_cuda:
            raise TypeError("text_embeddings must live on a CUDA device (zutis_b200 has no CPU path)")
        self.text_embeddings = text_embeddings.detach().float().contiguous()

    predict = predict
    get_mask_proposals = get_mask_proposals
    image_to_text_space = image_to_text_space

    def decode_and_score(self, patch_tokens, label_trues, size, meter: RunningScore, want_labels: bool = False,
                         precision: Optional[str] = None):
        return decode_and_score(self.text_embeddings, patch_tokens, label_trues, size, meter, want_labels, precision)


def _wants_autograd(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def install(zutis_cls, inference_forward_ops: bool = False) -> None:
    """Bind the B200 decode path onto the reference model class (``networks.zutis.ZUTIS``).

    By default only ``predict`` is replaced: it is the evaluation entry (``@torch.no_grad()`` in the reference,
    zutis.py:340) and nothing in ``forward`` calls it, so ``trainer.fit`` trains exactly as before.
    ``get_mask_proposals`` and ``image_to_text_space`` ARE called by ``forward`` (zutis.py:522-530), whose outputs feed
    the training criterion; the kernels behind this module have no backward.  With ``inference_forward_ops=True``
    they are bound as well, behind a guard: whenever autograd is recording and an input (or the projection) requires
    grad, the call goes to the reference's own method, untouched; the kernels only serve no-grad forwards
    (``trainer.evaluate``, ``coco20k_eval.py``)."""
    zutis_cls.predict = predict
    if not inference_forward_ops:
        return
    ref_proposals = zutis_cls.get_mask_proposals
    ref_text_space = zutis_cls.image_to_text_space

    def guarded_get_mask_proposals(self, queries, patch_tokens, return_binary_masks: bool = True):
        if _wants_autograd(queries, patch_tokens) or not patch_tokens.is_cuda:
            return ref_proposals(self, queries, patch_tokens, return_binary_masks)
        return get_mask_proposals(self, queries, patch_tokens, return_binary_masks)

    def guarded_image_to_text_space(self, patch_tokens, proj, channel_last, layer_norm: bool = True):
        if (_wants_autograd(patch_tokens, proj) or not patch_tokens.is_cuda or not channel_last
                or "RN" in getattr(self, "clip_arch", "ViT")):
            return ref_text_space(self, patch_tokens, proj, channel_last, layer_norm)
        return image_to_text_space(self, patch_tokens, proj, channel_last, layer_norm)

    guarded_get_mask_proposals.__wrapped__ = ref_proposals
    guarded_image_to_text_space.__wrapped__ = ref_text_space
    zutis_cls.get_mask_proposals = guarded_get_mask_proposals
    zutis_cls.image_to_text_space = guarded_image_to_text_space
