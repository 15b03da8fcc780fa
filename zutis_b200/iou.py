"""Drop-in for the reference's ``utils/iou.py`` (binary-mask IoU), computed on the GPU.

``compute_iou(pred_mask, gt_mask, threshold=None, eps=1e-7)`` keeps the reference signature and
return types (utils/iou.py:6-38): numpy inputs give a numpy float64 scalar, torch inputs give a
0-d CPU tensor.  Pixels whose ground truth lies outside [0, 1] are dropped, an optional strict
threshold binarises the prediction, and ``iou = |p & g| / (|p | g| + eps)``.

The counting is a 2 x 2 confusion histogram (ground truth x prediction), so it runs through the
same scoring kernel as RunningScore (zutis_score_labels): intersection = hist[1,1],
union = hist[0,1] + hist[1,0] + hist[1,1].  Inside NMS the pairwise form on bit-packed masks is
used instead (ops.pairwise_mask_intersections).
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch

from . import ops


def _binary_u8(x, threshold: Optional[float], device) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    t = t.to(device, non_blocking=True)
    if threshold is not None:
        return (t > threshold).to(torch.uint8)
    if t.dtype == torch.bool:
        return t.to(torch.uint8)
    return (t != 0).to(torch.uint8)


def compute_iou(pred_mask: Union[np.ndarray, torch.Tensor], gt_mask: Union[np.ndarray, torch.Tensor],
                threshold: Optional[float] = None, eps: float = 1e-7) -> Union[np.ndarray, torch.Tensor]:
    assert pred_mask.shape == gt_mask.shape, f"{pred_mask.shape} != {gt_mask.shape}"
    assert len(pred_mask.shape) == 2, ValueError(f"{len(pred_mask.shape)} != 2")
    assert len(gt_mask.shape) == 2, ValueError(f"{len(gt_mask.shape)} != 2")
    if not torch.cuda.is_available():
        raise RuntimeError("zutis_b200.compute_iou needs a CUDA device (sm_100); there is no CPU path")
    is_torch = isinstance(pred_mask, torch.Tensor)
    device = pred_mask.device if (is_torch and pred_mask.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    g = gt_mask if isinstance(gt_mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gt_mask))
    g = g.to(device, non_blocking=True)
    # ground truth outside [0,1] is ignored (utils/iou.py:23): encode it as label 255 so the scoring
    # kernel drops it through its 0 <= gt < n rule; otherwise the label is "gt is truthy"
    gl = torch.where((g >= 0) & (g <= 1), (g != 0).to(torch.uint8), torch.full_like(g, 255, dtype=torch.uint8))
    p = _binary_u8(pred_mask, threshold, device)
    partial = torch.zeros(4, dtype=torch.int32, device=device)
    ops.score_labels(gl, p, partial, 2)
    h = partial.cpu().numpy().astype(np.int64)
    intersection = h[3]
    union = h[1] + h[2] + h[3]
    if is_torch:
        return (torch.tensor(intersection) / (torch.tensor(union) + eps))
    return intersection / (union + eps)
