"""Drop-in for the reference's ``utils/running_score.py`` with the counting done on the GPU.

Same surface as the reference class (utils/running_score.py:5-50): ``RunningScore(n_classes)``,
``update(label_trues, label_preds)``, ``get_scores() -> (dict, dict)``, ``reset()``, attributes
``n_classes`` and ``confusion_matrix`` (float64 ``[n,n]``, rows = ground truth, columns =
prediction).  Callers: trainer.py:126/314 (construction), :177/:347 (update), :178/:348 (get_scores).

Behind it the matrix lives on the device as int64; every update enqueues a histogram kernel
(zutis_score_labels, or the fused zutis_decode_score through ``update_from_logits``) into an int32
per-launch partial that the merge kernel folds into the int64 matrix.  Nothing is copied back
until ``get_scores()`` / ``confusion_matrix`` is asked for.  Host numpy inputs (what the reference
passes: trainer.py:320) are copied to the device and counted there -- there is no CPU counting
path.  Additive API: ``update_from_logits`` (fused decode+score), ``all_reduce`` (multi-GPU sum).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch

from . import ops

_INT_DTYPES = (torch.uint8, torch.int16, torch.int32, torch.int64)


class RunningScore(object):
    def __init__(self, n_classes: int, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("zutis_b200.RunningScore needs a CUDA device (sm_100); there is no CPU path")
        self.n_classes = int(n_classes)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n2 = self.n_classes * self.n_classes
        self._hist = torch.zeros(n2, dtype=torch.int64, device=self.device)        # persistent counts
        self._partial = torch.zeros(n2, dtype=torch.int32, device=self.device)     # per-launch partial
        self._pending = 0          # pixels counted into the partial since the last merge
        self._host_extra = None    # float64 matrix assigned by a caller through the attribute, if any
        # non-blocking read-out (get_scores_async / poll_scores): pinned host mirror of the matrix + completion event
        self._pinned = None
        self._copy_event = None
        self._last_scores = None

    # ------------------------------------------------------------------ device-side plumbing
    def _as_device_labels(self, x) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            t = x
        else:
            a = np.asarray(x)
            if a.dtype.kind not in "iub":
                raise TypeError(f"labels must be integer arrays, got {a.dtype}")
            if a.dtype == np.bool_:
                a = a.astype(np.uint8)
            elif a.dtype not in (np.uint8, np.int16, np.int32, np.int64):
                a = a.astype(np.int64)
            t = torch.from_numpy(np.ascontiguousarray(a))
        if t.dtype not in _INT_DTYPES:
            if t.dtype == torch.bool:
                t = t.to(torch.uint8)
            elif t.dtype in (torch.int8,):
                t = t.to(torch.int16)
            else:
                raise TypeError(f"labels must be integer tensors, got {t.dtype}")
        return t.to(self.device, non_blocking=True)

    def _note_pixels(self, n: int) -> None:
        self._pending += n
        if self._pending >= (1 << 30):          # keep the int32 partial far from overflow
            self._merge()

    def _merge(self) -> None:
        if self._pending:
            ops.hist_merge(self._partial, self._hist, clear=True)
            self._pending = 0

    # ------------------------------------------------------------------------- reference API
    def update(self, label_trues, label_preds) -> None:
        """Accumulate confusion counts (running_score.py:18-20).

        ``label_trues`` / ``label_preds`` are zipped over their first dimension exactly like the
        reference: a ``[B,H,W]`` array/tensor or a list of ``[H,W]`` arrays of different sizes
        are both legal.  A same-shaped pair of tensors/arrays is counted in one launch.
        """
        same_block = (hasattr(label_trues, "shape") and hasattr(label_preds, "shape")
                      and tuple(label_trues.shape) == tuple(label_preds.shape) and len(label_trues.shape) >= 1)
        if same_block:
            pairs = [(label_trues, label_preds)]
        else:
            pairs = list(zip(label_trues, label_preds))
        for lt, lp in pairs:
            t = self._as_device_labels(lt)
            p = self._as_device_labels(lp)
            if t.numel() != p.numel():
                raise ValueError(f"label_true and label_pred differ in size: {tuple(t.shape)} vs {tuple(p.shape)}")
            if t.numel() == 0:
                continue
            ops.score_labels(t, p, self._partial, self.n_classes)
            self._note_pixels(t.numel())

    def update_from_logits(self, lowres_logits: torch.Tensor, label_trues, size=None, want_labels: bool = False,
                           workspace=None):
        """Fused path: upsample + argmax + count in one kernel; full-resolution logits never exist.

        Equivalent to ``update(label_trues, argmax(interpolate(lowres_logits, size)))``
        (zutis.py:366-372 + running_score.py:18-20).  Returns int16 labels when asked.
        """
        gt = self._as_device_labels(label_trues)
        labels = ops.decode_score(lowres_logits, size, gt=gt, hist_partial=self._partial, n_classes=self.n_classes,
                                  want_labels=want_labels, workspace=workspace)
        self._note_pixels(gt.numel())
        return labels

    def all_reduce(self, group=None, nccl_comm: Optional[int] = None, peer=None) -> None:
        """Sum the per-GPU matrices (one int64 all-reduce of n^2 counts): over the torch.distributed process group; for
        hosts that own a raw NCCL communicator over ``nccl_comm`` (the ncclComm_t as an integer) through the C ABI; or over
        NVLink peer memory with a ``distributed.PeerReducer`` (one node, small matrices: a third of NCCL's latency)."""
        if peer is not None and self._hist.numel() <= peer.max_elements:
            if self._pending:                                 # merge and all-reduce in one launch
                if getattr(self, "_reduced", None) is None:
                    self._reduced = torch.empty_like(self._hist)
                peer.merge_all_reduce(self._hist, self._partial, self._reduced)
                self._hist, self._reduced = self._reduced, self._hist
                self._pending = 0
            else:
                peer.all_reduce(self._hist)
            return
        self._merge()
        if nccl_comm is not None:
            with torch.cuda.device(self.device):
                ops.F.call("zutis_allreduce_hist", self._hist.data_ptr(), self._hist.numel(), nccl_comm,
                       torch.cuda.current_stream(self.device).cuda_stream)
            return
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self._hist, op=dist.ReduceOp.SUM, group=group)

    def counts(self) -> torch.Tensor:
        """The int64 ``[n,n]`` device matrix (merged, not synchronised)."""
        self._merge()
        return self._hist.view(self.n_classes, self.n_classes)

    @property
    def confusion_matrix(self) -> np.ndarray:
        m = self.counts().cpu().numpy().astype(np.float64)
        if self._host_extra is not None:
            m = m + self._host_extra
        return m

    @confusion_matrix.setter
    def confusion_matrix(self, value) -> None:
        value = np.asarray(value, dtype=np.float64)
        if value.shape != (self.n_classes, self.n_classes):
            raise ValueError(f"confusion_matrix must be {(self.n_classes, self.n_classes)}, got {value.shape}")
        self._hist.zero_(); self._partial.zero_(); self._pending = 0
        self._host_extra = value.copy()

    # ------------------------------------------------------------------ non-blocking read-out (additive)
    def get_scores_async(self) -> None:
        """Enqueue merge + device-to-host copy of the matrix into pinned memory and return at once.

        For callers that refresh a progress bar every batch (trainer.py:178, :348): the copy rides the stream behind
        the batch's kernels, nothing waits for it, and ``poll_scores()`` hands out the scores of the most recent copy
        that has completed.  A copy still in flight is not overtaken: a second request before it lands is dropped."""
        if self._copy_event is not None and not self._copy_event.query():
            return
        self._harvest()
        self._merge()
        if self._pinned is None:
            self._pinned = torch.empty(self._hist.shape, dtype=torch.int64, pin_memory=True)
        self._pinned.copy_(self._hist, non_blocking=True)
        self._copy_event = torch.cuda.Event()
        self._copy_event.record(torch.cuda.current_stream(self.device))

    def _harvest(self) -> None:
        if self._copy_event is not None and self._copy_event.query():
            m = self._pinned.numpy().astype(np.float64).reshape(self.n_classes, self.n_classes)
            if self._host_extra is not None:
                m = m + self._host_extra
            self._last_scores = self._scores_of(m)
            self._copy_event = None

    def poll_scores(self) -> Optional[Tuple[Dict[str, float], Dict[int, float]]]:
        """Scores of the latest completed ``get_scores_async`` copy (None until one has landed).  Never blocks."""
        self._harvest()
        return self._last_scores

    def get_scores(self) -> Tuple[Dict[str, float], Dict[int, float]]:
        """Pixel / mean / frequency-weighted accuracy and mean IoU (running_score.py:22-47).

        float64 numpy on the n x n matrix, same operations in the same order as the reference, so
        equal counts give bit-identical scores (NaN for absent classes, skipped by nanmean).
        Synchronises (it returns numbers); see ``get_scores_async`` / ``poll_scores`` for per-batch polling.
        """
        return self._scores_of(self.confusion_matrix)

    @staticmethod
    def _scores_of(m: np.ndarray) -> Tuple[Dict[str, float], Dict[int, float]]:
        n_classes = m.shape[0]
        diag = np.diag(m)
        rows = m.sum(axis=1)
        cols = m.sum(axis=0)
        total = m.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                acc = diag.sum() / total
                acc_cls = np.nanmean(diag / rows)
                iu = diag / (rows + cols - diag)
                mean_iu = np.nanmean(iu)
                freq = rows / total
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
        return (
            {"Pixel Acc": acc, "Mean Acc": acc_cls, "FreqW Acc": fwavacc, "Mean IoU": mean_iu},
            dict(zip(range(n_classes), iu)),
        )

    def reset(self) -> None:
        self._hist.zero_()
        self._partial.zero_()
        self._pending = 0
        self._host_extra = None
        self._copy_event = None
        self._last_scores = None
