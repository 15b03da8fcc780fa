"""Multi-GPU plumbing for the decode+score path: one process per GPU, images sharded, one all-reduce.

The path has no data exchange between images: rank r of G decodes and scores the contiguous image
range ``shard_range(n_images, r, G)`` into its private int64 confusion matrix, and the matrices are
summed once, when scores are requested (``RunningScore.all_reduce``: a single all-reduce of Q*Q int64
counts over NVLink -- NCCL, or for small matrices ``PeerReducer``, the library's own kernel over peer
memory; 52 KB for Q=81, 6.8 MB for Q=920).  Integer addition makes the result
bit-identical for any GPU count, which is what tests/test_distributed_cpu.py checks with gloo.
The reference itself is single-process, single-GPU (main.py:54); this module is new.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``n_items % world_size`` ranks get one extra item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_distributed(backend: str = "nccl"):
    """Join the torchrun rendezvous described by RANK / WORLD_SIZE / MASTER_* (no-op when single)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def all_reduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of an int64 count tensor over the group (works for nccl and gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class PeerReducer:
    """All-reduce of small int64 count matrices over NVLink peer memory (csrc/p2p_reduce.cu) for the ranks of ONE node.

    Built collectively: every rank of ``group`` constructs it at the same point (the IPC handles are exchanged with
    ``all_gather_object``).  ``all_reduce(counts)`` then sums in place on the current stream with one kernel launch per rank;
    every rank must call it the same number of times.  Matrices larger than ``max_elements`` belong to NCCL."""

    def __init__(self, max_elements: int = 128 * 128, group=None, device=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _ffi as F
        self._F = F
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.max_elements = int(max_elements)
        handle = (C.c_ubyte * 64)()
        ctx = C.c_int(-1)
        with torch.cuda.device(self.device):
            F.call("zutis_p2p_create", self.world, self.rank, self.max_elements, C.addressof(handle), C.addressof(ctx))
            self.ctx = ctx.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
            F.call("zutis_p2p_connect", self.ctx, C.addressof(blob))
        dist.barrier(group=group)                       # every rank has mapped every block before anyone signals

    def all_reduce(self, counts: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        if counts.dtype != torch.int64 or not counts.is_cuda or not counts.is_contiguous() or counts.numel() > self.max_elements:
            raise ValueError("PeerReducer.all_reduce wants a contiguous int64 CUDA tensor of at most max_elements elements")
        out = counts if out is None else out
        with torch.cuda.device(self.device):
            self._F.call("zutis_allreduce_hist_p2p", self.ctx, counts.data_ptr(), counts.numel(), out.data_ptr(),
                         torch.cuda.current_stream(self.device).cuda_stream)
        return out

    def merge_all_reduce(self, counts: torch.Tensor, partial: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """counts += partial; partial = 0; out = sum over ranks of counts -- the merge kernel and the all-reduce in one launch."""
        if counts.numel() > self.max_elements or out.data_ptr() == counts.data_ptr():
            raise ValueError("PeerReducer.merge_all_reduce: matrix too large for this reducer, or out aliases counts")
        with torch.cuda.device(self.device):
            self._F.call("zutis_merge_allreduce_hist_p2p", self.ctx, counts.data_ptr(), partial.data_ptr(), counts.numel(), out.data_ptr(),
                         torch.cuda.current_stream(self.device).cuda_stream)
        return out

    def close(self) -> None:
        if getattr(self, "ctx", -1) >= 0:
            with torch.cuda.device(self.device):
                self._F.call("zutis_p2p_destroy", self.ctx)
            self.ctx = -1


def scores_from_counts(counts) -> tuple:
    """get_scores (utils/running_score.py:22-47) on an [n,n] count matrix, float64, reference op order."""
    m = np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts, dtype=np.float64)
    diag = np.diag(m); rows = m.sum(axis=1); cols = m.sum(axis=0); total = m.sum()
    import warnings
    with np.errstate(divide="ignore", invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        acc = diag.sum() / total
        acc_cls = np.nanmean(diag / rows)
        iu = diag / (rows + cols - diag)
        mean_iu = np.nanmean(iu)
        freq = rows / total
    fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return ({"Pixel Acc": acc, "Mean Acc": acc_cls, "FreqW Acc": fwavacc, "Mean IoU": mean_iu}, dict(zip(range(m.shape[0]), iu)))
