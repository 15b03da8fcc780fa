"""Multi-GPU plumbing for the decode+score path: one process per GPU, images sharded, one all-reduce.

The path has no data exchange between images: rank r of G decodes and scores the contiguous image
range ``shard_range(n_images, r, G)`` into its private int64 confusion matrix, and the matrices are
summed once, when scores are requested (``RunningScore.all_reduce``: a single NCCL all-reduce of
Q*Q int64 counts over NVLink; 52 KB for Q=81, 6.8 MB for Q=920).  Integer addition makes the result
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
