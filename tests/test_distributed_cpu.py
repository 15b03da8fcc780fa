"""World-size-2 gloo test of the multi-GPU host logic (runs on CPU): every rank scores its own
image shard into a private int64 confusion matrix, one all-reduce sums them, and the result is
bit-identical to the single-process matrix (integer sums do not depend on the GPU count)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_images, n_classes, out_dir):
    import torch.distributed as dist
    from oracle import oracle as O
    from zutis_b200.distributed import all_reduce_counts, scores_from_counts, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                       # every rank generates the same dataset ...
    gt = rng.integers(0, n_classes, (n_images, 17, 13)); gt[:, :2] = 255
    pred = rng.integers(0, n_classes, (n_images, 17, 13))
    a, b = shard_range(n_images, rank, world)            # ... and scores only its shard
    local = O.c_fast_hist(gt[a:b], pred[a:b], n_classes) if b > a else np.zeros((n_classes, n_classes), np.int64)
    counts = torch.from_numpy(local.copy())
    all_reduce_counts(counts)
    s, _ = scores_from_counts(counts)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), counts.numpy())
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.array([s["Mean IoU"], s["Pixel Acc"]]))
    dist.destroy_process_group()


def test_sharded_confusion_matrix_equals_single_process(tmp_path):
    from oracle import oracle as O
    world, n_images, n_classes = 2, 7, 11
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, n_classes, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    gt = rng.integers(0, n_classes, (n_images, 17, 13)); gt[:, :2] = 255
    pred = rng.integers(0, n_classes, (n_images, 17, 13))
    whole = O.c_fast_hist(gt, pred, n_classes)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), whole)
    s, _ = O.scores_from_hist(whole)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"s{r}.npy"), np.array([s["Mean IoU"], s["Pixel Acc"]]))
