"""Host-side COCO string compression of device run lengths (zutis_b200.decode._rle_strings) against the scalar
restatement of cocoapi's rleToString in oracle/ (no GPU needed)."""
import numpy as np

from oracle import oracle as O
from zutis_b200.decode import _rle_strings


def test_rle_strings_match_scalar_restatement():
    rng = np.random.default_rng(11)
    masks = []
    for _ in range(40):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        masks.append(rng.random((H, W)) < rng.random())
    masks.append(np.zeros((480, 640), bool))
    masks.append(np.ones((480, 640), bool))
    stripes = np.zeros((480, 640), bool); stripes[::2] = True             # 153600 transitions + long final runs
    masks.append(stripes)
    counts = [O.rle_counts_numpy(m) for m in masks]
    n_runs = np.array([len(c) for c in counts], np.int64)
    runs = np.concatenate(counts).astype(np.uint32)
    got = _rle_strings(n_runs, runs)
    assert len(got) == len(masks)
    for s, c in zip(got, counts):
        assert s == O.rle_to_string(c)
    # large values and negative differences
    c = [0, 5, 2 ** 30, 1, 7, 2 ** 30 + 3, 1, 2 ** 31 - 1, 2]
    assert _rle_strings(np.array([len(c)]), np.array(c, np.uint32))[0] == O.rle_to_string(c)
    assert _rle_strings(np.zeros(0, np.int64), np.zeros(0, np.uint32)) == []
