"""The candidate-pruning rule of csrc/decode_pruned.cu, restated in numpy float32 and checked against the CPU oracle's
exact labels: on every cell, the label of every pixel must be among the categories the rule keeps, and a cell that
qualifies for the whole-cell shortcut must carry its champion everywhere.  No GPU needed: this pins the argument
(monotone rounding + margin) independently of the kernel that implements it."""
import numpy as np
import pytest

from oracle import oracle as O


def _cells(n_in, n_out):
    """first output index of every low-res cell along one axis (cell c = outputs whose first tap is c)"""
    i0, _, _, _ = O.c_axis_table(n_in, n_out)
    starts = np.searchsorted(i0, np.arange(n_in + 1), side="left")
    return starts


def _survivors(A, B, C, D, margin):
    """float32 restatement of the dominance test: keep q unless some corner champion k leads it at all four corners,
    by >= 0 when k < q and by >= margin when k >= q."""
    Q = A.shape[0]
    corners = (A, B, C, D)
    champs = []
    for X in corners:
        k = int(np.argmax(X))                       # first maximum
        if k not in champs:
            champs.append(k)
    keep = np.ones(Q, bool)
    q = np.arange(Q)
    for k in champs:
        lead = np.minimum.reduce([np.float32(X[k]) - X for X in corners]).astype(np.float32)   # RN(champion - q), fp32
        thr = np.where(k < q, np.float32(0), margin).astype(np.float32)
        keep &= ~(lead >= thr)
    return keep, champs


def _shortcut(A, B, C, D, margin):
    ks = {int(np.argmax(X)) for X in (A, B, C, D)}
    if len(ks) != 1:
        return None
    k = ks.pop()
    for X in (A, B, C, D):
        before = X[:k].max() if k > 0 else np.float32(-np.inf)
        if not (np.float32(X[k] - before) >= margin):
            return None
    return k


@pytest.mark.parametrize("case", ["smooth", "noise", "near_ties", "duplicates", "regions", "noninteger"])
def test_pruning_rule_never_drops_the_winner(case):
    rng = np.random.default_rng(len(case) * 13 + 5)
    Q, h, w, H, W = 24, 7, 9, 56, 72
    if case == "noninteger":
        H, W = 53, 70
    coarse = rng.standard_normal((Q, 3, 4)).astype(np.float32)
    yy = np.linspace(0, 2, h)[:, None]; xx = np.linspace(0, 3, w)[None, :]
    y0 = np.minimum(yy.astype(int), 1); x0 = np.minimum(xx.astype(int), 2)
    fy = (yy - y0).astype(np.float32); fx = (xx - x0).astype(np.float32)
    smooth = ((1 - fy) * (1 - fx) * coarse[:, y0, x0] + (1 - fy) * fx * coarse[:, y0, x0 + 1] +
              fy * (1 - fx) * coarse[:, y0 + 1, x0] + fy * fx * coarse[:, y0 + 1, x0 + 1]).astype(np.float32)
    if case in ("smooth", "noninteger"):
        lo = smooth
    elif case == "noise":
        lo = rng.standard_normal((Q, h, w)).astype(np.float32)
    elif case == "near_ties":
        base = np.tile(smooth[:6], (4, 1, 1))
        scale = np.abs(base).max() * np.float32(2.0) ** -rng.integers(18, 25, base.shape).astype(np.float32)
        lo = (base + rng.integers(-1, 2, base.shape).astype(np.float32) * scale).astype(np.float32)
    elif case == "duplicates":
        lo = smooth.copy(); lo[5] = lo[17]; lo[20] = lo[2]; lo[:, :3] = np.round(lo[:, :3] * 2) / 2
    else:
        lo = (0.05 * rng.standard_normal((Q, h, w))).astype(np.float32)
        region = rng.integers(0, Q, (2, 3)).repeat(4, 0).repeat(3, 1)[:h, :w]
        lo[region, np.arange(h)[:, None], np.arange(w)[None, :]] += 1.0
    labels = O.c_decode_semantic(lo[None], (H, W))[0]
    ys, xs = _cells(h, H), _cells(w, W)
    margin = np.float32(max(np.float32(np.abs(lo).max()) * np.float32(2.0 ** -20), np.float32(1e-37)))
    kept_total, shortcuts = 0, 0
    for cy in range(h):
        for cx in range(w):
            cy1, cx1 = min(cy + 1, h - 1), min(cx + 1, w - 1)
            A, B, C, D = lo[:, cy, cx], lo[:, cy, cx1], lo[:, cy1, cx], lo[:, cy1, cx1]
            cell = labels[ys[cy]:ys[cy + 1], xs[cx]:xs[cx + 1]]
            if cell.size == 0:
                continue
            keep, _ = _survivors(A, B, C, D, margin)
            assert keep[np.unique(cell)].all(), (case, cy, cx, np.unique(cell), np.flatnonzero(keep))
            kept_total += int(keep.sum())
            k = _shortcut(A, B, C, D, margin)
            if k is not None:
                shortcuts += 1
                assert (cell == k).all(), (case, cy, cx, k, np.unique(cell))
    assert kept_total >= h * w                                          # at least the winner survives everywhere
    if case == "regions":
        assert shortcuts > 0                                            # the shortcut is exercised
    if case == "smooth":
        assert kept_total < 0.5 * Q * h * w                             # and the rule actually prunes
