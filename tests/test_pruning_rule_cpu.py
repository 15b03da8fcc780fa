"""The candidate-pruning rule of csrc/decode_cells.cu, restated in numpy float32 and checked against the CPU oracle's
exact labels: on every cell, the label of every pixel must be among the categories the rule keeps, a cell whose list
holds a single category must carry it everywhere, and ANY category may serve as dominator without losing a winner.
No GPU needed: this pins the argument (monotone rounding + margin) independently of the kernel that implements it."""
import numpy as np
import pytest

from oracle import oracle as O


def _cells(n_in, n_out):
    """first output index of every low-res cell along one axis (cell c = outputs whose first tap is c)"""
    i0, _, _, _ = O.c_axis_table(n_in, n_out)
    starts = np.searchsorted(i0, np.arange(n_in + 1), side="left")
    return starts


def _dominator(A, B, C, D):
    """k* = argmax_q min_corner L_q (first such q): the category with the best guaranteed value in the cell"""
    return int(np.argmax(np.minimum.reduce([A, B, C, D])))


def _survivors(A, B, C, D, k):
    """float32 restatement of the test: keep q unless dominator k leads it at all four corners by
    m = 2^-20 * max_c |K_c| (never 0), the differences rounded to fp32 like the kernel's fma(q, -1, K)."""
    corners = (A, B, C, D)
    K = [np.float32(X[k]) for X in corners]
    margin = np.float32(max(np.float32(max(abs(v) for v in K)) * np.float32(2.0 ** -20), np.float32(1e-37)))
    lead = np.minimum.reduce([(Kc - X).astype(np.float32) for Kc, X in zip(K, corners)])
    return ~(lead >= margin)


def _survivors_virtual(A, B, C, D):
    """The second attempt of the kernel (cells whose first list overflows): dominator V = average of the four corner
    champions, V_c = 0.25 * ((L_kA[c] + L_kB[c]) + (L_kC[c] + L_kD[c])) in fp32, margin 2^-19 * max|those 16 taps|."""
    corners = (A, B, C, D)
    champs = [int(np.argmax(X)) for X in corners]
    V, big = [], np.float32(0)
    for X in corners:
        t = [np.float32(X[k]) for k in champs]
        V.append(np.float32(0.25) * np.float32(np.float32(t[0] + t[1]) + np.float32(t[2] + t[3])))
        big = max(big, np.float32(max(abs(v) for v in t)))
    margin = np.float32(max(big * np.float32(2.0 ** -19), np.float32(1e-37)))
    lead = np.minimum.reduce([(Vc - X).astype(np.float32) for Vc, X in zip(V, corners)])
    return ~(lead >= margin)


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
    kept_total, singles, virtual_total = 0, 0, 0
    for cy in range(h):
        for cx in range(w):
            cy1, cx1 = min(cy + 1, h - 1), min(cx + 1, w - 1)
            A, B, C, D = lo[:, cy, cx], lo[:, cy, cx1], lo[:, cy1, cx], lo[:, cy1, cx1]
            cell = labels[ys[cy]:ys[cy + 1], xs[cx]:xs[cx + 1]]
            if cell.size == 0:
                continue
            k = _dominator(A, B, C, D)
            keep = _survivors(A, B, C, D, k)
            assert keep[k]                                              # the dominator never drops itself
            assert keep[np.unique(cell)].all(), (case, cy, cx, np.unique(cell), np.flatnonzero(keep))
            kept_total += int(keep.sum())
            if keep.sum() == 1:
                singles += 1
                assert (cell == k).all(), (case, cy, cx, k, np.unique(cell))
            # the virtual dominator (average of the corner champions) keeps every winner too, and never empties the list
            keep_v = _survivors_virtual(A, B, C, D)
            assert keep_v.any() and keep_v[np.unique(cell)].all(), (case, cy, cx, np.unique(cell), np.flatnonzero(keep_v))
            if keep_v.sum() == 1:
                assert (cell == np.flatnonzero(keep_v)[0]).all()
            virtual_total += int(keep_v.sum())
            # which category dominates only changes how much is pruned, never whether a winner survives
            for other in (0, Q - 1, int(rng.integers(0, Q))):
                assert _survivors(A, B, C, D, other)[np.unique(cell)].all(), (case, cy, cx, other)
    assert kept_total >= h * w                                          # at least the winner survives everywhere
    if case == "regions":
        assert singles > 0                                              # single-survivor cells are exercised
        assert virtual_total < kept_total                               # across region borders the virtual dominator prunes more
    if case == "smooth":
        assert kept_total < 0.5 * Q * h * w                             # and the rule actually prunes
