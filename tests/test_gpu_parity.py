"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI,
against the CPU oracle and the golden fixtures captured from the reference.

Bars (BASELINE.json north_star): integer work (labels for identical logits, confusion matrices)
bit-exact; low-res logits within 1e-5 of max|logit| (fp32-grade paths) or 2e-2 (bf16); label
agreement with the reference path >= 99.99 % with disagreements only at fp ties; mIoU within 1e-4.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = ["int8x", "nonint", "x16", "same", "down", "wideq", "ties", "nan"]


@pytest.fixture(scope="module")
def zb(cuda_device):
    import zutis_b200
    from zutis_b200 import _ffi, ops
    _ffi.check(_ffi.lib().zutis_device_check(0))
    return zutis_b200


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def pixel_major(logits_bqhw: torch.Tensor) -> torch.Tensor:
    """[B,Q,h,w] view over [B,h,w,Qp] memory (what the contraction kernel emits)."""
    B, Q, h, w = logits_bqhw.shape
    Qp = (Q + 3) & ~3
    buf = torch.zeros(B, h, w, Qp, device=logits_bqhw.device)
    buf[..., :Q] = logits_bqhw.permute(0, 2, 3, 1)
    return buf[..., :Q].permute(0, 3, 1, 2)


def tie_gap_ok(full_ref: np.ndarray, a: np.ndarray, b: np.ndarray, ulps: float = 16.0) -> bool:
    """Every pixel where label maps a and b differ must be a near-tie of the reference's full-res logits."""
    bad = np.argwhere(a != b)
    scale = np.abs(full_ref).max() * np.finfo(np.float32).eps
    for (bi, y, x) in bad:
        col = full_ref[bi, :, y, x]
        if abs(float(col[a[bi, y, x]]) - float(col[b[bi, y, x]])) > ulps * scale:
            return False
    return True


# ------------------------------------------------------------------------------ contraction
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("auto", 1e-5), ("tf32x3", 1e-5), ("tf32", 2e-2)])
def test_contraction_model_cfg1(zb, golden, precision, tol):
    g = golden("model_cfg1")
    out = zb.ops.contraction(dev(g["text"]), dev(g["tokens"]), precision=precision)
    torch.cuda.synchronize()
    ref = g["lowres_logits"]
    assert tuple(out.shape) == ref.shape and out.stride(1) == 1          # pixel-major memory
    assert np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max() <= tol
    out2 = zb.ops.contraction(dev(g["text"]), dev(g["tokens"]), precision=precision, pixel_major=False)
    assert out2.is_contiguous() and torch.equal(out2, out.contiguous())
    if precision == "tf32":
        # the single-pass mode really is reduced precision (it must NOT be what feeds labels)
        assert np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max() > 1e-5


def test_tcgen05_path_is_taken_and_matches_fp32_kernel(zb):
    """K % 32 == 0 shapes run on the tensor-core kernel; compare with the FFMA kernel and float64."""
    gen = torch.Generator().manual_seed(8)
    for (B, M, h, w, K) in [(3, 81, 40, 40, 512), (2, 920, 9, 11, 512), (2, 100, 15, 20, 768), (1, 16, 3, 5, 32), (2, 257, 12, 12, 64)]:
        text = torch.nn.functional.normalize(torch.randn(M, K, generator=gen), dim=-1).cuda()
        tok = torch.nn.functional.normalize(torch.randn(B, h, w, K, generator=gen), dim=-1).cuda()
        ref = torch.einsum("nc,bhwc->bnhw", text.double(), tok.double())
        scale = float(ref.abs().max())
        for pm in (True, False):
            tc = zb.ops.contraction(text, tok, precision="tf32x3", pixel_major=pm)
            ff = zb.ops.contraction(text, tok, precision="fp32", pixel_major=pm)
            assert float((tc.double() - ref).abs().max()) / scale <= 1e-5, (B, M, h, w, K, pm)
            assert float((ff.double() - ref).abs().max()) / scale <= 2e-6
        one = zb.ops.contraction(text, tok, precision="tf32")
        assert 1e-5 < float((one.double() - ref).abs().max()) / scale <= 2e-2
    # more tiles than CTAs: every CTA runs several tiles back to back with all rings full (24*13 = 312 tiles on
    # 148 SMs).  A slot-reuse race in the pipeline only shows up here, so check EVERY image.
    text = torch.nn.functional.normalize(torch.randn(81, 512, generator=gen), dim=-1).cuda()
    tok = torch.nn.functional.normalize(torch.randn(24, 40, 40, 512, generator=gen), dim=-1).cuda()
    ref = torch.einsum("nc,bhwc->bnhw", text.double(), tok.double())
    for _ in range(3):
        for prec in ("tf32x3", "tf32"):
            got = zb.ops.contraction(text, tok, precision=prec)
            per_image = (got.double() - ref).abs().amax(dim=(1, 2, 3)) / float(ref.abs().max())
            assert float(per_image.max()) <= (1e-5 if prec == "tf32x3" else 2e-2), per_image.tolist()
    wide = torch.nn.functional.normalize(torch.randn(920, 512, generator=gen), dim=-1).cuda()
    tokw = torch.nn.functional.normalize(torch.randn(6, 56, 56, 512, generator=gen), dim=-1).cuda()      # 6*25*4 = 600 tiles
    refw = torch.einsum("nc,bhwc->bnhw", wide.double(), tokw.double())
    gotw = zb.ops.contraction(wide, tokw, precision="tf32x3")
    assert float((gotw.double() - refw).abs().amax() / refw.abs().max()) <= 1e-5
    # per-image A operand (queries) with the fused sigmoid epilogue
    q = torch.nn.functional.normalize(torch.randn(3, 100, 768, generator=gen), dim=-1).cuda()
    feats = torch.randn(3, 15, 20, 768, generator=gen).cuda()
    ref = torch.sigmoid(torch.einsum("bqc,bhwc->bqhw", q.double(), feats.double()))
    got = zb.ops.contraction(q, feats, precision="tf32x3", sigmoid=True, pixel_major=False)
    assert float((got.double() - ref).abs().max()) <= 5e-6
    with pytest.raises(zb.ZutisUnsupported):
        zb.ops.contraction(torch.randn(5, 24).cuda(), torch.randn(1, 4, 4, 24).cuda(), precision="tf32x3")   # K % 32 != 0


@pytest.mark.parametrize("name", ["int8x", "nonint", "x16", "wideq"])
def test_contraction_small_k_cases(zb, golden, name):
    g = golden("decode_cases")
    out = zb.ops.contraction(dev(g[f"{name}_text"]), dev(g[f"{name}_tokens"]))      # K = 32/16/24: SIMT or tcgen05
    ref = g[f"{name}_lowres"]
    assert np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max() <= 1e-5


def test_contraction_batched_queries_with_sigmoid(zb):
    gen = torch.Generator().manual_seed(3)
    q3 = torch.randn(2, 10, 96, generator=gen); q3 = q3 / q3.norm(dim=-1, keepdim=True)
    q4 = torch.randn(2, 3, 10, 96, generator=gen); q4 = q4 / q4.norm(dim=-1, keepdim=True)
    feats = torch.randn(2, 5, 7, 96, generator=gen)
    dec = zb.ZutisDecoder(torch.zeros(2, 96).cuda())
    for q in (q3, q4):
        ref = O.torch_mask_proposals(q, feats).numpy()
        got = dec.get_mask_proposals(q.cuda(), feats.cuda(), return_binary_masks=False)
        assert tuple(got.shape) == ref.shape
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=5e-6)
    raw, one_hot = dec.get_mask_proposals(q3.cuda(), feats.cuda(), return_binary_masks=True)
    ref_raw = torch.einsum("bqc,bhwc->bqhw", q3, feats)
    np.testing.assert_allclose(raw.cpu().numpy(), ref_raw.numpy(), atol=1e-5 * float(ref_raw.abs().max()))
    assert one_hot.dtype == torch.bool and tuple(one_hot.shape) == (2, 10, 5, 7)
    assert torch.equal(one_hot.long().argmax(1).cpu(), raw.cpu().argmax(1))
    assert int(one_hot.sum()) == 2 * 5 * 7


# ------------------------------------------------------------------- fused decode, fixed logits
@pytest.mark.parametrize("layout", ["bqhw", "pixel_major"])
@pytest.mark.parametrize("name", CASES)
def test_decode_labels_bit_exact_vs_c_oracle(zb, golden, name, layout):
    from zutis_b200 import _ffi
    g = golden("decode_cases")
    H, W = (int(v) for v in g[f"{name}_size"])
    lo = g[f"{name}_lowres"]
    want = O.c_decode_semantic(lo, (H, W))
    t = dev(lo)
    if layout == "pixel_major":
        t = pixel_major(t)
    modes = [_ffi.DECODE_AUTO, _ffi.DECODE_GENERIC]
    for mode in modes:
        labels = zb.ops.decode_score(t, (H, W), mode=mode)
        assert labels.dtype == torch.int16
        assert np.array_equal(labels.cpu().numpy().astype(np.int64), want), f"mode {mode}"
    # the golden labels come from ATen itself; identical except (possibly) at fp ties for tiny outputs
    ref = g[f"{name}_labels"].astype(np.int64)
    if max(H, W) > 64 or name in ("ties", "nan"):
        assert np.array_equal(want, ref)
    else:
        assert (want == ref).mean() >= 0.999


def test_decode_size_none_is_plain_argmax(zb, golden):
    g = golden("model_cfg1")
    labels = zb.ops.decode_score(dev(g["lowres_logits"]), None)
    assert np.array_equal(labels.cpu().numpy(), g["labels_lowres"])


def test_tiled_kernel_is_selected_and_equals_generic(zb):
    """The two kernels are independent implementations; they must agree bit for bit."""
    from zutis_b200 import _ffi
    gen = torch.Generator().manual_seed(11)
    for (B, Q, h, w, H, W) in [(3, 81, 20, 20, 160, 160), (2, 81, 13, 17, 107, 139), (1, 130, 9, 9, 150, 150),
                               (2, 7, 6, 8, 96, 128), (1, 81, 14, 14, 224, 224)]:
        lo = torch.randn(B, Q, h, w, generator=gen).cuda()
        for t in (lo, pixel_major(lo)):
            a = zb.ops.decode_score(t, (H, W), mode=_ffi.DECODE_TILED)
            b = zb.ops.decode_score(t, (H, W), mode=_ffi.DECODE_GENERIC)
            assert torch.equal(a, b)
        want = O.c_decode_semantic(lo.cpu().numpy(), (H, W))
        assert np.array_equal(a.cpu().numpy().astype(np.int64), want)
    with pytest.raises(zb.ZutisUnsupported):
        zb.ops.decode_score(torch.randn(1, 5, 30, 30).cuda(), (40, 40), mode=_ffi.DECODE_TILED)   # scale too small


def test_nan_and_inf_follow_torch_argmax(zb):
    from zutis_b200 import _ffi
    gen = torch.Generator().manual_seed(5)
    lo = torch.randn(2, 9, 6, 6, generator=gen)
    lo[0, 3, 2, 2] = float("nan"); lo[0, 5, 2, 3] = float("nan"); lo[1, 4, 0, 0] = float("inf"); lo[1, 2, 5, 5] = float("-inf")
    want = torch.argmax(torch.nn.functional.interpolate(lo, size=(96, 96), mode="bilinear"), dim=1).numpy()
    want_c = O.c_decode_semantic(lo.numpy(), (96, 96))
    assert np.array_equal(want, want_c)
    for mode in (_ffi.DECODE_GENERIC, _ffi.DECODE_TILED, _ffi.DECODE_AUTO):
        got = zb.ops.decode_score(lo.cuda(), (96, 96), mode=mode).cpu().numpy()
        assert np.array_equal(got, want), f"mode {mode}"
    wide = torch.randn(1, 200, 5, 5, generator=gen); wide[0, 150, 1, 1] = float("nan")       # NaN in a later chunk
    want = O.c_decode_semantic(wide.numpy(), (80, 80))
    assert np.array_equal(zb.ops.decode_score(wide.cuda(), (80, 80), mode=_ffi.DECODE_TILED).cpu().numpy(), want)


def _smooth_logits(B, Q, h, w, gen, contrast=1.0):
    """spatially coherent logits: a coarse random field up-sampled x4 (what a segmentation model emits)"""
    coarse = torch.randn(B, Q, (h + 3) // 4 + 1, (w + 3) // 4 + 1, generator=gen)
    return (contrast * torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)).contiguous()


@pytest.mark.parametrize("case", ["smooth", "smooth_noninteger", "x16", "x6", "iid", "ties", "nonfinite_mix", "wide", "wide_overflow", "hist_only", "uniform_regions", "near_ties",
                                  "region_borders", "region_borders_wide", "region_borders_crowded"])
def test_pruned_kernel_is_exact(zb, case):
    """The candidate-pruning (cell) kernel must give the generic kernel's labels and histogram bit for bit -- on coherent
    logits (where it prunes), on noise (where nearly everything survives and lists overflow), on exact ties and
    duplicated categories (first maximum wins), and on cells with NaN/inf taps (its NaN-aware brute-force path)."""
    from zutis_b200 import _ffi
    gen = torch.Generator().manual_seed(hash(case) % 1000 if False else len(case) * 7 + 1)
    want_labels = True
    if case == "smooth":
        B, Q, h, w, H, W = 3, 81, 20, 24, 160, 192; lo = _smooth_logits(B, Q, h, w, gen, 0.1)
    elif case == "smooth_noninteger":
        B, Q, h, w, H, W = 2, 81, 13, 17, 107, 139; lo = _smooth_logits(B, Q, h, w, gen)
    elif case == "x16":
        B, Q, h, w, H, W = 2, 81, 14, 14, 224, 224; lo = _smooth_logits(B, Q, h, w, gen)
    elif case == "x6":
        B, Q, h, w, H, W = 2, 21, 16, 20, 96, 120; lo = _smooth_logits(B, Q, h, w, gen)
    elif case == "iid":
        B, Q, h, w, H, W = 2, 81, 10, 12, 80, 96; lo = torch.randn(B, Q, h, w, generator=gen)
    elif case == "ties":
        B, Q, h, w, H, W = 3, 24, 9, 9, 72, 72
        lo = _smooth_logits(B, Q, h, w, gen)
        lo[:, 5] = lo[:, 17]; lo[:, 20] = lo[:, 2]                     # duplicated categories: the smaller index must win
        lo[1] = 0.0                                                       # a constant image: label 0 everywhere
        lo[2] = torch.round(lo[2] * 2) / 2                                # heavy exact ties between different categories
    elif case == "nonfinite_mix":
        B, Q, h, w, H, W = 4, 40, 8, 10, 64, 80
        lo = _smooth_logits(B, Q, h, w, gen)
        lo[1, 7, 3, 4] = float("nan"); lo[3, 2, 0, 0] = float("inf"); lo[3, 9, 7, 9] = float("-inf")
    elif case == "uniform_regions":                                       # big single-category regions: the whole-cell shortcut
        B, Q, h, w, H, W = 3, 21, 16, 20, 128, 160
        lo = 0.05 * torch.randn(B, Q, h, w, generator=gen)
        region = torch.randint(0, Q, (B, 4, 5), generator=gen).repeat_interleave(4, 1).repeat_interleave(4, 2)
        lo.scatter_add_(1, region[:, None], torch.ones(B, 1, h, w))
        lo[1, 3] = lo[1, 7]                                               # a duplicated category inside the uniform regions
        lo[2, :, :8] = 0.25                                               # a region where every category ties exactly: label 0
    elif case == "near_ties":                                             # differences right at the scale of the pruning margin
        B, Q, h, w, H, W = 4, 32, 12, 12, 96, 96
        base = _smooth_logits(B, 8, h, w, gen)
        lo = base.repeat(1, 4, 1, 1)                                      # every category has three exact copies ...
        scale = lo.abs().amax() * torch.tensor([2.0 ** -e for e in (24, 23, 22, 21, 20, 19, 18)])
        pick = torch.randint(0, 7, lo.shape, generator=gen)
        sign = torch.randint(0, 3, lo.shape, generator=gen).float() - 1.0  # ... perturbed by 0 or +-2^-24..2^-18 of the maximum
        lo = (lo + sign * scale[pick]).contiguous()
    elif case.startswith("region_borders"):
        # trained-model-like: every low-res pixel is confident about its region's category, regions of 3x3 pixels, the other
        # categories are low noise.  On a cell across a border no single category dominates (the first attempt's list
        # overflows); the average of the corner champions does (second attempt, virtual dominator).
        Q = 300 if case.endswith("wide") else 81
        B, h, w, H, W = 2, 12, 15, 96, 120
        lo = 0.02 * torch.randn(B, Q, h, w, generator=gen)
        region = torch.randint(0, Q, (B, 4, 5), generator=gen).repeat_interleave(3, 1).repeat_interleave(3, 2)
        lo.scatter_add_(1, region[:, None], torch.ones(B, 1, h, w))
        if case.endswith("crowded"):
            # every other category sits within 2^-18 of 0.5, i.e. within the margin of k* on the far side of a border (k*
            # is 1 on its side and ~0.5 on the other) but 0.25 below the virtual dominator
            lo = 0.5 + 2.0 ** -18 * torch.randn(B, Q, h, w, generator=gen)
            lo.scatter_(1, region[:, None], torch.ones(B, 1, h, w))
    elif case == "wide":
        B, Q, h, w, H, W = 1, 920, 7, 8, 56, 64; lo = _smooth_logits(B, Q, h, w, gen, 0.2)
    elif case == "wide_overflow":                                         # noise: ~390 of 600 categories survive, more than the 256 slots
        B, Q, h, w, H, W = 1, 600, 5, 6, 40, 48; lo = torch.randn(B, Q, h, w, generator=gen)
    else:
        B, Q, h, w, H, W = 2, 81, 10, 10, 80, 80; lo = _smooth_logits(B, Q, h, w, gen); want_labels = False
    gt = torch.randint(0, Q, (B, H, W), generator=gen); gt[:, :3] = 1000
    t = pixel_major(lo.cuda())
    ref_part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
    ref = zb.ops.decode_score(t, (H, W), gt=gt.cuda(), hist_partial=ref_part, mode=_ffi.DECODE_GENERIC)
    if not case.startswith("wide"):
        assert np.array_equal(ref.cpu().numpy().astype(np.int64), O.c_decode_semantic(lo.numpy(), (H, W)))
    if case.startswith("region_borders"):                               # the case is what it claims: k* alone would overflow
        cells = torch.stack([lo[:, :, :-1, :-1], lo[:, :, :-1, 1:], lo[:, :, 1:, :-1], lo[:, :, 1:, 1:]])     # 4,B,Q,h-1,w-1
        kstar = cells.amin(0).argmax(1, keepdim=True)                                                           # B,1,h-1,w-1
        K = torch.gather(cells, 2, kstar[None].expand(4, -1, -1, -1, -1))
        survivors = (~((K - cells) >= K.abs().amax(0) * 2.0 ** -20).all(0)).sum(1)
        assert (survivors > 40).float().mean() > 0.2
    for mode in (_ffi.DECODE_CELLS, _ffi.DECODE_AUTO):
        part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
        got = zb.ops.decode_score(t, (H, W), gt=gt.cuda(), hist_partial=part, mode=mode, want_labels=want_labels)
        if want_labels:
            assert torch.equal(got, ref), f"{case}: labels differ in mode {mode}: {(got != ref).sum().item()} pixels"
        assert torch.equal(part, ref_part), f"{case}: histogram differs in mode {mode}"
        if want_labels:                                                  # labels without a histogram
            assert torch.equal(zb.ops.decode_score(t, (H, W), mode=mode), ref)
    if case == "smooth":                                                  # every ground-truth storage type (the kernel is templated on it)
        gt8 = gt.clamp(max=255)
        for dt in (torch.uint8, torch.int16, torch.int32):
            g = (gt8 if dt == torch.uint8 else gt).to(dt).cuda()
            want = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
            zb.ops.decode_score(t, (H, W), gt=g, hist_partial=want, mode=_ffi.DECODE_GENERIC, want_labels=False)
            part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
            zb.ops.decode_score(t, (H, W), gt=g, hist_partial=part, mode=_ffi.DECODE_CELLS, want_labels=False)
            assert torch.equal(part, want), f"cell kernel: histogram differs for {dt}"
    # the pruned kernel needs contiguous categories; a query-major tensor must be refused in forced mode, not mis-read
    with pytest.raises(zb.ZutisUnsupported):
        zb.ops.decode_score(lo.cuda(), (H, W), mode=_ffi.DECODE_CELLS)
    assert torch.equal(zb.ops.decode_score(lo.cuda(), (H, W)), ref)


def test_cell_kernel_work_distribution_and_workspace(zb):
    """The cell kernel hands out its runs through a global counter when it gets a workspace (dynamic) and by cell rows
    otherwise (static): same labels and histogram, the workspace is left zeroed, a dirty workspace is re-armed unless
    the caller vouches for it, in-place changes of the logits are seen, and decode_and_score matches the generic kernel."""
    import ctypes as C
    from zutis_b200 import _ffi
    gen = torch.Generator().manual_seed(3)
    B, Q, D, h, w, H, W = 5, 81, 512, 20, 24, 160, 192
    text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).cuda()
    coarse = torch.randn(B, D, h // 2, w // 2, generator=gen)
    tokens = torch.nn.functional.normalize(torch.nn.functional.interpolate(coarse, scale_factor=2, mode="bilinear").permute(0, 2, 3, 1), dim=-1).contiguous().cuda()
    tokens[3] = torch.nn.functional.normalize(torch.randn(h, w, D, generator=gen), dim=-1).cuda()     # an incoherent image: long lists, overflow
    tokens[4, 2, 3, 7] = float("nan")                                                                  # a poisoned pixel: brute-force cells
    gt = torch.randint(0, Q, (B, H, W), generator=gen).cuda()
    lo = zb.ops.contraction(text, tokens, precision="tf32x3")
    ref_part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
    ref = zb.ops.decode_score(lo, (H, W), gt=gt, hist_partial=ref_part, mode=_ffi.DECODE_GENERIC)
    lib = _ffi.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def raw(ws, mode):
        labels = torch.empty(B, H, W, dtype=torch.int16, device="cuda")
        part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
        args = (lo.data_ptr(), lo.stride(0), lo.stride(1), lo.stride(2), lo.stride(3), B, Q, h, w, H, W, gt.data_ptr(), _ffi.GT_I64, H * W,
                labels.data_ptr(), part.data_ptr(), Q)
        if ws is None:
            _ffi.check(lib.zutis_decode_score(*args, mode, stream))
        else:
            _ffi.check(lib.zutis_decode_score_ws(*args, mode, ws.data_ptr(), ws.numel(), stream))
        return labels, part

    assert lib.zutis_decode_workspace_bytes(B, Q, h, w, H, W) == 16
    cases = ((None, _ffi.DECODE_CELLS),                                                                  # static rows
             (torch.zeros(16, dtype=torch.uint8, device="cuda"), _ffi.DECODE_CELLS | _ffi.DECODE_WORKSPACE_ZEROED),   # global run counter
             (torch.full((16,), 0xAB, dtype=torch.uint8, device="cuda"), _ffi.DECODE_AUTO))              # dirty counter: the call re-arms it
    for ws, mode in cases:
        for _ in range(3):                                                                           # the workspace is reusable
            labels, part = raw(ws, mode)
            assert torch.equal(labels, ref) and torch.equal(part, ref_part)
            if ws is not None:
                assert int(ws[:16].view(torch.int32).abs().sum()) == 0
            mode |= _ffi.DECODE_WORKSPACE_ZEROED if ws is not None else 0
    # logits changed in place after the contraction are simply decoded as they are
    lo3 = zb.ops.contraction(text, tokens, precision="tf32x3")
    lo3[:, 5] += 1.0
    want3 = zb.ops.decode_score(lo3, (H, W), mode=_ffi.DECODE_GENERIC)
    assert torch.equal(zb.ops.decode_score(lo3, (H, W), workspace=zb.ops.DecodeWorkspace()), want3)
    meter = zb.RunningScore(Q)
    labels = zb.decode_and_score(text, tokens, gt, (H, W), meter, want_labels=True, precision="tf32x3")
    assert torch.equal(labels, ref) and torch.equal(meter.counts().view(-1).to(torch.int32), ref_part)


@pytest.mark.parametrize("gt_dtype", [torch.uint8, torch.int16, torch.int32, torch.int64])
def test_fused_histogram_bit_exact(zb, gt_dtype):
    from zutis_b200 import _ffi
    gen = torch.Generator().manual_seed(21)
    B, Q, h, w, H, W = 3, 81, 12, 15, 96, 120
    lo = torch.randn(B, Q, h, w, generator=gen)
    gt = torch.randint(0, Q, (B, H, W), generator=gen)
    gt[:, :5] = 255; gt[1, 40:50, 10:60] = 200
    if gt_dtype != torch.uint8:
        gt[2, 7] = -1
        gt[0, 9, :30] = 1000
    labels_ref = O.c_decode_semantic(lo.numpy(), (H, W))
    gt_np = gt.to(gt_dtype).numpy().astype(np.int64)
    want = O.c_fast_hist(gt_np, labels_ref, Q)
    for mode in (_ffi.DECODE_TILED, _ffi.DECODE_GENERIC):
        part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
        labels = zb.ops.decode_score(pixel_major(lo.cuda()), (H, W), gt=gt.to(gt_dtype).cuda(), hist_partial=part, mode=mode)
        assert np.array_equal(labels.cpu().numpy(), labels_ref)
        assert np.array_equal(part.view(Q, Q).cpu().numpy().astype(np.int64), want)
        hist = torch.zeros(Q * Q, dtype=torch.int64, device="cuda")
        zb.ops.hist_merge(part, hist, clear=True)
        zb.ops.hist_merge(part, hist, clear=True)              # partial was cleared: second merge adds nothing
        assert np.array_equal(hist.view(Q, Q).cpu().numpy(), want) and int(part.abs().sum()) == 0
    # histogram without labels, wide class count (global-atomic path), all-ignored image
    Qw = 300
    low = torch.randn(2, Qw, 5, 6, generator=gen)
    gtw = torch.randint(0, Qw, (2, 80, 96), generator=gen); gtw[1] = 1000
    part = torch.zeros(Qw * Qw, dtype=torch.int32, device="cuda")
    assert zb.ops.decode_score(low.cuda(), (80, 96), gt=gtw.cuda(), hist_partial=part, want_labels=False) is None
    want = O.c_fast_hist(gtw.numpy(), O.c_decode_semantic(low.numpy(), (80, 96)), Qw)
    assert np.array_equal(part.view(Qw, Qw).cpu().numpy().astype(np.int64), want)
    assert want.sum() == 80 * 96


# ------------------------------------------------------------------ drop-in predict + RunningScore
def test_semantic_predict_drop_in_model_cfg1(zb, golden):
    """ZUTIS.predict('semantic') + RunningScore on the reference's own random-init ViT-B/32 outputs."""
    g = golden("model_cfg1")
    dec = zb.ZutisDecoder(dev(g["text"]))
    pred = dec.predict({"patch_tokens": dev(g["tokens"])}, "semantic", size=(224, 224))
    assert isinstance(pred, np.ndarray) and pred.dtype == np.int64 and pred.shape == (2, 224, 224)
    ref = g["labels"].astype(np.int64)
    agree = (pred == ref).mean()
    assert agree >= 0.9999, agree
    full = O.torch_semantic_predict(torch.from_numpy(g["text"]), torch.from_numpy(g["tokens"]), (224, 224), return_logits=True).numpy()
    assert tie_gap_ok(full, pred, ref)
    # confusion matrix: bit-exact for the labels the kernel produced; scores within 1e-4 of the reference's
    meter = zb.RunningScore(81)
    meter.update(g["gt"].astype(np.int64), pred)                      # host numpy, like trainer.py:347
    assert np.array_equal(meter.confusion_matrix, O.c_fast_hist(g["gt"], pred, 81).astype(np.float64))
    s, cls = meter.get_scores()
    ref_scores = g["scores"]
    got = np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]])
    assert np.abs(got - ref_scores).max() <= 1e-4
    if agree == 1.0:
        assert np.array_equal(got, ref_scores) and np.array_equal(meter.confusion_matrix, g["confusion"].astype(np.float64))
    # fused path gives the same matrix without host labels
    fused = zb.RunningScore(81)
    dec.decode_and_score(dev(g["tokens"]), dev(g["gt"]), (224, 224), fused)
    assert np.array_equal(fused.confusion_matrix, meter.confusion_matrix)
    # return_logits=True still returns full-resolution fp32 logits
    logits = dec.predict({"patch_tokens": dev(g["tokens"])}, "semantic", size=(224, 224), return_logits=True)
    assert tuple(logits.shape) == (2, 81, 224, 224) and logits.dtype == torch.float32
    assert np.abs(logits.cpu().numpy() - full).max() <= 1e-5 * np.abs(full).max() + 1e-6
    lowres_labels = dec.predict({"patch_tokens": dev(g["tokens"])}, "semantic")          # size=None
    assert (lowres_labels == g["labels_lowres"]).mean() >= 0.99


@pytest.mark.parametrize("name", ["int8x", "nonint", "x16", "down", "wideq", "ties"])
def test_semantic_predict_drop_in_cases(zb, golden, name):
    g = golden("decode_cases")
    H, W = (int(v) for v in g[f"{name}_size"])
    dec = zb.ZutisDecoder(dev(g[f"{name}_text"]))
    size = (H, W) if name != "nonint" else [torch.tensor([H]), torch.tensor([W])]        # trainer.py:322-323 form
    pred = dec.predict({"patch_tokens": dev(g[f"{name}_tokens"])}, "semantic", size=size)
    ref = g[f"{name}_labels"].astype(np.int64)
    full = O.torch_semantic_predict(torch.from_numpy(g[f"{name}_text"]), torch.from_numpy(g[f"{name}_tokens"]), (H, W),
                                    return_logits=True).numpy()
    if name == "ties":
        # exact duplicates of a category: the first index must win everywhere (torch.argmax rule)
        assert not np.isin(pred, [4, 5]).any() and np.array_equal(pred, ref)
    else:
        assert tie_gap_ok(full, pred, ref)
        assert (pred == ref).mean() >= (0.9999 if pred.size >= 20000 else 0.995)


def test_upsample_bilinear_bit_exact(zb, golden):
    g = golden("decode_cases")
    for name in ("int8x", "nonint", "x16", "same"):
        H, W = (int(v) for v in g[f"{name}_size"])
        out = zb.ops.upsample_bilinear(dev(g[f"{name}_lowres"][:, :4]), (H, W)).cpu().numpy()
        assert np.array_equal(out.view(np.int32), g[f"{name}_full4"].view(np.int32))


def test_running_score_drop_in(zb, golden):
    g = golden("scoring")
    for n in (3, 4):
        m = zb.RunningScore(n)
        m.update(g["ka_gt"][None], g["ka_pred"][None])
        assert np.array_equal(m.confusion_matrix, g[f"ka{n}_confusion"]) and m.confusion_matrix.dtype == np.float64
        s, c = m.get_scores()
        assert np.array_equal(np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]]), g[f"ka{n}_scores"])
        assert np.array_equal(np.array([c[i] for i in range(n)]), g[f"ka{n}_class_iou"], equal_nan=True)
        assert list(s) == ["Pixel Acc", "Mean Acc", "FreqW Acc", "Mean IoU"] and list(c) == list(range(n))
    # wide matrix, two updates, ignore labels 1000 and negatives
    gt, pr = g["wide_gt"].astype(np.int64), g["wide_pred"].astype(np.int64)
    m = zb.RunningScore(920)
    m.update(gt, pr); m.update(torch.from_numpy(gt[:1]).cuda(), torch.from_numpy(pr[:1]).cuda().to(torch.int16))
    ref = np.zeros((920, 920)); r, c, v = g["wide_confusion_nz"]; ref[r.astype(int), c.astype(int)] = v
    assert np.array_equal(m.confusion_matrix, ref)
    s, cls = m.get_scores()
    assert np.array_equal(np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]]), g["wide_scores"])
    assert np.array_equal(np.array([cls[i] for i in range(920)]), g["wide_class_iou"], equal_nan=True)
    # ragged list of differently sized images
    m = zb.RunningScore(7)
    m.update([g["rag_ga"], g["rag_gb"]], [g["rag_pa"], g["rag_pb"]])
    assert np.array_equal(m.confusion_matrix, g["rag_confusion"])
    m.reset()
    assert m.confusion_matrix.sum() == 0
    s, _ = m.get_scores()                                    # empty matrix -> nan / nan / 0 / nan
    assert np.isnan(s["Pixel Acc"]) and np.isnan(s["Mean Acc"]) and s["FreqW Acc"] == 0.0 and np.isnan(s["Mean IoU"])
    # an all-ignored image changes nothing
    m.update(np.full((1, 5, 5), 255), np.zeros((1, 5, 5), np.int64))
    assert m.confusion_matrix.sum() == 0


def test_compute_iou_drop_in(zb, golden):
    g = golden("scoring")
    a = np.array([[1, 1, 0], [0, 1, 0]], bool); b = np.array([[1, 0, 0], [0, 1, 1]], bool)
    r = zb.compute_iou(a, b)
    assert isinstance(r, np.floating) and r == float(g["iou_bool"]) == 0.4999999875000003
    pf = np.array([[.6, .4, .9], [.1, .7, .2]])
    assert zb.compute_iou(pf, b, threshold=0.5) == float(g["iou_thr"])
    assert zb.compute_iou(np.zeros((2, 3), bool), np.zeros((2, 3), bool)) == 0.0
    assert zb.compute_iou(g["iou_rand_a"], g["iou_rand_b"]) == float(g["iou_rand"])
    t = zb.compute_iou(torch.from_numpy(a), torch.from_numpy(b))
    assert isinstance(t, torch.Tensor) and t.dim() == 0 and not t.is_cuda
    assert np.array_equal(t.numpy(), g["iou_torch"])
    with pytest.raises(AssertionError):
        zb.compute_iou(np.zeros((2, 3)), np.zeros((3, 2)))


def test_streaming_scorer_equals_batch_by_batch(zb):
    """StreamingScorer (decode + scoring on its own stream) against decode_and_score batch by batch: same matrix, and
    the caller's stream really is left free (the scoring of a batch is still pending when submit returns)."""
    gen = torch.Generator().manual_seed(31)
    Q, D, h, w, H, W = 21, 64, 14, 14, 112, 112
    text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).cuda()
    batches = [(torch.randn(3, h, w, D, generator=gen).cuda(), torch.randint(0, Q, (3, H, W), generator=gen).cuda()) for _ in range(5)]
    ref = zb.RunningScore(Q, device="cuda")
    for tok, gt in batches:
        zb.decode_and_score(text, tok, gt, (H, W), ref)
    want = ref.confusion_matrix
    meter = zb.RunningScore(Q, device="cuda")
    scorer = zb.StreamingScorer(text, meter, (H, W))
    for tok, gt in batches:
        scorer.submit(tok, gt)
    scores, _ = scorer.get_scores()
    assert np.array_equal(meter.confusion_matrix, want)
    assert scores["Mean IoU"] == ref.get_scores()[0]["Mean IoU"]
    # the scoring stream is not the caller's
    assert scorer._stream.cuda_stream != torch.cuda.current_stream().cuda_stream


# --------------------------------------------------------------------------------- instance path
def test_threshold_masks_bit_exact(zb):
    gen = torch.Generator().manual_seed(4)
    probs = torch.sigmoid(3 * torch.randn(2, 10, 6, 8, generator=gen))
    for size in [(48, 64), (50, 70), None]:
        want = O.c_decode_threshold(probs.numpy(), size, 0.5)
        bits, areas = zb.ops.decode_threshold(probs.cuda(), size, 0.5)
        W = want.shape[-1]
        got = zb.ops.unpack_mask_bits(bits, W).cpu().numpy()
        assert np.array_equal(got, want)
        assert np.array_equal(areas.cpu().numpy(), want.sum((-2, -1)))
        inter = zb.ops.pairwise_mask_intersections(bits[0]).cpu().numpy()
        flat = want[0].reshape(10, -1).astype(np.int64)
        assert np.array_equal(inter, flat @ flat.T)


@pytest.mark.parametrize("fixture,prefix", [("instance_cases", ""), ("model_cfg1", "inst_")])
@pytest.mark.parametrize("tag", ["hard", "none", "linear", "gaussian"])
def test_instance_predict_drop_in(zb, golden, fixture, prefix, tag):
    g = golden(fixture)
    if f"{prefix}{tag}_score" not in g:
        pytest.skip("soft-NMS outputs of the reference are stored for the synthetic proposals only")
    size = tuple(int(v) for v in g["size"]) if "size" in g else (224, 224)
    ids = [5, 6] if fixture == "instance_cases" else [11, 22]
    dec = zb.ZutisDecoder(dev(g["text"]))
    preds = dec.predict({"mask_proposals": dev(g["proposals"]), "patch_tokens": dev(g["tokens"])}, "instance",
                        size=size, image_ids=ids, nms_type=None if tag == "none" else tag)
    assert [p["category_id"] for p in preds] == g[f"{prefix}{tag}_category"].tolist()
    assert [p["image_id"] for p in preds] == g[f"{prefix}{tag}_image_id"].tolist()
    np.testing.assert_allclose(np.array([p["score"] for p in preds], np.float64).reshape(-1), g[f"{prefix}{tag}_score"], rtol=5e-6, atol=1e-9)
    ref_bits = g[f"{prefix}{tag}_mask_bits"]
    H, W = size
    for p, rb, box in zip(preds, ref_bits, g[f"{prefix}{tag}_bbox"]):
        assert set(p) >= {"category_id", "segmentation", "score", "image_id", "image_size", "bbox"}
        assert tuple(p["image_size"]) == (H, W)
        mask = np.unpackbits(rb)[: H * W].reshape(H, W).astype(bool)
        seg = p["segmentation"]                                   # pycocotools.mask.encode's dict, built on the device
        assert seg["size"] == [H, W] and isinstance(seg["counts"], bytes)
        assert seg["counts"] == O.rle_to_string(O.rle_counts_numpy(mask))
        flat = np.zeros(H * W, np.uint8); pos = 0; val = 0
        for run in O.rle_from_string(seg["counts"]):
            flat[pos:pos + run] = val; pos += run; val ^= 1
        assert pos == H * W and np.array_equal(flat.reshape(W, H).T.astype(bool), mask)
        assert p["bbox"] == list(box)


def _pack_rows(masks: np.ndarray) -> np.ndarray:
    """bool [n,H,W] -> int32 [n,H,ceil(W/32)], bit x%32 of word [y][x/32] = pixel (y,x)."""
    n, H, W = masks.shape
    words = (W + 31) // 32
    padded = np.zeros((n, H, words * 32), np.uint8)
    padded[:, :, :W] = masks
    return np.packbits(padded.reshape(n, H, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(n, H, words).view(np.int32)

def test_device_hard_nms_matches_reference_loop(zb):
    """The device suppression (ops.instance_nms_hard) against the host replay of zutis.py:232-282 on random overlapping
    masks: few categories so that chains of suppression are long, low scores around the 0.001 floor, empty masks, a
    background category, and -- flagged, not resolved -- duplicated scores."""
    from zutis_b200.decode import _nms_keep, _ordered_picks
    rng = np.random.default_rng(21)
    for trial, (B, M, n_cat) in enumerate([(3, 40, 3), (2, 100, 5), (1, 7, 2), (2, 64, 1), (1, 1, 2)]):
        H, W = 24, 40
        masks = np.zeros((B, M, H, W), bool)
        for b in range(B):
            for i in range(M):
                if rng.random() < 0.1:
                    continue                                   # empty mask
                y0, x0 = rng.integers(0, H - 4), rng.integers(0, W - 4)
                masks[b, i, y0:y0 + rng.integers(2, 12), x0:x0 + rng.integers(2, 16)] = True
        cats = rng.integers(0, n_cat + 1, (B, M)).astype(np.int64)
        scores = rng.random((B, M)).astype(np.float32)
        scores[rng.random((B, M)) < 0.2] *= np.float32(0.002)       # around the floor
        if trial == 3:
            scores[1, 5] = scores[1, 9]; cats[1, 5] = cats[1, 9] = 1   # a tie inside one category
        bits = torch.from_numpy(np.stack([_pack_rows(m) for m in masks])).cuda()
        inter = torch.stack([zb.ops.pairwise_mask_intersections(bits[b]) for b in range(B)])
        pick, tie = zb.ops.instance_nms_hard(inter, torch.from_numpy(cats).to(torch.int32).cuda(), torch.from_numpy(scores).cuda())
        pick, tie, inter_h = pick.cpu().numpy(), tie.cpu().numpy(), inter.cpu().numpy()
        for b in range(B):
            if trial == 3 and b == 1:
                assert tie[b] != 0
                continue
            assert tie[b] == 0
            want = _nms_keep(cats[b], scores[b], inter_h[b], "hard")
            got = _ordered_picks(cats[b], scores[b], pick[b], inter_h[b].diagonal())
            assert [(int(c), q, float(s)) for c, q, s in got] == [(int(c), q, float(s)) for c, q, s in want]



@pytest.mark.parametrize("H,W", [(1, 1), (1, 70), (33, 1), (5, 32), (31, 33), (64, 64), (50, 70), (480, 640), (97, 1000)])
def test_mask_rle_and_boxes_match_oracle(zb, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    n = 12
    masks = np.zeros((n, H, W), bool)
    masks[1] = True                                                                  # full: runs [0, H*W]
    masks[2, 0, 0] = True                                                            # starts with a set pixel
    masks[3, -1, -1] = True                                                          # ends with a set pixel
    masks[4] = rng.random((H, W)) < 0.5                                              # noise: ~H*W/2 runs
    masks[5, :, W // 2:] = True                                                      # right half: one transition
    masks[6, H // 2:, :] = True                                                      # bottom half: a run per column
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(7, n):                                                            # blobs
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(0.5, max(H, W) / 2 + 1)
        masks[i] = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
    bits = torch.from_numpy(_pack_rows(masks)).cuda()
    order = [3, 0, 7, 1, 2, 4, 5, 6, 8, 9, 10, 11, 4]                                # selection, repeats allowed
    for ids in (None, torch.tensor(order)):
        sel = masks if ids is None else masks[order]
        n_runs, runs, boxes = zb.ops.mask_rle(bits, W, mask_ids=ids)
        offs = np.cumsum(n_runs) - n_runs
        for i, m in enumerate(sel):
            want = O.rle_counts_numpy(m)
            got = runs[offs[i]: offs[i] + n_runs[i]].astype(np.int64)
            assert np.array_equal(got, want), (H, W, i)
            if m.any():
                assert [float(v) for v in boxes[i]] == O.mask_to_box(m)
            else:
                assert boxes[i].tolist() == [-1, -1, -1, -1]
        from zutis_b200.decode import _rle_strings
        strings = _rle_strings(n_runs, runs)                                         # host (numpy) compressor
        dev_strings, dev_boxes = zb.ops.mask_rle_strings(bits, W, mask_ids=ids)      # device compressor
        assert np.array_equal(dev_boxes, boxes)
        for i, m in enumerate(sel):
            want = O.rle_to_string(O.rle_counts_numpy(m))
            assert strings[i] == want and dev_strings[i] == want


# ----------------------------------------------------------------------------- host-buffer entry
def test_semantic_eval_host_entry(zb, golden):
    import ctypes as C
    from zutis_b200 import _ffi
    g = golden("model_cfg1")
    text = np.ascontiguousarray(g["text"]); tokens = np.ascontiguousarray(g["tokens"])
    gt = np.ascontiguousarray(g["gt"].astype(np.int64))
    hist = np.zeros((81, 81), np.int64); labels = np.zeros((2, 224, 224), np.int16)
    _ffi.call("zutis_semantic_eval_host", text.ctypes.data, tokens.ctypes.data, gt.ctypes.data, _ffi.GT_I64,
              2, 81, 512, 14, 14, 224, 224, hist.ctypes.data, labels.ctypes.data, _ffi.GEMM_FP32_SIMT, 0)
    assert (labels == g["labels"]).mean() >= 0.9999
    assert np.array_equal(hist, O.c_fast_hist(gt, labels.astype(np.int64), 81))


# ------------------------------------------------------- BASELINE-size runs: size-independent properties
FULL = {
    "cfg2": dict(B=64, Q=81, D=512, h=40, w=40, H=320, W=320, ignore=255),
    "cfg3": dict(B=32, Q=81, D=512, h=64, w=64, H=512, W=512, ignore=255),
    "cfg4": dict(B=32, Q=920, D=512, h=56, w=56, H=448, W=448, ignore=1000),
}


@pytest.mark.parametrize("cfg", list(FULL))
def test_full_size_properties(zb, cfg):
    """At BASELINE.json's sizes the oracle is too slow, so check what must hold regardless of size:
    counts conserve pixels, rows reproduce the ground-truth class histogram, columns the label
    histogram, the fused matrix equals scoring the emitted labels, and two kernels agree."""
    from zutis_b200 import _ffi
    c = FULL[cfg]
    B, Q, D, h, w, H, W = (c[k] for k in ("B", "Q", "D", "h", "w", "H", "W"))
    gen = torch.Generator(device="cuda").manual_seed(0)
    text = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=gen), dim=-1)
    coarse = torch.randn(B, D, h // 2, w // 2, device="cuda", generator=gen)
    tokens = torch.nn.functional.normalize(torch.nn.functional.interpolate(coarse, scale_factor=2, mode="bilinear").permute(0, 2, 3, 1), dim=-1).contiguous()
    gt = torch.randint(0, Q, (B, H, W), device="cuda", generator=gen)
    gt[:, :9] = c["ignore"]; gt[0] = c["ignore"]
    meter = zb.RunningScore(Q)
    labels = zb.decode_and_score(text, tokens, gt, (H, W), meter, want_labels=True)
    counts = meter.counts()
    valid = (gt >= 0) & (gt < Q)
    assert int(counts.sum()) == int(valid.sum())
    assert torch.equal(counts.sum(1), torch.bincount(gt[valid], minlength=Q))
    assert torch.equal(counts.sum(0), torch.bincount(labels[valid].long(), minlength=Q))
    assert int(labels.min()) >= 0 and int(labels.max()) < Q
    again = zb.RunningScore(Q)
    again.update(gt, labels)                                  # scoring the emitted labels separately
    assert torch.equal(again.counts(), counts)
    # first two images: independent generic kernel and the C oracle's decode on the same low-res logits
    lowres = zb.ops.contraction(text, tokens[:2])
    a = zb.ops.decode_score(lowres, (H, W), mode=_ffi.DECODE_GENERIC)
    assert torch.equal(a, labels[:2])
    if Q <= 128:
        want = O.c_decode_semantic(lowres[:1].cpu().numpy(), (H, W))
        assert np.array_equal(labels[:1].cpu().numpy().astype(np.int64), want)
    ref_lo = torch.einsum("nc,bhwc->bnhw", text.double(), tokens[:2].double())
    assert float((lowres.double() - ref_lo).abs().max() / ref_lo.abs().max()) <= 1e-5
    s, _ = meter.get_scores()
    assert 0.0 <= s["Mean IoU"] <= 1.0


def test_cfg5_instance_full_size(zb):
    """BASELINE config 5 (COCO-20K instance path): 100 queries, 60x80 -> 480x640, batch 16.
    Contraction+sigmoid, low-res statistics, full-resolution bit masks and pairwise intersections at
    full size; the oracle checks two images, size-independent identities cover the batch."""
    B, Q, D, h, w, H, W = 16, 100, 768, 60, 80, 480, 640
    gen = torch.Generator(device="cuda").manual_seed(5)
    queries = torch.nn.functional.normalize(torch.randn(B, Q, D, device="cuda", generator=gen), dim=-1)
    coarse = torch.randn(B, D, h // 2, w // 2, device="cuda", generator=gen)
    feats = (4 * torch.nn.functional.interpolate(coarse, scale_factor=2, mode="bilinear").permute(0, 2, 3, 1)).contiguous()
    dec = zb.ZutisDecoder(torch.nn.functional.normalize(torch.randn(81, 512, device="cuda", generator=gen), dim=-1))
    probs = dec.get_mask_proposals(queries, feats, return_binary_masks=False)
    assert tuple(probs.shape) == (B, Q, h, w)
    ref = torch.sigmoid(torch.einsum("bqc,bhwc->bqhw", queries[:2].double(), feats[:2].double()))
    assert float((probs[:2].double() - ref).abs().max()) <= 1e-5
    assert 0.0 <= float(probs.min()) and float(probs.max()) <= 1.0
    bits, areas = zb.ops.decode_threshold(probs, (H, W), 0.5)
    assert tuple(bits.shape) == (B, Q, H, W // 32)
    want = O.c_decode_threshold(probs[:2].cpu().numpy(), (H, W), 0.5)
    got = zb.ops.unpack_mask_bits(bits[:2], W).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(areas[:2].cpu().numpy(), want.sum((-2, -1)))
    inter = zb.ops.pairwise_mask_intersections(bits[3])
    assert torch.equal(inter.diagonal(), areas[3]) and torch.equal(inter, inter.t())
    assert int((inter > torch.minimum(areas[3][:, None], areas[3][None, :])).sum()) == 0
    # low-res statistics against the torch port on two images (0.98 GB broadcast per image in the reference)
    tokens = torch.nn.functional.normalize(torch.randn(2, h, w, 512, device="cuda", generator=gen), dim=-1)
    sizes, psum, mean = zb.ops.instance_lowres_stats(probs[:2], tokens, 0.5)
    conf_ref, cat_ref, sizes_ref = O.torch_instance_lowres(dec.text_embeddings.cpu(), probs[:2].cpu(), tokens.cpu())
    assert np.array_equal(sizes.cpu().numpy(), sizes_ref)
    cat, prob = zb.ops.instance_categories(mean, dec.text_embeddings, 5.0)
    conf = (psum.cpu().numpy() / (sizes.cpu().numpy().astype(np.float32) + np.float32(1e-7))) * prob.cpu().numpy()
    np.testing.assert_allclose(conf, conf_ref, rtol=2e-5, atol=1e-7)
    assert (cat.cpu().numpy() == cat_ref).mean() >= 0.99          # argmax over sigmoid(5*cos): ties only


def test_host_entry_chunks_and_overlaps(zb):
    """zutis_semantic_eval_host with more images than one chunk (two streams, H2D overlapped with kernels)."""
    from zutis_b200 import _ffi
    B, Q, D, h, w, H, W = 24, 81, 512, 40, 40, 320, 320
    gen = torch.Generator().manual_seed(2)
    text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).numpy()
    tokens = torch.nn.functional.normalize(torch.randn(B, h, w, D, generator=gen), dim=-1).numpy()
    gt = torch.randint(0, Q, (B, H, W), generator=gen).numpy().astype(np.int32); gt[:, :3] = 255
    hist = np.zeros((Q, Q), np.int64); labels = np.zeros((B, H, W), np.int16)
    _ffi.call("zutis_semantic_eval_host", text.ctypes.data, tokens.ctypes.data, gt.ctypes.data, _ffi.GT_I32,
              B, Q, D, h, w, H, W, hist.ctypes.data, labels.ctypes.data, _ffi.GEMM_TF32X3, 0)
    dev_labels = zb.ZutisDecoder(torch.from_numpy(text).cuda()).predict({"patch_tokens": torch.from_numpy(tokens).cuda()}, "semantic", size=(H, W))
    assert np.array_equal(labels.astype(np.int64), dev_labels)
    assert np.array_equal(hist, O.c_fast_hist(gt, labels, Q)) and hist.sum() == B * (H - 3) * W
    _ffi.call("zutis_semantic_eval_host", text.ctypes.data, tokens.ctypes.data, gt.ctypes.data, _ffi.GT_I32,
              B, Q, D, h, w, H, W, hist.ctypes.data, None, _ffi.GEMM_TF32X3, 0)          # accumulates into hist
    assert hist.sum() == 2 * B * (H - 3) * W


@pytest.mark.parametrize("Q,gt_dtype", [(81, np.int64), (300, np.int64), (300, np.int32), (81, np.int16), (255, np.int64), (256, np.int64)])
def test_host_entry_narrows_labels_exactly(zb, Q, gt_dtype):
    """The host entry ships int64 / int32 / int16 labels as uint8 or int16 (csrc/host_eval.cu).  Everything the reference
    ignores (negative, >= n, values whose low byte would look valid) must stay ignored; big batch = threaded path."""
    from zutis_b200 import _ffi
    B, D, h, w, H, W = (20 if Q == 81 else 3), 64, 12, 12, 96, 96
    if Q == 81 and gt_dtype == np.int64:
        H, W = 320, 320                                                # > 4 MB of labels: the worker threads run
    gen = torch.Generator().manual_seed(Q)
    text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).numpy()
    tokens = torch.nn.functional.normalize(torch.randn(B, h, w, D, generator=gen), dim=-1).numpy()
    gt = torch.randint(0, Q, (B, H, W), generator=gen).numpy().astype(gt_dtype)
    big = np.iinfo(gt_dtype).max
    odd = [-1, -256, Q, Q + 1, 255, 256, 32767, big, big - 250, np.iinfo(gt_dtype).min]
    if gt_dtype == np.int64:
        odd += [2 ** 40 + 5, 2 ** 32, -(2 ** 33) + 7, 65536 + 3]
    for k, v in enumerate(odd):
        gt[:, k % H, (7 * k) % W::11] = v
    code = {np.int64: _ffi.GT_I64, np.int32: _ffi.GT_I32, np.int16: _ffi.GT_I16}[gt_dtype]
    hist = np.zeros((Q, Q), np.int64); labels = np.zeros((B, H, W), np.int16)
    _ffi.call("zutis_semantic_eval_host", text.ctypes.data, tokens.ctypes.data, gt.ctypes.data, code,
              B, Q, D, h, w, H, W, hist.ctypes.data, labels.ctypes.data, _ffi.GEMM_TF32X3, 0)
    assert np.array_equal(hist, O.c_fast_hist(gt.astype(np.int64), labels.astype(np.int64), Q))
    assert hist.sum() == int(((gt >= 0) & (gt < Q)).sum())


def test_allreduce_hist_over_a_raw_nccl_communicator(zb):
    """zutis_allreduce_hist: the int64 matrix summed in place over an ncclComm_t made with NCCL's own C API.
    One GPU: a communicator of size 1 (the sum is the matrix itself; checks symbol resolution and the call).  Two or more
    GPUs: tools/nccl_abi_probe.py under torchrun, every rank compares with the gathered per-rank matrices."""
    import ctypes, subprocess, sys
    nccl = ctypes.CDLL("libnccl.so.2")

    class UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_char * 128)]

    uid = UniqueId()
    assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
    assert nccl.ncclCommInitRank(ctypes.byref(comm), 1, uid, 0) == 0
    meter = zb.RunningScore(21, device="cuda")
    gen = torch.Generator().manual_seed(8)
    meter.update(torch.randint(0, 21, (2, 40, 40), generator=gen).cuda(), torch.randint(0, 21, (2, 40, 40), generator=gen).cuda())
    before = meter.counts().clone()
    meter.all_reduce(nccl_comm=comm.value)
    torch.cuda.synchronize()
    assert torch.equal(meter.counts(), before) and int(before.sum()) == 2 * 40 * 40
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
    if torch.cuda.device_count() >= 2:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29547", os.path.join(root, "tools", "nccl_abi_probe.py")], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_peer_memory_allreduce(zb):
    """zutis_allreduce_hist_p2p (csrc/p2p_reduce.cu).  One GPU: a context of world size 1 through the raw C ABI (staging,
    handshake with itself, sum, epochs across repeated calls, in place and out of place).  Two or more GPUs:
    tools/p2p_reduce_probe.py under torchrun compares with NCCL on every rank, eagerly and as a replayed CUDA graph."""
    import ctypes as C, subprocess, sys
    from zutis_b200 import _ffi
    handle = (C.c_ubyte * 64)(); ctx = C.c_int(-1)
    _ffi.call("zutis_p2p_create", 1, 0, 81 * 81, C.addressof(handle), C.addressof(ctx))
    _ffi.call("zutis_p2p_connect", ctx.value, C.addressof(handle))
    gen = torch.Generator().manual_seed(12)
    stream = torch.cuda.current_stream().cuda_stream
    for rep in range(5):
        hist = torch.randint(0, 1 << 50, (81 * 81,), generator=gen).cuda()
        out = torch.zeros_like(hist)
        _ffi.call("zutis_allreduce_hist_p2p", ctx.value, hist.data_ptr(), hist.numel(), out.data_ptr(), stream)
        assert torch.equal(out, hist)
        keep = hist.clone()
        _ffi.call("zutis_allreduce_hist_p2p", ctx.value, hist.data_ptr(), 21 * 21, hist.data_ptr(), stream)      # in place, smaller
        assert torch.equal(hist, keep)
    with pytest.raises(zb.ZutisBadArgument):
        _ffi.call("zutis_allreduce_hist_p2p", ctx.value, hist.data_ptr(), 2 * 81 * 81, hist.data_ptr(), stream)     # does not fit the context
    torch.cuda.synchronize()
    _ffi.call("zutis_p2p_destroy", ctx.value)
    if torch.cuda.device_count() >= 2:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29549", os.path.join(root, "tools", "p2p_reduce_probe.py")], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_image_to_text_space_drop_in(zb, golden):
    """projection + joint layer norm + per-pixel L2 norm (zutis.py:319-322)."""
    g = golden("text_space")
    dec = zb.ZutisDecoder(torch.zeros(2, 32).cuda())
    dec.clip_arch = "ViT-B/16"
    for key, ln in (("out_ln", True), ("out_noln", False)):
        got = dec.image_to_text_space(dev(g["tokens"]), dev(g["proj"]), channel_last=True, layer_norm=ln)
        assert tuple(got.shape) == g[key].shape and got.is_contiguous()
        np.testing.assert_allclose(got.cpu().numpy(), g[key], rtol=0, atol=2e-6)
    with pytest.raises(NotImplementedError):
        dec.image_to_text_space(dev(g["tokens"]), dev(g["proj"]), channel_last=False)
    # BASELINE cfg2 shape: 768 -> 512 projection of 64 x 40 x 40 tokens; unit rows, and the chain into the decoder
    gen = torch.Generator(device="cuda").manual_seed(1)
    tok = torch.randn(8, 40, 40, 768, device="cuda", generator=gen)
    proj = torch.randn(768, 512, device="cuda", generator=gen) * 0.03
    got = dec.image_to_text_space(tok, proj, channel_last=True)
    ref = O.torch_image_to_text_space(tok[:2].cpu(), proj.cpu())
    np.testing.assert_allclose(got[:2].cpu().numpy(), ref.numpy(), rtol=0, atol=3e-6)
    assert float((got.norm(dim=-1) - 1).abs().max()) < 1e-5


def test_polling_scores_every_batch_does_not_serialise_the_stream(zb):
    """trainer.py:178 / :348 read the scores after EVERY batch.  With get_scores_async + poll_scores the host never waits:
    after enqueueing all batches the stream must still be busy, the polled scores are those of an earlier batch (or
    None), and once everything has landed they equal the blocking get_scores() bit for bit."""
    import time
    gen = torch.Generator().manual_seed(2)
    B, Q, D, h, w, H, W = 32, 81, 512, 40, 40, 320, 320
    text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).cuda()
    coarse = torch.randn(B, D, h // 2, w // 2, generator=gen)
    tokens = torch.nn.functional.normalize(torch.nn.functional.interpolate(coarse, scale_factor=2, mode="bilinear").permute(0, 2, 3, 1), dim=-1).contiguous().cuda()
    gt = torch.randint(0, Q, (B, H, W), generator=gen).cuda()
    meter = zb.RunningScore(Q)
    zb.decode_and_score(text, tokens, gt, (H, W), meter)           # warm-up: workspace, prepared operand, pinned buffer
    meter.get_scores_async(); torch.cuda.synchronize(); assert meter.poll_scores() is not None
    meter.reset()
    assert meter.poll_scores() is None
    n_batches, seen = 40, []
    torch.cuda.synchronize()
    # keep the device busy for ~60 ms first: every call below is then enqueued behind a running kernel, and anything that
    # waited for the device would make the host loop at least that long
    clock_khz = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1_900_000
    t0 = time.perf_counter()
    torch.cuda._sleep(int(0.06 * clock_khz * 1e3))
    for _ in range(n_batches):
        zb.decode_and_score(text, tokens, gt, (H, W), meter)
        meter.get_scores_async()
        seen.append(meter.poll_scores())
    host_s = time.perf_counter() - t0
    still_busy = not torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    assert still_busy and host_s < 0.05, f"the host loop waited for the device (host {host_s * 1e3:.1f} ms, device done at {gpu_s * 1e3:.1f} ms)"
    assert all(x is None for x in seen)                              # nothing can land while the device is still asleep
    meter.get_scores_async(); torch.cuda.synchronize()
    polled = meter.poll_scores()
    blocking = meter.get_scores()
    assert polled is not None and polled[0] == blocking[0] and polled[1] == blocking[1]
    assert int(meter.counts().sum()) == n_batches * B * H * W


@pytest.mark.parametrize("shape", [(2, 100, 15, 20, 512), (3, 37, 9, 7, 96), (1, 100, 60, 80, 512)])
def test_masked_average_on_tensor_cores_matches_simt_and_reference(zb, shape):
    """zutis_instance_lowres_stats_ws computes the masked average (zutis.py:404-406) as a mask x tokens contraction on the
    tcgen05 kernel; sizes and probability sums are the SIMT kernel's, the averages agree with it and with the torch-CPU
    restatement of the reference's 5-D broadcast to fp32 summation-order accuracy (incl. h*w not a multiple of 32)."""
    from zutis_b200 import _ffi
    B, Q, h, w, D = shape
    gen = torch.Generator().manual_seed(B * 100 + Q)
    probs = torch.sigmoid(2.0 * torch.randn(B, Q, h, w, generator=gen))
    probs[0, 0] = 0.1                                                  # an empty mask: 0 / 1e-7 = 0
    tokens = torch.nn.functional.normalize(torch.randn(B, h, w, D, generator=gen), dim=-1)
    pc, tc = probs.cuda(), tokens.cuda()
    sizes, psum, mean = zb.ops.instance_lowres_stats(pc, tc, 0.5)                     # workspace path
    s2 = torch.empty_like(sizes); p2 = torch.empty_like(psum); m2 = torch.empty_like(mean)
    _ffi.call("zutis_instance_lowres_stats", pc.data_ptr(), *pc.stride(), tc.data_ptr(), B, Q, h, w, D, 0.5,
              s2.data_ptr(), p2.data_ptr(), m2.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(sizes, s2) and torch.equal(psum, p2)
    mask = (probs > 0.5)
    want = (tokens[:, None] * mask[..., None]).sum(dim=(2, 3)) / (mask.sum(dim=(2, 3)).float()[..., None] + 1e-7)   # the reference's expression
    scale = float(want.abs().max())
    assert float((mean.cpu() - want).abs().max()) <= 2e-6 * max(scale, 1.0)
    assert float((m2.cpu() - want).abs().max()) <= 2e-6 * max(scale, 1.0)
    assert float(mean[0, 0].abs().max()) == 0.0
