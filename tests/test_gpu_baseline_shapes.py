"""north_star parity bars at BASELINE.json's shapes, GPU path vs the CPU oracle on IDENTICAL tensors.

Inputs are generated on the CPU with bench.py's model-like generator (seeded), copied to the GPU, and pushed through
the public API (`ZutisDecoder.predict`, `decode_and_score`); the oracle (torch-CPU restatement of zutis.py:355-372 +
running_score.py) sees the same arrays.  Asserted per config:
  * pixel-label agreement >= 99.99 %, and every disagreeing pixel is a near-tie of the reference's full-resolution logits
  * |mIoU difference| <= 1e-4
  * the device confusion matrix equals the reference's `_fast_hist` (running_score.py:11-16) fed the device's labels,
    element for element (bit-exact wherever labels agree)
  * low-res logits within 1e-5 of max|logit|
  * the C oracle's decode of the device's own low-res logits reproduces the device labels bit for bit on a full image,
    through both decode kernels (cell kernel and tiled brute force), including Q = 920.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb(cuda_device):
    import zutis_b200
    from zutis_b200 import _ffi
    _ffi.check(_ffi.lib().zutis_device_check(0))
    return zutis_b200


def near_ties_only(text, tokens, size, ref_labels, got_labels, ulps=64.0):
    """Every pixel where the label maps differ must be a near-tie of the REFERENCE's full-resolution fp32 logits
    (re-derived per image so that cfg4's 739 MB/image tensor exists once)."""
    worst = 0.0
    for b in np.unique(np.argwhere(ref_labels != got_labels)[:, 0]):
        full = O.torch_semantic_predict(text, tokens[b:b + 1], size, return_logits=True)[0].numpy()     # [Q,H,W]
        scale = np.abs(full).max() * np.finfo(np.float32).eps
        ys, xs = np.nonzero(ref_labels[b] != got_labels[b])
        gap = np.abs(full[ref_labels[b, ys, xs], ys, xs] - full[got_labels[b, ys, xs], ys, xs]) / scale
        worst = max(worst, float(gap.max()))
    return worst <= ulps, worst


@pytest.mark.parametrize("cfg,n_img,tokens_mode", [("cfg1", 2, False), ("cfg2", 8, False), ("cfg3", 4, False), ("cfg4", 2, False),
                                                   ("cfg2", 4, True), ("cfg2", 4, "segmented")])
def test_label_agreement_and_scores_vs_oracle(zb, cfg, n_img, tokens_mode):
    import bench
    from zutis_b200 import _ffi
    c = dict(bench.WORKLOADS[cfg], B=n_img)
    Q, H, W = c["Q"], c["H"], c["W"]
    text, tokens, gt = bench.make_inputs_torch(c, "cpu", 7, tokens_mode)
    gt_np = gt.numpy()

    # ---- reference arithmetic (oracle) on the CPU
    ref_labels = O.torch_semantic_predict(text, tokens, (H, W))
    ref_meter = O.OracleRunningScore(Q)
    ref_meter.update(gt_np, ref_labels)
    ref_scores, _ = ref_meter.get_scores()

    # ---- device path through the public API, same tensors
    dec = zb.ZutisDecoder(text.cuda())
    got_labels = dec.predict({"patch_tokens": tokens.cuda()}, "semantic", size=(H, W))
    assert got_labels.dtype == np.int64 and got_labels.shape == ref_labels.shape
    meter = zb.RunningScore(Q)
    dev_labels = dec.decode_and_score(tokens.cuda(), gt.cuda(), (H, W), meter, want_labels=True)
    assert np.array_equal(dev_labels.cpu().numpy().astype(np.int64), got_labels)       # both entry points agree
    scores, _ = meter.get_scores()

    agreement = float((got_labels == ref_labels).mean())
    ok, worst = near_ties_only(text, tokens, (H, W), ref_labels, got_labels)
    print(f"{cfg} tokens={tokens_mode}: agreement {agreement:.6f}, worst tie gap {worst:.1f} ulp, "
          f"mIoU {scores['Mean IoU']:.6f} vs {ref_scores['Mean IoU']:.6f}")
    assert agreement >= 0.9999, f"{cfg}: label agreement {agreement}"
    assert ok, f"{cfg}: a disagreeing pixel is {worst} ulp(max|logit|) apart in the reference's logits"
    assert abs(scores["Mean IoU"] - ref_scores["Mean IoU"]) <= 1e-4
    for k in ("Pixel Acc", "Mean Acc", "FreqW Acc"):
        assert abs(scores[k] - ref_scores[k]) <= 1e-4, k
    # confusion matrix: the reference's own counting (running_score.py:11-16) fed the DEVICE labels
    want_hist = np.zeros((Q, Q), np.int64)
    for lt, lp in zip(gt_np, got_labels):
        m = (lt >= 0) & (lt < Q)
        want_hist += np.bincount(Q * lt[m].astype(int) + lp[m], minlength=Q * Q).reshape(Q, Q)
    assert np.array_equal(meter.counts().cpu().numpy(), want_hist)
    assert np.array_equal(meter.confusion_matrix, want_hist.astype(np.float64))

    # ---- low-res logits bar
    low = zb.ops.contraction(text.cuda(), tokens.cuda())
    ref_low = O.torch_lowres_logits(text, tokens).numpy()
    assert np.abs(low.cpu().numpy() - ref_low).max() / np.abs(ref_low).max() <= 1e-5

    # ---- C oracle on one FULL image of the device's own logits: both decode kernels, bit for bit (Q = 920 included)
    one = low[:1]
    want = O.c_decode_semantic(one.cpu().numpy(), (H, W))
    ws = zb.ops.DecodeWorkspace()
    for kwargs in (dict(workspace=ws), dict(mode=_ffi.DECODE_TILED), dict(mode=_ffi.DECODE_CELLS, workspace=zb.ops.DecodeWorkspace())):
        got = zb.ops.decode_score(one, (H, W), **kwargs)
        assert np.array_equal(got.cpu().numpy().astype(np.int64), want), f"{cfg}: decode {kwargs} differs from the C oracle"


def test_text_embedding_swap_is_noticed(zb):
    """update_text_embeddings (zutis.py:333-338) replaces the tensor by another of the same size; the caching allocator
    hands the freed address back.  The prepared tensor-core operand must follow the new embeddings."""
    gen = torch.Generator().manual_seed(5)
    tokens = torch.nn.functional.normalize(torch.randn(2, 20, 20, 512, generator=gen), dim=-1).cuda()
    dec = zb.ZutisDecoder(torch.nn.functional.normalize(torch.randn(81, 512, generator=gen), dim=-1).cuda())
    first = dec.predict({"patch_tokens": tokens}, "semantic", size=(160, 160))
    old_ptr = dec.text_embeddings.data_ptr()
    new_text = torch.nn.functional.normalize(torch.randn(81, 512, generator=gen), dim=-1)
    dec.text_embeddings = None                       # frees the old tensor ...
    dec.text_embeddings = new_text.cuda()            # ... and the allocator may return the same address
    second = dec.predict({"patch_tokens": tokens}, "semantic", size=(160, 160))
    want = O.torch_semantic_predict(new_text, tokens.cpu(), (160, 160))
    assert (second == want).mean() >= 0.9999, f"stale text operand (same address: {dec.text_embeddings.data_ptr() == old_ptr})"
    assert (first != second).mean() > 0.5


@pytest.mark.parametrize("value", [float("inf"), float("-inf"), float("nan")])
def test_nonfinite_tokens_follow_the_reference(zb, value):
    """A +-inf / NaN token reaches the labels the way torch's einsum + argmax treat it (zutis.py:361-372)."""
    gen = torch.Generator().manual_seed(11)
    text = torch.nn.functional.normalize(torch.randn(81, 512, generator=gen), dim=-1)
    tokens = torch.nn.functional.normalize(torch.randn(2, 10, 12, 512, generator=gen), dim=-1)
    tokens[1, 4, 5, 17] = value
    ref_low = O.torch_lowres_logits(text, tokens).numpy()
    low = zb.ops.contraction(text.cuda(), tokens.cuda()).cpu().numpy()
    bad = ~np.isfinite(ref_low)
    assert bad.any()
    assert np.array_equal(np.isnan(low), np.isnan(ref_low)), "NaN pattern of the low-res logits differs"
    assert np.array_equal(low[bad & ~np.isnan(ref_low)], ref_low[bad & ~np.isnan(ref_low)]), "+-inf logits differ"
    assert np.abs(low[~bad] - ref_low[~bad]).max() / np.abs(ref_low[~bad]).max() <= 1e-5
    want = O.torch_semantic_predict(text, tokens, (80, 96))
    got = zb.ZutisDecoder(text.cuda()).predict({"patch_tokens": tokens.cuda()}, "semantic", size=(80, 96))
    assert (got == want).mean() >= 0.9999
    assert np.array_equal(got[1, 32:48, 40:56], want[1, 32:48, 40:56])      # around the poisoned pixel
