"""Pin the CPU oracle (oracle/) against outputs captured from the reference itself.

The fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which
imports NoelShin/zutis from /root/reference and runs ZUTIS.predict / RunningScore /
compute_iou verbatim.  Both restatements (plain C and torch-CPU ops) must reproduce them.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

CASES = ["int8x", "nonint", "x16", "same", "down", "wideq"]


@pytest.mark.parametrize("name", CASES)
def test_c_bilinear_bit_exact_with_reference(golden, name):
    g = golden("decode_cases")
    H, W = g[f"{name}_size"]
    full = O.c_bilinear(g[f"{name}_lowres"][:, :4], (H, W))
    ref = g[f"{name}_full4"]
    assert full.dtype == ref.dtype and full.shape == ref.shape
    if max(H, W) > 64:
        assert np.array_equal(full.view(np.int32), ref.view(np.int32)), "bilinear restatement is not bit-exact"
    else:
        # Observed in torch 2.11 CPU: when both output sides are <= 64 ATen sums the four taps as one
        # fma chain over pre-multiplied 2-D weights (w01*b, +w00*a, +w10*c, +w11*d) instead of the
        # separable width-then-height pattern it uses for every realistic image size.  The oracle
        # restates the separable pattern only; the two agree to a few ulp ("oracle").
        assert np.abs(full - ref).max() <= 4 * np.finfo(np.float32).eps * np.abs(ref).max()


@pytest.mark.parametrize("name", CASES + ["ties", "nan"])
def test_c_decode_labels_equal_reference(golden, name):
    g = golden("decode_cases")
    H, W = g[f"{name}_size"]
    labels = O.c_decode_semantic(g[f"{name}_lowres"], (H, W))
    assert np.array_equal(labels, g[f"{name}_labels"].astype(np.int64))


@pytest.mark.parametrize("name", CASES + ["ties", "nan"])
def test_torch_port_labels_equal_reference(golden, name):
    g = golden("decode_cases")
    H, W = g[f"{name}_size"]
    labels = O.torch_semantic_predict(torch.from_numpy(g[f"{name}_text"]), torch.from_numpy(g[f"{name}_tokens"]),
                                      (int(H), int(W)))
    assert np.array_equal(labels, g[f"{name}_labels"].astype(np.int64))


def test_size_as_one_element_tensors(golden):
    g = golden("decode_cases")
    labels = O.c_decode_semantic(g["nonint_lowres"], [torch.tensor([97]), torch.tensor([131])])
    assert np.array_equal(labels, g["tensorsize_labels"].astype(np.int64))
    assert np.array_equal(g["tensorsize_labels"], g["nonint_labels"])


def test_c_logits_within_tolerance(golden):
    g = golden("model_cfg1")
    lo = O.c_logits(g["text"], g["tokens"])
    ref = g["lowres_logits"]
    assert np.abs(lo - ref).max() / np.abs(ref).max() <= 1e-5


def test_model_cfg1_decode_and_score(golden):
    g = golden("model_cfg1")
    labels = O.c_decode_semantic(g["lowres_logits"], (224, 224))
    assert np.array_equal(labels, g["labels"].astype(np.int64))
    assert np.array_equal(O.c_decode_semantic(g["lowres_logits"], None), g["labels_lowres"].astype(np.int64))
    hist = O.c_fast_hist(g["gt"], labels, 81)
    assert np.array_equal(hist, g["confusion"])
    summary, cls = O.scores_from_hist(hist)
    got = np.array([summary["Pixel Acc"], summary["Mean Acc"], summary["FreqW Acc"], summary["Mean IoU"]])
    assert np.array_equal(got, g["scores"])                       # same float64 ops, same order
    assert np.array_equal(np.array([cls[i] for i in range(81)]), g["class_iou"], equal_nan=True)


@pytest.mark.parametrize("n", [3, 4])
def test_running_score_known_answers(golden, n):
    g = golden("scoring")
    hist = O.c_fast_hist(g["ka_gt"], g["ka_pred"], n)
    assert np.array_equal(hist, g[f"ka{n}_confusion"].astype(np.int64))
    if n == 3:
        assert hist.tolist() == [[2, 1, 0], [0, 3, 1], [1, 1, 1]]
    meter = O.OracleRunningScore(n)
    meter.update(g["ka_gt"][None], g["ka_pred"][None])
    assert np.array_equal(meter.confusion_matrix, g[f"ka{n}_confusion"])
    s, c = meter.get_scores()
    assert np.array_equal(np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]]), g[f"ka{n}_scores"])
    assert np.array_equal(np.array([c[i] for i in range(n)]), g[f"ka{n}_class_iou"], equal_nan=True)
    assert s["Pixel Acc"] == 0.6 and abs(s["Mean IoU"] - 0.41666666) < 1e-6 and abs(s["FreqW Acc"] - 0.425) < 1e-12


def test_running_score_wide_ragged_empty(golden):
    g = golden("scoring")
    gt, pr = g["wide_gt"].astype(np.int64), g["wide_pred"].astype(np.int64)
    hist = O.c_fast_hist(gt, pr, 920)
    O.c_fast_hist(gt[:1], pr[:1], 920, hist)
    ref = np.zeros((920, 920), np.int64)
    r, c, v = g["wide_confusion_nz"]
    ref[r.astype(int), c.astype(int)] = v.astype(np.int64)
    assert np.array_equal(hist, ref)
    s, cls = O.scores_from_hist(hist)
    assert np.array_equal(np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]]), g["wide_scores"])
    assert np.array_equal(np.array([cls[i] for i in range(920)]), g["wide_class_iou"], equal_nan=True)
    m = O.OracleRunningScore(7)
    m.update([g["rag_ga"], g["rag_gb"]], [g["rag_pa"], g["rag_pb"]])
    assert np.array_equal(m.confusion_matrix, g["rag_confusion"])
    s, _ = O.OracleRunningScore(5).get_scores()
    e = g["empty_scores"]
    assert np.isnan(s["Pixel Acc"]) and np.isnan(s["Mean Acc"]) and s["FreqW Acc"] == 0.0 and np.isnan(s["Mean IoU"])
    assert np.isnan(e[0]) and np.isnan(e[1]) and e[2] == 0.0 and np.isnan(e[3])


def test_mask_iou_known_answers(golden):
    g = golden("scoring")
    a = np.array([[1, 1, 0], [0, 1, 0]], bool); b = np.array([[1, 0, 0], [0, 1, 1]], bool)
    assert O.c_mask_iou(a, b) == float(g["iou_bool"]) == 0.4999999875000003
    pf = np.array([[.6, .4, .9], [.1, .7, .2]])
    assert O.c_mask_iou(pf, b, threshold=0.5) == float(g["iou_thr"])
    assert O.c_mask_iou(np.zeros((2, 3), bool), np.zeros((2, 3), bool)) == float(g["iou_empty"]) == 0.0
    assert O.c_mask_iou(g["iou_rand_a"], g["iou_rand_b"]) == float(g["iou_rand"])


def _instance_oracle(text, proposals, tokens, size, nms):
    """Instance decode through the oracle pieces: low-res stats (torch port), masks (C), NMS (numpy)."""
    conf, cat, _ = O.torch_instance_lowres(torch.from_numpy(text), torch.from_numpy(proposals), torch.from_numpy(tokens))
    mp = proposals[:, -1] if proposals.ndim == 5 else proposals
    masks = O.c_decode_threshold(mp, size, 0.5)
    out = []
    for b in range(masks.shape[0]):
        if nms:
            kept = O.hard_nms(masks[b], conf[b], cat[b], nms_type=nms)
        else:
            kept = [(int(c), i, float(s)) for i, (s, c) in enumerate(zip(conf[b], cat[b])) if c != 0 and masks[b, i].any()]
        out.append((b, kept, masks[b]))
    return out


@pytest.mark.parametrize("fixture,prefix", [("instance_cases", ""), ("model_cfg1", "inst_")])
@pytest.mark.parametrize("tag", ["hard", "none", "linear", "gaussian"])
def test_instance_decode_matches_reference(golden, fixture, prefix, tag):
    g = golden(fixture)
    if f"{prefix}{tag}_score" not in g:
        pytest.skip("soft-NMS outputs of the reference are stored for the synthetic proposals only")
    size = tuple(int(v) for v in g["size"]) if "size" in g else (224, 224)
    res = _instance_oracle(g["text"], g["proposals"], g["tokens"], size, None if tag == "none" else tag)
    cats, scores, bits = [], [], []
    for b, kept, masks in res:
        for c, i, s in kept:
            cats.append(c); scores.append(s); bits.append(np.packbits(masks[i].reshape(-1)))
    assert cats == g[f"{prefix}{tag}_category"].tolist()
    np.testing.assert_allclose(np.array(scores, np.float64).reshape(-1), g[f"{prefix}{tag}_score"], rtol=2e-6, atol=1e-9)
    ref_bits = g[f"{prefix}{tag}_mask_bits"]
    assert len(bits) == len(ref_bits)
    if bits:
        assert np.array_equal(np.stack(bits), ref_bits)


@pytest.mark.parametrize("nms_type", ["hard", "linear", "gaussian"])
def test_host_nms_replay_matches_oracle(golden, nms_type):
    """The product's host NMS (driven by intersection counts, used for soft NMS and for tied scores) against the oracle's
    mask-based loop on the reference fixture's masks and scores."""
    from zutis_b200.decode import _nms_keep
    g = golden("instance_cases")
    size = tuple(int(v) for v in g["size"])
    conf, cat, _ = O.torch_instance_lowres(torch.from_numpy(g["text"]), torch.from_numpy(g["proposals"]), torch.from_numpy(g["tokens"]))
    masks = O.c_decode_threshold(g["proposals"][:, -1], size, 0.5)
    for b in range(masks.shape[0]):
        flat = masks[b].reshape(masks.shape[1], -1).astype(np.int64)
        inter = flat @ flat.T
        got = _nms_keep(cat[b], conf[b], inter, nms_type)
        want = O.hard_nms(masks[b], conf[b], cat[b], nms_type=nms_type)
        assert [(int(c), q) for c, q, _ in got] == [(c, q) for c, q, _ in want]
        assert [float(s) for _, _, s in got] == [s for _, _, s in want]


def test_image_to_text_space_port_matches_reference(golden):
    g = golden("text_space")
    for key, ln in (("out_ln", True), ("out_noln", False)):
        got = O.torch_image_to_text_space(torch.from_numpy(g["tokens"]), torch.from_numpy(g["proj"]), ln).numpy()
        assert np.array_equal(got, g[key])


# ------------------------------------------------------------------------------- COCO RLE restatement
def test_rle_known_answers_and_round_trip():
    """Hand-derived vectors for cocoapi's rleEncode / rleToString (pycocotools itself is not installed here, so
    the compressed string is pinned by these, by the decoder round trip and by the vectorised/product
    implementations agreeing with this scalar restatement)."""
    m = np.array([[0, 1, 1], [0, 0, 1]], bool)                 # column-major: 0 0 | 1 0 | 1 1
    assert O.rle_counts(m) == [2, 1, 1, 2]
    assert O.rle_counts(np.ones((2, 2), bool)) == [0, 4]       # a mask that starts set has an empty first run
    assert O.rle_counts(np.zeros((3, 2), bool)) == [6]
    # 5-bit groups, +48: 3 -> '3'; 100 = 0b11_00100 -> (4|32)+48 = 'T', then 3 -> '3'; 40 = 0b1_01000 -> 'X','1'
    assert O.rle_to_string([3, 2, 1]) == b"321"
    assert O.rle_to_string([3, 2, 1, 2]) == b"3210"            # fourth value is stored as 2 - 2 = 0
    assert O.rle_to_string([100, 40, 7, 3]) == b"T3X17kN"      # 3 - 40 = -37 -> 27|32 -> 'k', then 30 -> 'N' (sign-extended)
    assert O.rle_to_string([307200]) == b"PP\\9"
    rng = np.random.default_rng(7)
    for _ in range(300):
        H, W = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        m = rng.random((H, W)) < rng.random()
        c = O.rle_counts(m)
        assert sum(c) == H * W and list(O.rle_counts_numpy(m)) == c
        assert O.rle_from_string(O.rle_to_string(c)) == c
    big = [0, 5, 2 ** 30, 1, 7, 2 ** 30 + 3, 1]
    assert O.rle_from_string(O.rle_to_string(big)) == big


def test_mask_to_box_matches_torchvision():
    from torchvision.ops import masks_to_boxes
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = rng.random((int(rng.integers(1, 20)), int(rng.integers(1, 20)))) < 0.3
        if not m.any():
            continue
        assert O.mask_to_box(m) == masks_to_boxes(torch.from_numpy(m)[None])[0].tolist()
