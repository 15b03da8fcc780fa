"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (it needs the read-only checkout at /root/reference):

    python tests/golden/make_golden.py

The reference (NoelShin/zutis) ships no tests or golden vectors for the mask-decode +
scoring path, so the oracle in oracle/ is pinned against outputs of the reference's own
code captured here: ``ZUTIS.predict`` (networks/zutis.py:340-470), ``ZUTIS.forward``
(:472-532) with random-init weights, ``RunningScore`` (utils/running_score.py) and
``compute_iou`` (utils/iou.py).  Two third-party imports of networks/zutis.py are absent
offline and are stubbed exactly as SURVEY.md Appendix A.6 describes: ``clip`` (load ->
random-init networks.clip_arch.CLIP, tokenize -> deterministic tokens) and
``pycocotools.mask.encode`` (returns the mask it was given, so fixtures can store it).
Nothing from the reference is copied into this repository; only its outputs are stored.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
OUT = os.environ.get("ZUTIS_GOLDEN_OUT", os.path.dirname(os.path.abspath(__file__)))


def install_stubs():
    sys.path.insert(0, REF)
    import networks.clip_arch as clip_arch

    hp = {
        "ViT-B/32": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
                         vision_patch_size=32, context_length=77, vocab_size=49408,
                         transformer_width=512, transformer_heads=8, transformer_layers=12),
        "ViT-B/16": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
                         vision_patch_size=16, context_length=77, vocab_size=49408,
                         transformer_width=512, transformer_heads=8, transformer_layers=12),
    }

    def load(name, device="cpu"):
        torch.manual_seed(0)
        return clip_arch.CLIP(**hp[name]).float().eval(), None

    def tokenize(texts):
        toks = torch.zeros(len(texts), 77, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [49406] + [1 + (sum(map(ord, w)) * 31 + 7 * k) % 49000 for k, w in enumerate(t.split())][:75]
            toks[i, :len(ids)] = torch.tensor(ids)
            toks[i, len(ids)] = 49407
        return toks

    clip = types.ModuleType("clip"); clip.load = load; clip.tokenize = tokenize
    sys.modules["clip"] = clip
    pkg = types.ModuleType("pycocotools"); mask = types.ModuleType("pycocotools.mask")
    mask.encode = lambda m: np.asarray(m)
    pkg.mask = mask
    sys.modules["pycocotools"] = pkg; sys.modules["pycocotools.mask"] = mask


def unit(x, dim=-1):
    return x / x.norm(dim=dim, keepdim=True)


def main():
    install_stubs()
    from networks.zutis import ZUTIS
    from utils.running_score import RunningScore
    from utils.iou import compute_iou

    def ref_semantic(text, tokens, size, return_logits=False):
        return ZUTIS.predict(SimpleNamespace(text_embeddings=text), {"patch_tokens": tokens}, "semantic",
                             size=size, return_logits=return_logits)

    # ---------------------------------------------------------------- 1. model-like, cfg-1
    cats = ["background"] + [f"category number {i}" for i in range(80)]
    net = ZUTIS(categories=cats, clip_arch="ViT-B/32", device=torch.device("cpu")).eval()
    torch.manual_seed(0)
    images = torch.randn(2, 3, 224, 224)
    with torch.no_grad():
        out = net(images)
    text = net.text_embeddings.clone()
    tokens = out["patch_tokens"].contiguous()                 # [2,14,14,512]
    proposals = out["mask_proposals"][:, -1].contiguous()     # [2,100,14,14], last decoder layer
    lo = ref_semantic(text, tokens, None, return_logits=True).contiguous()
    labels = ref_semantic(text, tokens, (224, 224))
    labels_lo = ref_semantic(text, tokens, None)
    g = torch.Generator().manual_seed(1)
    gt = torch.randint(0, 81, (2, 224, 224), generator=g).numpy()
    gt[:, :7] = 255
    gt[1, 100:120, 30:90] = -1
    meter = RunningScore(81)
    meter.update(gt, labels)
    s, c = meter.get_scores()
    # instance branch, hard NMS and no NMS (stub encode returns the mask itself)
    inst = {}
    for tag, nms in (("hard", "hard"), ("none", None)):
        preds = net.predict({"mask_proposals": out["mask_proposals"], "patch_tokens": out["patch_tokens"]},
                            "instance", size=(224, 224), image_ids=[11, 22], nms_type=nms)
        inst[f"inst_{tag}_category"] = np.array([p["category_id"] for p in preds], np.int64)
        inst[f"inst_{tag}_score"] = np.array([p["score"] for p in preds], np.float64)
        inst[f"inst_{tag}_image_id"] = np.array([p["image_id"] for p in preds], np.int64)
        inst[f"inst_{tag}_bbox"] = np.array([p["bbox"] for p in preds], np.float64).reshape(-1, 4)
        inst[f"inst_{tag}_mask_bits"] = np.packbits(
            np.stack([np.ascontiguousarray(p["segmentation"]) for p in preds]).astype(bool).reshape(len(preds), -1),
            axis=1) if preds else np.zeros((0, 224 * 224 // 8), np.uint8)
    np.savez_compressed(
        os.path.join(OUT, "model_cfg1.npz"),
        text=text.numpy(), tokens=tokens.numpy(), proposals=proposals.numpy(),
        lowres_logits=lo.numpy(), labels=labels.astype(np.int16), labels_lowres=labels_lo.astype(np.int16),
        gt=gt.astype(np.int16), confusion=meter.confusion_matrix.astype(np.int64),
        scores=np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]]),
        class_iou=np.array([c[i] for i in range(81)]), **inst)

    # ------------------------------------------------ 2. synthetic decode cases (small D)
    cases = {}
    specs = {
        "int8x": (2, 21, 32, 10, 12, 80, 96),        # integer scale x8, non-square
        "nonint": (1, 21, 32, 13, 17, 97, 131),      # non-integer scale (427x640-style)
        "x16": (1, 9, 16, 7, 7, 112, 112),           # ViT-B/32 style x16
        "same": (1, 5, 16, 8, 9, 8, 9),              # h == H: ATen copy branch
        "down": (1, 6, 16, 12, 10, 5, 7),            # size smaller than the grid
        "wideq": (1, 300, 24, 6, 5, 44, 37),         # wide Q
    }
    for case_index, (name, (B, Q, D, h, w, H, W)) in enumerate(specs.items()):
        gen = torch.Generator().manual_seed(100 + case_index)
        t = unit(torch.randn(Q, D, generator=gen))
        x = unit(torch.randn(B, h, w, D, generator=gen))
        cases[f"{name}_text"] = t.numpy()
        cases[f"{name}_tokens"] = x.numpy()
        cases[f"{name}_size"] = np.array([H, W])
        cases[f"{name}_lowres"] = ref_semantic(t, x, None, return_logits=True).contiguous().numpy()
        # full-resolution logits (return_logits=True, zutis.py:369-370): first 4 categories only, to keep
        # the fixture small; they pin the bilinear arithmetic bit-for-bit
        cases[f"{name}_full4"] = ref_semantic(t, x, (H, W), return_logits=True)[:, :4].contiguous().numpy()
        cases[f"{name}_labels"] = ref_semantic(t, x, (H, W)).astype(np.int16)
    # exact ties (duplicate category rows -> first index wins) and NaN (counts as max)
    gen = torch.Generator().manual_seed(5)
    t = unit(torch.randn(6, 8, generator=gen)); t[4] = t[1]; t[5] = t[1]
    x = unit(torch.randn(1, 5, 5, 8, generator=gen))
    cases["ties_text"] = t.numpy(); cases["ties_tokens"] = x.numpy(); cases["ties_size"] = np.array([40, 40])
    cases["ties_labels"] = ref_semantic(t, x, (40, 40)).astype(np.int16)
    cases["ties_lowres"] = ref_semantic(t, x, None, return_logits=True).contiguous().numpy()
    xn = x.clone(); xn[0, 2, 2, :] = float("nan")
    cases["nan_text"] = t.numpy(); cases["nan_tokens"] = xn.numpy(); cases["nan_size"] = np.array([40, 40])
    cases["nan_labels"] = ref_semantic(t, xn, (40, 40)).astype(np.int16)
    cases["nan_lowres"] = ref_semantic(t, xn, None, return_logits=True).contiguous().numpy()
    # size given as a pair of 1-element tensors (trainer.py:322-323)
    cases["tensorsize_labels"] = ref_semantic(
        torch.from_numpy(cases["nonint_text"]), torch.from_numpy(cases["nonint_tokens"]),
        [torch.tensor([97]), torch.tensor([131])]).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "decode_cases.npz"), **cases)

    # ------------------------------------------------------------- 3. scoring known answers
    sc = {}
    gt3 = np.array([[0, 0, 1, 1], [2, 2, 255, 1], [0, 1, 2, -1]])
    pr3 = np.array([[0, 1, 1, 1], [2, 0, 2, 2], [0, 1, 1, 0]])
    for n in (3, 4):
        m = RunningScore(n); m.update(gt3[None], pr3[None]); s, c = m.get_scores()
        sc[f"ka{n}_confusion"] = m.confusion_matrix
        sc[f"ka{n}_scores"] = np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]])
        sc[f"ka{n}_class_iou"] = np.array([c[i] for i in range(n)])
    sc["ka_gt"] = gt3; sc["ka_pred"] = pr3
    rng = np.random.default_rng(7)
    gtr = rng.integers(0, 920, (3, 61, 47)); gtr[0, :5] = 1000; gtr[2, 3, :] = -3
    prr = rng.integers(0, 920, (3, 61, 47))
    m = RunningScore(920); m.update(gtr, prr); m.update(gtr[:1], prr[:1]); s, c = m.get_scores()
    sc["wide_gt"] = gtr.astype(np.int16); sc["wide_pred"] = prr.astype(np.int16)
    sc["wide_confusion_nz"] = np.stack(np.nonzero(m.confusion_matrix) + (m.confusion_matrix[np.nonzero(m.confusion_matrix)],))
    sc["wide_scores"] = np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]])
    sc["wide_class_iou"] = np.array([c[i] for i in range(920)])
    m = RunningScore(5); s, c = m.get_scores()              # empty matrix -> nan/nan/0/nan
    sc["empty_scores"] = np.array([s["Pixel Acc"], s["Mean Acc"], s["FreqW Acc"], s["Mean IoU"]])
    # ragged list of differently sized images (running_score.py:19 zips over the first dim)
    ga = rng.integers(0, 7, (9, 11)); gb = rng.integers(0, 7, (5, 4)); gb[0, 0] = 255
    pa = rng.integers(0, 7, (9, 11)); pb = rng.integers(0, 7, (5, 4))
    m = RunningScore(7); m.update([ga, gb], [pa, pb])
    sc["rag_ga"] = ga; sc["rag_gb"] = gb; sc["rag_pa"] = pa; sc["rag_pb"] = pb
    sc["rag_confusion"] = m.confusion_matrix
    # compute_iou known answers (iou.py)
    a = np.array([[1, 1, 0], [0, 1, 0]], bool); b = np.array([[1, 0, 0], [0, 1, 1]], bool)
    pf = np.array([[.6, .4, .9], [.1, .7, .2]])
    sc["iou_bool"] = np.array(compute_iou(a, b))
    sc["iou_thr"] = np.array(compute_iou(pf, b, threshold=0.5))
    sc["iou_empty"] = np.array(compute_iou(np.zeros((2, 3), bool), np.zeros((2, 3), bool)))
    sc["iou_torch"] = compute_iou(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    ma = rng.random((37, 53)) > 0.6; mb = rng.random((37, 53)) > 0.5
    sc["iou_rand_a"] = ma; sc["iou_rand_b"] = mb; sc["iou_rand"] = np.array(compute_iou(ma, mb))
    np.savez_compressed(os.path.join(OUT, "scoring.npz"), **sc)

    # -------------------------------------------- 4. instance decode on synthetic proposals
    ic = {}
    gen = torch.Generator().manual_seed(9)
    t = unit(torch.randn(12, 32, generator=gen))
    x = unit(torch.randn(2, 9, 11, 32, generator=gen))
    base = torch.randn(2, 3, 20, 5, 6, generator=gen)
    up = torch.nn.functional.interpolate(base.flatten(0, 1), size=(9, 11), mode="bilinear").view(2, 3, 20, 9, 11)
    mp = torch.sigmoid(3 * up)                                 # 5-D: the last layer is taken (:379-382)
    self_ns = SimpleNamespace(text_embeddings=t)
    self_ns.non_maximum_suppression = types.MethodType(ZUTIS.non_maximum_suppression, self_ns)
    for tag, nms in (("hard", "hard"), ("none", None), ("linear", "linear"), ("gaussian", "gaussian")):
        preds = ZUTIS.predict(self_ns, {"mask_proposals": mp, "patch_tokens": x}, "instance",
                              size=(70, 90), image_ids=[5, 6], nms_type=nms)
        ic[f"{tag}_category"] = np.array([p["category_id"] for p in preds], np.int64)
        ic[f"{tag}_score"] = np.array([p["score"] for p in preds], np.float64)
        ic[f"{tag}_image_id"] = np.array([p["image_id"] for p in preds], np.int64)
        ic[f"{tag}_bbox"] = np.array([p["bbox"] for p in preds], np.float64).reshape(-1, 4)
        ic[f"{tag}_mask_bits"] = np.packbits(
            np.stack([np.ascontiguousarray(p["segmentation"]) for p in preds]).astype(bool).reshape(len(preds), -1), axis=1)
    ic["text"] = t.numpy(); ic["tokens"] = x.numpy(); ic["proposals"] = mp.numpy(); ic["size"] = np.array([70, 90])
    np.savez_compressed(os.path.join(OUT, "instance_cases.npz"), **ic)
    # ------------------------------------------- 5. image_to_text_space (zutis.py:301-331)
    ts = {}
    gen = torch.Generator().manual_seed(13)
    pt = torch.randn(2, 5, 6, 64, generator=gen) * 3 + 0.5
    pj = torch.randn(64, 32, generator=gen) * 0.2
    ns = SimpleNamespace(clip_arch="ViT-B/16")
    ts["tokens"] = pt.numpy(); ts["proj"] = pj.numpy()
    ts["out_ln"] = ZUTIS.image_to_text_space(ns, pt, pj, channel_last=True, layer_norm=True).numpy()
    ts["out_noln"] = ZUTIS.image_to_text_space(ns, pt, pj, channel_last=True, layer_norm=False).numpy()
    np.savez_compressed(os.path.join(OUT, "text_space.npz"), **ts)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
