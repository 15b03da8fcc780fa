"""End-to-end predict(mask_type="instance") at BASELINE config 5 (B=16, 100 queries, 60x80 -> 480x640): wall time per
image and a cProfile of the host side."""
import sys, os, time, cProfile, pstats, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zutis_b200
from zutis_b200 import decode
B, Q, h, w, H, W = 16, 100, 60, 80, 480, 640
g = torch.Generator(device="cuda").manual_seed(5)
# blobby proposals: low-frequency random fields through a sigmoid, ~half of the queries nearly empty
coarse = torch.randn(B, Q, 8, 10, device="cuda", generator=g) * 3 - 2.0
probs = torch.sigmoid(torch.nn.functional.interpolate(coarse, size=(h, w), mode="bicubic"))
probs = probs.clamp(0, 1).contiguous()
tokens = torch.nn.functional.normalize(torch.randn(B, h, w, 512, device="cuda", generator=g), dim=-1)
text = torch.nn.functional.normalize(torch.randn(81, 512, device="cuda", generator=g), dim=-1)
model = types.SimpleNamespace(text_embeddings=text)
out = {"mask_proposals": probs, "patch_tokens": tokens}
for nms in ("hard", None):
    preds = decode.predict(model, out, "instance", size=(H, W), nms_type=nms)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        preds = decode.predict(model, out, "instance", size=(H, W), nms_type=nms)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"nms={nms}: {dt * 1e3:.1f} ms per batch of {B} -> {B / dt:.1f} images/s, {len(preds)} predictions")
pr = cProfile.Profile()
pr.enable()
decode.predict(model, out, "instance", size=(H, W), nms_type="hard")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
