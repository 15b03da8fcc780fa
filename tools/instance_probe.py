"""Time the instance-path kernels at BASELINE config 5 (B=16, 100 queries, 60x80 -> 480x640)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zutis_b200
from zutis_b200 import ops
B, Q, D, h, w, H, W = 16, 100, 768, 60, 80, 480, 640
gen = torch.Generator(device="cuda").manual_seed(5)
queries = torch.nn.functional.normalize(torch.randn(B, Q, D, device="cuda", generator=gen), dim=-1)
feats = (4 * torch.randn(B, h, w, D, device="cuda", generator=gen)).contiguous()
tokens = torch.nn.functional.normalize(torch.randn(B, h, w, 512, device="cuda", generator=gen), dim=-1)
text = torch.nn.functional.normalize(torch.randn(81, 512, device="cuda", generator=gen), dim=-1)
def timeit(name, fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:34s} {e0.elapsed_time(e1) * 1e3 / reps:10.1f} us")
    return out
probs = timeit("contraction+sigmoid (q-major)", lambda: ops.contraction(queries, feats, sigmoid=True, pixel_major=False))
probs_pm = timeit("contraction+sigmoid (pixel-major)", lambda: ops.contraction(queries, feats, sigmoid=True, pixel_major=True))
timeit("lowres stats + mean tokens", lambda: ops.instance_lowres_stats(probs, tokens, 0.5))
sizes, psum, mean = ops.instance_lowres_stats(probs, tokens, 0.5)
timeit("categories", lambda: ops.instance_categories(mean, text, 5.0))
bits, areas = timeit("threshold 480x640 (q-major probs)", lambda: ops.decode_threshold(probs, (H, W), 0.5))
timeit("threshold 480x640 (pixel-major)", lambda: ops.decode_threshold(probs_pm, (H, W), 0.5))
timeit("pairwise intersections (1 image)", lambda: ops.pairwise_mask_intersections(bits[0]))
timeit("unpack 100 masks", lambda: ops.unpack_mask_bits(bits[0], W))
