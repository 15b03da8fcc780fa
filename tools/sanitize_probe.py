"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel once at modest sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zutis_b200
from zutis_b200 import ops
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "model_cfg1.npz"))
text = torch.from_numpy(g["text"]).cuda(); tokens = torch.from_numpy(g["tokens"]).cuda(); gt = torch.from_numpy(g["gt"].astype(np.int64)).cuda()
tok_big = torch.nn.functional.normalize(torch.randn(14, 40, 40, 512, device="cuda"), dim=-1)      # 182 tiles > 148 SMs
lo = ops.contraction(text, tok_big)                         # tcgen05, multi-round
lo2 = ops.contraction(text, tokens, precision="fp32")      # SIMT
meter = zutis_b200.RunningScore(81)
labels = zutis_b200.decode_and_score(text, tokens, gt, (224, 224), meter, want_labels=True)
ops.decode_score(lo[:2], (320, 320))                        # AUTO: the cell kernel (TMA-staged taps)
from zutis_b200 import _ffi
coarse = torch.randn(3, 81, 6, 6, device="cuda")
smooth = torch.nn.functional.interpolate(coarse, size=(20, 20), mode="bilinear")
pm = torch.zeros(3, 20, 20, 84, device="cuda"); pm[..., :81] = smooth.permute(0, 2, 3, 1)
part = torch.zeros(81 * 81, dtype=torch.int32, device="cuda")
ops.decode_score(pm[..., :81].permute(0, 3, 1, 2), (160, 160), gt=torch.randint(0, 81, (3, 160, 160), device="cuda"),
                 hist_partial=part, mode=_ffi.DECODE_CELLS)   # cell kernel forced, with histogram
ws = ops.DecodeWorkspace()
lo3 = ops.contraction(text, tok_big[:3])
ops.decode_score(lo3, (320, 320), workspace=ws)
wide = torch.nn.functional.interpolate(torch.randn(1, 300, 4, 4, device="cuda"), size=(9, 10), mode="bilinear")
pw = torch.zeros(1, 9, 10, 300, device="cuda"); pw.copy_(wide.permute(0, 2, 3, 1))
ops.decode_score(pw.permute(0, 3, 1, 2), (72, 80), mode=_ffi.DECODE_CELLS)     # wide Q: taps from global memory
ops.decode_score(lo[:1], (320, 320), mode=_ffi.DECODE_TILED)
ops.decode_score(lo2, (100, 90), mode=1)
meter.update(gt, labels); meter.get_scores()
probs = torch.sigmoid(3 * torch.randn(2, 100, 15, 20, device="cuda"))
bits, areas = ops.decode_threshold(probs, (120, 160), 0.5)
ops.pairwise_mask_intersections(bits[0]); ops.unpack_mask_bits(bits[0], 160)
ops.mask_rle(bits, 160); ops.mask_rle_strings(bits, 160, mask_ids=torch.tensor([3, 0, 5]))
s, p, m = ops.instance_lowres_stats(probs, torch.randn(2, 15, 20, 512, device="cuda"), 0.5)
ops.instance_categories(m, text, 5.0)
ops.upsample_bilinear(lo2, (50, 60))
torch.cuda.synchronize()
print("sanitize probe done", float(meter.get_scores()[0]["Mean IoU"]))
