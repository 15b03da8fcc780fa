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
cat_dev, prob_dev = ops.instance_categories(m, text, 5.0)
inter = torch.stack([ops.pairwise_mask_intersections(bits[b]) for b in range(2)])
ops.instance_nms_hard(inter, cat_dev.to(torch.int32).view(2, 100), torch.rand(2, 100, device="cuda"))      # device hard NMS
# region borders: the first pruning attempt overflows, the second (virtual dominator) runs -- narrow and wide Q
for Qb in (81, 300):
    lob = 0.02 * torch.randn(2, Qb, 12, 15, device="cuda")
    region = torch.randint(0, Qb, (2, 4, 5), device="cuda").repeat_interleave(3, 1).repeat_interleave(3, 2)
    lob.scatter_add_(1, region[:, None], torch.ones(2, 1, 12, 15, device="cuda"))
    Qp = (Qb + 3) & ~3
    pb = torch.zeros(2, 12, 15, Qp, device="cuda"); pb[..., :Qb] = lob.permute(0, 2, 3, 1)
    ops.decode_score(pb[..., :Qb].permute(0, 3, 1, 2), (96, 120), mode=_ffi.DECODE_CELLS)
# host-buffer entry: chunked lanes + label narrowing threads
B2 = 20
th = torch.nn.functional.normalize(torch.randn(B2, 40, 40, 512), dim=-1).numpy(); gh = np.random.randint(0, 81, (B2, 320, 320)).astype(np.int64)
hist = np.zeros((81, 81), np.int64)
_ffi.call("zutis_semantic_eval_host", text.cpu().numpy().ctypes.data, th.ctypes.data, gh.ctypes.data, _ffi.GT_I64, B2, 81, 512, 40, 40, 320, 320,
          hist.ctypes.data, None, _ffi.GEMM_TF32X3, 0)
# peer-memory all-reduce, world size 1 (the multi-GPU exchange is exercised by tools/p2p_reduce_probe.py)
import ctypes as C
handle = (C.c_ubyte * 64)(); ctx = C.c_int(-1)
_ffi.call("zutis_p2p_create", 1, 0, 81 * 81, C.addressof(handle), C.addressof(ctx))
_ffi.call("zutis_p2p_connect", ctx.value, C.addressof(handle))
hh = torch.arange(81 * 81, device="cuda", dtype=torch.int64); oo = torch.empty_like(hh)
for _ in range(3):
    _ffi.call("zutis_allreduce_hist_p2p", ctx.value, hh.data_ptr(), hh.numel(), oo.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); assert torch.equal(hh, oo)
_ffi.call("zutis_p2p_destroy", ctx.value)
scorer = zutis_b200.StreamingScorer(text, zutis_b200.RunningScore(81), (224, 224))
scorer.submit(tokens, gt); scorer.submit(tokens, gt); scorer.get_scores()
ops.upsample_bilinear(lo2, (50, 60))
torch.cuda.synchronize()
print("sanitize probe done", float(meter.get_scores()[0]["Mean IoU"]))
