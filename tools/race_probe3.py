import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zutis_b200 import ops
Q, D, h, w = 81, 512, 40, 40
gen = torch.Generator().manual_seed(2)
text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).cuda()
tokens = torch.nn.functional.normalize(torch.randn(12, h, w, D, generator=gen), dim=-1).cuda()
ref = ops.contraction(text, tokens, precision="fp32").contiguous().permute(0, 2, 3, 1).reshape(12 * 1600, Q)   # [pixels, Q]
lo = ops.contraction(text, tokens, precision="tf32x3").contiguous().permute(0, 2, 3, 1).reshape(12 * 1600, Q)
torch.cuda.synchronize()
def tile_rows(t):
    b, pt = divmod(t, 13); r0 = b * 1600 + pt * 128; return slice(r0, min(r0 + 128, (b + 1) * 1600))
for t in (148, 149, 155):
    got = lo[tile_rows(t)]; want = ref[tile_rows(t)]
    first = ref[tile_rows(t - 148)]
    d = (got - want)
    print(f"tile {t}: err max {float(d.abs().max()):.3e} mean {float(d.abs().mean()):.3e}; |got| mean {float(got.abs().mean()):.3e} |want| mean {float(want.abs().mean()):.3e}")
    n = min(got.shape[0], first.shape[0])
    print("   corr(got, want) %.4f  corr(got, first-tile) %.4f  corr(got-want, first) %.4f" % (
        float(torch.corrcoef(torch.stack([got[:n].flatten(), want[:n].flatten()]))[0, 1]),
        float(torch.corrcoef(torch.stack([got[:n].flatten(), first[:n].flatten()]))[0, 1]),
        float(torch.corrcoef(torch.stack([d[:n].flatten(), first[:n].flatten()]))[0, 1])))
    # per-row error profile
    re = d.abs().amax(dim=1)
    print("   rows with err>1e-5:", int((re > 1e-5).sum()), "of", got.shape[0], " first bad row", int((re > 1e-5).nonzero()[0]) if (re>1e-5).any() else None)
    ce = d.abs().amax(dim=0)
    print("   cols with err>1e-5:", int((ce > 1e-5).sum()), "of", Q)
    # is got = want computed from a K-subset? ratio
    print("   got/want median ratio", float((got / want).median()))
print("---- bad row sets")
for t in range(148, 156):
    got = lo[tile_rows(t)]; want = ref[tile_rows(t)]
    re = (got - want).abs().amax(dim=1)
    bad = (re > 1e-5).nonzero().flatten().tolist()
    print(t, bad)
