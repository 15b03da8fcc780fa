import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zutis_b200 import ops
Q, D, h, w = 81, 512, 40, 40
gen = torch.Generator().manual_seed(2)
text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1).cuda()
tokens = torch.nn.functional.normalize(torch.randn(64, h, w, D, generator=gen), dim=-1).cuda()
for B in (2, 9, 11, 12, 13, 24, 64):
    ref = ops.contraction(text, tokens[:B], precision="fp32").contiguous()
    for it in range(3):
        lo = ops.contraction(text, tokens[:B], precision="tf32x3").contiguous()
        torch.cuda.synchronize()
        err = (lo - ref).abs().amax(dim=(1, 2, 3))
        bad = (err > 1e-5).nonzero().flatten().tolist()
        print(f"B={B:3d} run {it}: max err {float(err.max()):.2e} bad images {bad[:12]}{'...' if len(bad)>12 else ''}")
