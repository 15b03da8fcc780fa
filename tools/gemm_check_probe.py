"""tf32x3 contraction against the fp32 FFMA kernel on the whole cfg2 batch (every image, every tile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops
gen = torch.Generator(device="cuda").manual_seed(1)
text = torch.nn.functional.normalize(torch.randn(81, 512, device="cuda", generator=gen), dim=-1)
tokens = torch.nn.functional.normalize(torch.randn(64, 40, 40, 512, device="cuda", generator=gen), dim=-1)
ref = ops.contraction(text, tokens, precision="fp32")
for rep in range(3):
    lo = ops.contraction(text, tokens, precision="tf32x3")
    err = (lo - ref).abs().amax(dim=(1, 2, 3))
    print("rep", rep, "max err over images", float(err.max()), "worst image", int(err.argmax()), "scale", float(ref.abs().max()))
