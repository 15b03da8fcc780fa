import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zutis_b200
from zutis_b200 import _ffi, ops
B, Q, D, h, w, H, W = 24, 81, 512, 40, 40, 320, 320
gen = torch.Generator().manual_seed(2)
text = torch.nn.functional.normalize(torch.randn(Q, D, generator=gen), dim=-1)
tokens = torch.nn.functional.normalize(torch.randn(B, h, w, D, generator=gen), dim=-1)
tc, kc = text.cuda(), tokens.cuda()
ref = ops.contraction(tc, kc, precision="fp32").contiguous()
base = None
for it in range(30):
    lo = ops.contraction(tc, kc, precision="tf32x3").contiguous()
    torch.cuda.synchronize()
    if base is None: base = lo.clone()
    nd = int((lo != base).sum()); err = float((lo - ref).abs().max())
    if nd or err > 2e-6: print("device gemm run", it, "differs in", nd, "max err vs fp32", err)
print("device gemm determinism check done; err", float((base-ref).abs().max()))
# concurrent: two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for it in range(30):
    with torch.cuda.stream(s1): a = ops.contraction(tc, kc[:12], precision="tf32x3")
    with torch.cuda.stream(s2): b = ops.contraction(tc, kc[12:], precision="tf32x3")
    torch.cuda.synchronize()
    lo = torch.cat([a.contiguous(), b.contiguous()])
    nd = int((lo != base).sum())
    if nd: 
        bad = (lo != base).nonzero()
        print("concurrent run", it, "differs in", nd, "first", bad[0].tolist(), "max diff", float((lo-base).abs().max()))
print("concurrent check done")
lab0 = None
for it in range(10):
    lab = ops.decode_score(base, (H, W)); torch.cuda.synchronize()
    if lab0 is None: lab0 = lab.clone()
    if not torch.equal(lab, lab0): print("decode differs run", it)
print("decode determinism done")
