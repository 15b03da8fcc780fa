import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops
def run(B, Q, D=512, h=40, w=40, precision="tf32x3"):
    text = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda"), dim=-1)
    sets = [torch.nn.functional.normalize(torch.randn(B, h, w, D, device="cuda"), dim=-1) for _ in range(4)]
    for i in range(8):
        ops.contraction(text, sets[i % 4], precision=precision)
    torch.cuda.synchronize()
for Q in (16, 48, 81, 128, 250):
    run(64, Q)
run(64, 16, precision="tf32")
