import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
text = torch.nn.functional.normalize(torch.randn(81, 512, device="cuda"), dim=-1)
sets = [torch.nn.functional.normalize(torch.randn(64, 40, 40, 512, device="cuda"), dim=-1) for _ in range(2)]
cache = {}
for i in range(3): ops.contraction(text, sets[i % 2], precision=prec, a_cache=cache)
torch.cuda.synchronize()
