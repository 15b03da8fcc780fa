"""The cell decode kernel alone, small, for compute-sanitizer --tool racecheck."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops, _ffi
coarse = torch.randn(1, 21, 4, 4, device="cuda")
smooth = torch.nn.functional.interpolate(coarse, size=(8, 8), mode="bilinear")
pm = torch.zeros(1, 8, 8, 24, device="cuda"); pm[..., :21] = smooth.permute(0, 2, 3, 1)
part = torch.zeros(21 * 21, dtype=torch.int32, device="cuda")
ops.decode_score(pm[..., :21].permute(0, 3, 1, 2), (64, 64), gt=torch.randint(0, 21, (1, 64, 64), device="cuda"), hist_partial=part, mode=_ffi.DECODE_CELLS)
torch.cuda.synchronize()
print("done")
