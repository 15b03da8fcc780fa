"""Kernel-only timing of the tensor-core contraction (cfg2 shape) through a CUDA graph, so that the Python/ctypes launch
cost (~50 us per call, more than the kernel) does not hide differences.  usage: gemm_probe3.py [tf32x3|tf32]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
text = torch.nn.functional.normalize(torch.randn(81, 512, device="cuda"), dim=-1)
sets = [torch.nn.functional.normalize(torch.randn(64, 40, 40, 512, device="cuda"), dim=-1) for _ in range(4)]
cache = {}
for i in range(5): ops.contraction(text, sets[i % 4], precision=prec, a_cache=cache)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(20): ops.contraction(text, sets[i % 4], precision=prec, a_cache=cache)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
print("debug", os.environ.get("ZUTIS_GEMM_DEBUG"), "sb", os.environ.get("ZUTIS_GEMM_SB"), prec, "us/launch (graph)", round(e0.elapsed_time(e1) * 1e3 / 100, 2))
