"""Contraction through a CUDA graph with and without the champion by-product (cfg2 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops
text = torch.nn.functional.normalize(torch.randn(81, 512, device="cuda"), dim=-1)
sets = [torch.nn.functional.normalize(torch.randn(64, 40, 40, 512, device="cuda"), dim=-1) for _ in range(4)]
for with_ws in (False, True):
    cache = {}
    ws = ops.DecodeWorkspace() if with_ws else None
    for i in range(5): ops.contraction(text, sets[i % 4], precision="tf32x3", a_cache=cache, decode_ws=ws)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20): ops.contraction(text, sets[i % 4], precision="tf32x3", a_cache=cache, decode_ws=ws)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("champion by-product" if with_ws else "plain contraction  ", round(e0.elapsed_time(e1) * 1e3 / 100, 2), "us per launch (graph)")
