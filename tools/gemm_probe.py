"""Micro-benchmark of the contraction kernel alone: time per image for L2-resident vs HBM-streaming inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zutis_b200 import ops

def run(B, Q=81, D=512, h=40, w=40, reps=50, nsets=1, precision="tf32x3"):
    text = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda"), dim=-1)
    sets = [torch.nn.functional.normalize(torch.randn(B, h, w, D, device="cuda"), dim=-1) for _ in range(nsets)]
    for i in range(5):
        ops.contraction(text, sets[i % nsets], precision=precision)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.contraction(text, sets[i % nsets], precision=precision)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"B={B:3d} Q={Q} nsets={nsets} {precision}: {us:8.1f} us/launch  {us / B:6.2f} us/image  tokens {B*h*w*D*4/1e6:.0f} MB")

if __name__ == "__main__":
    run(8); run(16); run(64, nsets=4); run(64, nsets=4, precision="tf32"); run(8, precision="tf32")
    run(37, nsets=8)   # 37*13 = 481 tiles = 3.25 waves
