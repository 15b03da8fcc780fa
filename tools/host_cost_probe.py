"""Host-side cost of enqueueing one bench.py step (no GPU wait): is the timed loop launch-bound?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
cfg = dict(bench.WORKLOADS["cfg2"])
for pipe in (False, True):
    R = bench.SemanticRunner(cfg, torch.device("cuda:0"), 0, False, pipelined=pipe)
    for i in range(20): R.step(i)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(64)]
    for with_ev in (False, True):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(64): R.step(i, evs[i] if with_ev else None)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"pipelined={pipe} events={with_ev}: enqueue {1e6 * (t1 - t0) / 64:.1f} us/step, until idle {1e6 * (t2 - t0) / 64:.1f} us/step", flush=True)
