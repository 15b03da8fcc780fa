"""PeerReducer (all-reduce of the confusion matrix over NVLink peer memory) against NCCL: equality, latency, CUDA-graph replay.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/p2p_reduce_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from zutis_b200.distributed import PeerReducer, init_distributed

rank, local, world = init_distributed("nccl")
dev = torch.device("cuda", local)
red = PeerReducer(max_elements=128 * 128)
ok = True
for Q in (81, 21, 128):
    gen = torch.Generator(device=dev).manual_seed(1000 * Q + rank)
    mine = torch.randint(0, 1 << 40, (Q * Q,), device=dev, dtype=torch.int64, generator=gen)
    want = mine.clone(); dist.all_reduce(want)
    for rep in range(3):
        got = mine.clone()
        red.all_reduce(got)
        ok &= bool(torch.equal(got, want))
    out = torch.empty_like(mine)
    red.all_reduce(mine, out=out)
    ok &= bool(torch.equal(out, want))
# merge + all-reduce in one launch: counts += partial; partial = 0; out = sum over ranks
cnt = torch.randint(0, 1 << 40, (81 * 81,), device=dev, dtype=torch.int64)
part = torch.randint(0, 1 << 20, (81 * 81,), device=dev, dtype=torch.int32)
want = cnt + part.long(); want_local = want.clone(); dist.all_reduce(want)
fused_out = torch.empty_like(cnt)
red.merge_all_reduce(cnt, part, fused_out)
ok &= bool(torch.equal(fused_out, want)) and bool(torch.equal(cnt, want_local)) and int(part.abs().sum()) == 0
# RunningScore.all_reduce(peer=...) with counts still pending in the int32 partial, against the torch.distributed route
import zutis_b200
g2 = torch.Generator().manual_seed(7 + rank)
gt = torch.randint(0, 81, (2, 64, 64), generator=g2).cuda(); pr = torch.randint(0, 81, (2, 64, 64), generator=g2).cuda()
m1, m2 = zutis_b200.RunningScore(81, device=dev), zutis_b200.RunningScore(81, device=dev)
m1.update(gt, pr); m2.update(gt, pr)
m1.all_reduce(peer=red); m2.all_reduce()
ok &= bool(torch.equal(m1.counts(), m2.counts())) and int(m1.counts().sum()) == world * 2 * 64 * 64
m1.all_reduce(peer=red)                                   # nothing pending: the plain in-place call
ok &= int(m1.counts().sum()) == world * world * 2 * 64 * 64
# latency, back to back
Q = 81
buf = torch.ones(Q * Q, device=dev, dtype=torch.int64); out = torch.empty_like(buf)
def timed(fn, n=200):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
us_p2p = timed(lambda: red.all_reduce(buf, out=out))
us_nccl = timed(lambda: dist.all_reduce(out))
# graph replay
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    red.all_reduce(buf, out=out)
    with torch.cuda.graph(g, stream=side):
        red.all_reduce(buf, out=out)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
out.zero_()
us_graph = timed(lambda: g.replay())
ok &= bool((out == world).all())
print(f"rank {rank}/{world}: equal to NCCL {ok}; Q=81 all-reduce us/call: peer memory {us_p2p:.1f}, peer memory in a CUDA graph {us_graph:.1f}, NCCL {us_nccl:.1f}", flush=True)
dist.barrier()
red.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
