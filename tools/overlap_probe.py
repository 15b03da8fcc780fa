"""contraction(i+1) on one stream while decode(i) runs on another, against the same launches on one stream.

r02 measurements on cfg2 (us/step, serial -> two streams): product kernels 136.6 -> 125.8.  With temporary switches (not
committed) that shrink both kernels until one CTA of each fits an SM together -- contraction rings of 3 pixel + 2 category
slots (98 KB instead of 227 KB: 139.5 serial, i.e. the contraction barely needs its deep rings), decode CTAs of 8 / 10 / 12
warps -- the co-resident pair ran at 172.6 / 156.5 / 143.9: the decode kernel needs all 20 warps x 96 registers of the SM
to itself, and what it loses with fewer warps is more than the overlap returns."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from zutis_b200 import _ffi, ops
cfg = dict(bench.WORKLOADS["cfg2"])
dev = torch.device("cuda:0")
R = bench.SemanticRunner(cfg, dev, 0, False)
B, Q, D, h, w, H, W = (cfg[k] for k in ("B", "Q", "D", "h", "w", "H", "W"))
lib, F = R.lib, _ffi
Qp = R.Qp
bufs = [torch.zeros(B, h, w, Qp, device=dev) for _ in range(2)]
sg, sd = torch.cuda.Stream(), torch.cuda.Stream()
def gemm(i, stream):
    _, tokens, gt = R.sets[i % R.n_sets]
    F.check(lib.zutis_gemm_logits(R.text.data_ptr(), D, 0, tokens.data_ptr(), D, h * w * D, bufs[i % 2].data_ptr(), 1, Qp, h * w * Qp, Q, h * w, D, B,
                                  R.step_flags, R.ws.data_ptr(), R.ws_bytes, stream.cuda_stream))
def decode(i, stream):
    _, tokens, gt = R.sets[i % R.n_sets]
    lg = bufs[i % 2][..., :Q].permute(0, 3, 1, 2)
    F.check(lib.zutis_decode_score_ws(lg.data_ptr(), h * w * Qp, 1, w * Qp, Qp, B, Q, h, w, H, W, gt.data_ptr(), F.GT_I64, H * W, R.labels.data_ptr(),
                                      R.meter._partial.data_ptr(), Q, R.decode_mode | F.DECODE_WORKSPACE_ZEROED, R.dws.data_ptr(), R.dws_bytes, stream.cuda_stream))
def serial(K):
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(K):
        gemm(i, s); decode(i, s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
def piped(K):
    g_done = [torch.cuda.Event() for _ in range(K)]
    d_done = [torch.cuda.Event() for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(sg); sd.wait_event(e0)
    for i in range(K):
        if i >= 2: sg.wait_event(d_done[i - 2])       # logits buffer i%2 is free again
        gemm(i, sg); g_done[i].record(sg)
        sd.wait_event(g_done[i]); decode(i, sd); d_done[i].record(sd)
    e1.record(sd); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
for _ in range(2): serial(20); piped(20)
R.meter.reset() if hasattr(R.meter, "reset") else None
print("env", {k: v for k, v in os.environ.items() if k.startswith("ZUTIS_EXP")}, "serial %.1f us/step  piped %.1f us/step" % (serial(200), piped(200)), flush=True)
