"""Time the decode launches alone (champions already in the workspace) on bench.py's token modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from zutis_b200 import ops, _ffi
cfg = dict(bench.WORKLOADS["cfg2"])
B, Q, h, w, H, W = (cfg[k] for k in ("B", "Q", "h", "w", "H", "W"))
lib = _ffi.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in (False, "segmented", True):
    text, tokens, gt = bench.make_inputs_torch(cfg, "cuda", 0, mode)
    ws = ops.DecodeWorkspace()
    lo = ops.contraction(text, tokens, precision="tf32x3", decode_ws=ws)
    assert ws.ready_for is not None
    labels = torch.empty(B, H, W, dtype=torch.int16, device="cuda")
    part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
    dws_bytes = lib.zutis_decode_workspace_bytes(B, Q, h, w, H, W)
    stream = torch.cuda.current_stream().cuda_stream
    def run():
        _ffi.check(lib.zutis_decode_score_ws(lo.data_ptr(), lo.stride(0), lo.stride(1), lo.stride(2), lo.stride(3), B, Q, h, w, H, W,
                                             gt.data_ptr(), _ffi.GT_I64, H * W, labels.data_ptr(), part.data_ptr(), Q,
                                             _ffi.DECODE_AUTO | _ffi.DECODE_CHAMPIONS_READY, ws.buf.data_ptr(), dws_bytes, stream))
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): run()
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 10
    # the same with L2 flushed before every call (ground truth and labels from DRAM, as inside the pipeline)
    tot = 0.0
    for _ in range(30):
        flush.zero_()
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print({False: "model-like", True: "iid", "segmented": "segmented"}[mode], "decode alone: L2-warm", round(warm, 2), "us, L2-flushed", round(tot / 30 * 1e3, 2), "us per call")
