"""Time the decode kernels alone, per mode, on bench.py's workloads and token modes (L2 flushed before every call),
and check each mode's labels + histogram against the generic kernel.   python tools/decode_modes_probe.py [cfg2 cfg4 ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from zutis_b200 import ops, _ffi

names = [a for a in sys.argv[1:] if a in bench.WORKLOADS] or ["cfg2"]
token_modes = [False, "segmented", True] if "--all-tokens" in sys.argv else [False]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in names:
    cfg = dict(bench.WORKLOADS[name])
    B, Q, h, w, H, W = (cfg[k] for k in ("B", "Q", "h", "w", "H", "W"))
    for tm in token_modes:
        text, tokens, gt = bench.make_inputs_torch(cfg, "cuda", 0, tm)
        lo = ops.contraction(text, tokens, precision="tf32x3")
        ref_part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
        ref = ops.decode_score(lo, (H, W), gt=gt, hist_partial=ref_part, mode=_ffi.DECODE_GENERIC)
        for mode_name, mode in (("cells", _ffi.DECODE_CELLS), ("tiled", _ffi.DECODE_TILED)):
            part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
            def run():
                return ops.decode_score(lo, (H, W), gt=gt, hist_partial=part, mode=mode)
            try:
                got = run()
            except Exception as e:
                print(name, tm, mode_name, "unsupported:", str(e)[:80]); continue
            ok = bool(torch.equal(got, ref)) and bool(torch.equal(part, ref_part))
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for _ in range(20):
                flush.zero_()
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            bytes_alg = B * (4 * Q * h * w + 8 * H * W + 2 * H * W)
            us = tot / 20 * 1e3
            print(f"{name} tokens={tm} {mode_name}: {us:.1f} us/launch (incl. torch.empty for labels), exact={ok}, "
                  f"{bytes_alg / us / 1e3:.0f} GB/s algorithmic", flush=True)
