"""Host-side cost of enqueueing one bench step (contraction + decode through the C ABI) against its GPU time."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from zutis_b200 import ops, _ffi
cfg = dict(bench.WORKLOADS["cfg2"])
B, Q, D, h, w, H, W = (cfg[k] for k in ("B", "Q", "D", "h", "w", "H", "W"))
Qp = (Q + 3) & ~3
lib = _ffi.lib()
text, tokens, gt = bench.make_inputs_torch(cfg, "cuda", 0, "segmented")
logits = torch.zeros(B, h, w, Qp, device="cuda")
labels = torch.empty(B, H, W, dtype=torch.int16, device="cuda")
part = torch.zeros(Q * Q, dtype=torch.int32, device="cuda")
flags = _ffi.GEMM_TF32X3
ws_bytes = lib.zutis_gemm_workspace_bytes(Q, h * w, D, B, flags)
ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
dws_bytes = lib.zutis_decode_workspace_bytes(B, Q, h, w, H, W)
dws = torch.zeros(dws_bytes, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
def gemm(fl):
    _ffi.check(lib.zutis_gemm_logits(text.data_ptr(), D, 0, tokens.data_ptr(), D, h * w * D, logits.data_ptr(), 1, Qp, h * w * Qp,
                                     Q, h * w, D, B, fl, ws.data_ptr(), ws_bytes, stream))
def decode():
    _ffi.check(lib.zutis_decode_score_ws(logits.data_ptr(), h * w * Qp, 1, w * Qp, Qp, B, Q, h, w, H, W, gt.data_ptr(), _ffi.GT_I64, H * W,
                                         labels.data_ptr(), part.data_ptr(), Q, _ffi.DECODE_AUTO | _ffi.DECODE_WORKSPACE_ZEROED,
                                         dws.data_ptr(), dws_bytes, stream))
gemm(flags); decode(); torch.cuda.synchronize()
fl = flags | _ffi.GEMM_A_PREPARED
for name, fn in (("contraction", lambda: gemm(fl)), ("decode", decode), ("both", lambda: (gemm(fl), decode()))):
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:12s} host enqueue {1e6 * (t1 - t0) / n:7.1f} us per call, until the GPU is done {1e6 * (t2 - t0) / n:7.1f} us per call")

# the same with three rotating input sets (tokens and ground truth come from DRAM, as in bench.py)
sets = [bench.make_inputs_torch(cfg, "cuda", s, "segmented") for s in range(3)]
def step(i):
    global tokens, gt
    _, tokens, gt = sets[i % 3]
    gemm(fl); decode()
for i in range(6): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(300): step(i)
e1.record(); torch.cuda.synchronize()
print("rotating sets: GPU time per step", round(e0.elapsed_time(e1) / 300 * 1e3, 1), "us")

