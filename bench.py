#!/usr/bin/env python
"""Benchmark of the ZUTIS mask-decode + scoring hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one batch of synthetic input resident in HBM:
    contraction (text x patch tokens -> low-res logits)  ->  fused upsample/argmax/int16 labels/confusion
    histogram (int32 partial)  ->  histogram merge into the int64 matrix (lazily, as RunningScore does: every
    2^30 scored pixels and when scores are read).
The workload is BASELINE.json configs[1] ("cfg2": ViT-B/16 COCO2017-val shape, Q=81, 40x40 -> 320x320,
batch 64 per GPU).  Images shard across GPUs with no data-path collective (weak scaling); the per-GPU
int64 confusion matrices are summed by ONE NCCL all-reduce when scores are read, after the timed steps
(and inside the e2e region).  Prints ONE JSON line on rank 0.

Inputs are "model-like" synthetic tensors: unit-norm patch tokens obtained by x2 bilinear up-sampling
of coarse random features (what ZUTIS.forward does to ViT tokens, zutis.py:488-497), unit-norm random
text embeddings, int64 ground truth (the reference's dtype) with blocky regions and ignore rows.
`--iid` switches to spatially independent tokens.  Input sets rotate so that each step reads tensors
that are larger than L2 and were last touched several steps ago.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, Q, D, h, w, H, W, ignore label)
    "cfg1": dict(B=2, Q=81, D=512, h=14, w=14, H=224, W=224, ignore=255, desc="ViT-B/32 CPU config, 2x224x224, 81 queries"),
    "cfg2": dict(B=64, Q=81, D=512, h=40, w=40, H=320, W=320, ignore=255, desc="ViT-B/16 COCO2017-val shape, 81 queries, 320x320, batch 64 per GPU"),
    "cfg3": dict(B=32, Q=81, D=512, h=64, w=64, H=512, W=512, ignore=255, desc="ViT-B/16 CoCA shape, 81 queries, 512x512, batch 32 per GPU"),
    "cfg4": dict(B=32, Q=920, D=512, h=56, w=56, H=448, W=448, ignore=1000, desc="ViT-B/16 ImageNet-S919 shape, 920 queries, 448x448, batch 32 per GPU"),
}
METRIC = "mask-decode+mIoU images/sec"
UNIT = "images/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--iid", action="store_true", help="spatially independent tokens (adversarial for pruning)")
    ap.add_argument("--segmented", action="store_true",
                    help="tokens of a trained-model-like output: piecewise-constant label regions + noise (informational; "
                         "the default stays the random-init model-like field SURVEY 8d specifies)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32x3", "tf32"])
    ap.add_argument("--decode", default="auto", choices=["auto", "tiled", "cells", "generic"],
                    help="decode kernel: auto = exact per-cell candidate pruning (cells) when the shape allows")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------ synthetic inputs
def token_mode(args):
    return "segmented" if args.segmented else bool(args.iid)


def make_inputs_torch(cfg, device, seed, iid=False):
    """Model-like synthetic inputs on `device` (torch is input plumbing here, not the measured path)."""
    import torch
    import torch.nn.functional as F
    B, Q, D, h, w, H, W = (cfg[k] for k in ("B", "Q", "D", "h", "w", "H", "W"))
    gen = torch.Generator(device=device).manual_seed(seed)
    text = F.normalize(torch.randn(Q, D, device=device, generator=gen), dim=-1)
    if iid == "segmented":
        # these tokens are built around the text embeddings, and bench.py scores every rotating input set against the
        # first set's text: the text must not depend on the set's seed
        text = F.normalize(torch.randn(Q, D, device=device, generator=torch.Generator(device=device).manual_seed(4242)), dim=-1)
    if iid == "segmented":
        # what a trained model emits: each pixel's token is close to the text embedding of its region's category
        coarse = torch.randint(0, Q, (B, (h + 7) // 8, (w + 7) // 8), device=device, generator=gen)
        regions = coarse[:, torch.arange(h, device=device) // 8][:, :, torch.arange(w, device=device) // 8]      # 8x8-pixel regions
        tokens = F.normalize(text[regions] + 0.04 * torch.randn(B, h, w, D, device=device, generator=gen), dim=-1).contiguous()
    elif iid:
        tokens = F.normalize(torch.randn(B, h, w, D, device=device, generator=gen), dim=-1)
    else:
        coarse = torch.randn(B, D, h // 2, w // 2, device=device, generator=gen)
        up = F.interpolate(coarse, scale_factor=2, mode="bilinear").permute(0, 2, 3, 1)
        tokens = F.normalize(F.layer_norm(up, up.shape[1:]), dim=-1).contiguous()
    # blocky ground truth: <= 5 classes per image in rectangular regions, a few ignore rows
    gt = torch.empty(B, H, W, dtype=torch.int64, device=device)
    classes = torch.randint(0, Q, (B, 5), device=device, generator=gen)
    ys = torch.arange(H, device=device).view(1, H, 1)
    xs = torch.arange(W, device=device).view(1, 1, W)
    cut_y = torch.randint(H // 4, 3 * H // 4, (B, 1, 1), device=device, generator=gen)
    cut_x = torch.randint(W // 4, 3 * W // 4, (B, 1, 1), device=device, generator=gen)
    region = (ys >= cut_y).long() * 2 + (xs >= cut_x).long()
    centre = ((ys - H // 2).abs() < H // 8) & ((xs - W // 2).abs() < W // 8)
    region = torch.where(centre, torch.full_like(region, 4), region)
    gt.copy_(torch.gather(classes, 1, region.view(B, -1)).view(B, H, W))
    gt[:, :4] = cfg["ignore"]
    return text, tokens, gt


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _reasons(self, mask):
        nv = self.nv
        table = {
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        return {k for k, bit in table.items() if mask & bit}

    def _run(self):
        while not self._stop.is_set():
            self.sample_once()
            time.sleep(0.02)

    def sample_once(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.reasons |= self._reasons(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        except Exception:
            pass

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram bytes per launch from the committed ncu summary, if one exists (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


# ------------------------------------------------------------------------------ CPU reference leg
def cpu_reference_step(text, tokens, gt, cfg):
    """The reference's own CPU arithmetic for one batch (oracle.torch_semantic_predict + np.bincount)."""
    from oracle import oracle as O
    H, W = cfg["H"], cfg["W"]
    pred = O.torch_semantic_predict(text, tokens, (H, W))
    meter = O.OracleRunningScore(cfg["Q"])
    meter.update(gt.numpy(), pred)
    return meter.get_scores()


def time_cpu_baseline(cfg, sample_images, reps, iid):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    small = dict(cfg, B=sample_images)
    text, tokens, gt = make_inputs_torch(small, "cpu", 0, iid)
    cpu_reference_step(text, tokens, gt, small)                 # warm-up
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(text, tokens, gt, small)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    return {"value": sample_images / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_images} images of {cfg['desc']} per run, median of {reps} runs after 1 warm-up; "
                      f"oracle/oracle.py torch-CPU restatement of zutis.py:355-372 + running_score.py (torch {torch.__version__}, numpy {np.__version__})"}


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sample = min(16, cfg["B"])
    small = dict(cfg, B=sample)
    text, tokens, gt = make_inputs_torch(small, "cpu", 0, token_mode(args))
    t0 = time.perf_counter(); cpu_reference_step(text, tokens, gt, small); t_first = time.perf_counter() - t0
    budget = 150.0
    while sample > 1 and (args.steps + args.warmup) * t_first > budget:
        sample = max(1, sample // 2); t_first /= 2
    if sample != small["B"]:
        small = dict(cfg, B=sample)
        text, tokens, gt = make_inputs_torch(small, "cpu", 0, token_mode(args))
    for _ in range(args.warmup):
        cpu_reference_step(text, tokens, gt, small)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(text, tokens, gt, small)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    desc = (f"{sample} images of the workload per step; oracle/oracle.py torch-CPU restatement of the reference path "
            f"(the Python reference cannot travel to this box), {torch.get_num_threads()} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['desc']}", "tokens": {True: "iid", False: "model-like (x2-upsampled coarse features)", "segmented": "segmented (piecewise-constant label regions + noise)"}[token_mode(args)],
                   "gt_dtype": "int64", "sample_images_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- our arm
def run_ours(args, cfg, rank, local, world):
    import torch
    import torch.distributed as dist
    import zutis_b200
    from zutis_b200 import _ffi, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the zutis_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _ffi.check(_ffi.lib().zutis_device_check(local))
    B, Q, D, h, w, H, W = (cfg[k] for k in ("B", "Q", "D", "h", "w", "H", "W"))

    # rotating input sets: each > L2 (tokens alone are B*h*w*D*4 bytes), re-read only every n_sets steps
    tok_bytes = B * h * w * D * 4
    n_sets = max(3, int(np.ceil(3 * 126e6 / tok_bytes)))
    sets = [make_inputs_torch(cfg, device, 1000 * rank + s, token_mode(args)) for s in range(n_sets)]
    text = sets[0][0]
    meter = zutis_b200.RunningScore(Q, device=device)
    Qp = (Q + 3) & ~3
    logits_buf = torch.zeros(B, h, w, Qp, device=device)
    logits = logits_buf[..., :Q].permute(0, 3, 1, 2)
    labels = torch.empty(B, H, W, dtype=torch.int16, device=device)
    lib = _ffi.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def gemm_status(fl):
        wsb = lib.zutis_gemm_workspace_bytes(Q, h * w, D, B, fl)
        wsp = torch.empty(max(wsb, 1), dtype=torch.uint8, device=device)
        return lib.zutis_gemm_logits(text.data_ptr(), D, 0, sets[0][1].data_ptr(), D, h * w * D, logits_buf.data_ptr(), 1, Qp,
                                     h * w * Qp, Q, h * w, D, B, fl, wsp.data_ptr(), wsb, stream)

    flags = ops.gemm_flags(args.precision)
    if args.precision == "auto" and gemm_status(flags) == _ffi.ERR_UNSUPPORTED:
        flags = ops.gemm_flags("fp32")          # shape not taken by the tcgen05 kernel: fp32 FFMA kernel
    _ffi.check(gemm_status(flags))
    torch.cuda.synchronize()
    ws_bytes = _ffi.lib().zutis_gemm_workspace_bytes(Q, h * w, D, B, flags)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=device)

    # the text embeddings are constant: their hi/lo split is prepared once (first call below) and reused by every step
    _ffi.check(lib.zutis_gemm_logits(text.data_ptr(), D, 0, sets[0][1].data_ptr(), D, h * w * D, logits_buf.data_ptr(), 1, Qp, h * w * Qp,
                                     Q, h * w, D, B, flags, ws.data_ptr(), ws_bytes, stream))
    step_flags = flags | (_ffi.GEMM_A_PREPARED if (flags & 3) != 0 else 0)

    # the cell decode kernel's run counter: zero-filled once, left zero by every launch
    dws_bytes = _ffi.lib().zutis_decode_workspace_bytes(B, Q, h, w, H, W)
    dws = torch.zeros(max(dws_bytes, 16), dtype=torch.uint8, device=device)
    decode_mode = {"auto": _ffi.DECODE_AUTO, "tiled": _ffi.DECODE_TILED, "cells": _ffi.DECODE_CELLS, "generic": _ffi.DECODE_GENERIC}[args.decode]

    def step(i, ev=None):
        _, tokens, gt = sets[i % n_sets]
        if ev: ev[0].record()
        _ffi.check(lib.zutis_gemm_logits(text.data_ptr(), D, 0, tokens.data_ptr(), D, h * w * D, logits_buf.data_ptr(), 1, Qp, h * w * Qp,
                                         Q, h * w, D, B, step_flags, ws.data_ptr(), ws_bytes, stream))
        if ev: ev[1].record()
        mode = decode_mode | _ffi.DECODE_WORKSPACE_ZEROED
        _ffi.check(lib.zutis_decode_score_ws(logits.data_ptr(), h * w * Qp, 1, w * Qp, Qp, B, Q, h, w, H, W, gt.data_ptr(), _ffi.GT_I64, H * W,
                                             labels.data_ptr(), meter._partial.data_ptr(), Q, mode, dws.data_ptr(), dws_bytes, stream))
        if ev: ev[2].record()
        # RunningScore's own policy: the int32 per-launch partial is folded into the int64 matrix lazily, before it
        # could overflow (every 2^30 scored pixels) and whenever the matrix is read
        merges[0] += meter._pending + B * H * W >= (1 << 30)
        meter._note_pixels(B * H * W)
        if ev: ev[3].record()

    merges = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    meter.reset()
    merges[0] = 0
    barrier()
    # per-kernel events: every step of the timed region, on the launching stream
    n_ev = min(args.steps, 512)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_ev)]
    stride_ev = max(1, args.steps // n_ev)
    sampler = ClockSampler(local)
    sampler.sample_once()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    start.record()
    for i in range(args.steps):
        j = i // stride_ev
        step(i, evs[j] if (i % stride_ev == 0 and j < n_ev) else None)
    stop.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    used = [e for k, e in enumerate(evs) if k * stride_ev < args.steps]
    gemm_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in used]))
    decode_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in used]))
    merge_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in used]))
    meter.all_reduce()
    scores, _ = meter.get_scores()
    total_px = int(meter.counts().sum().item())
    # integrity of the timed path (untimed): the last step's logits against the fp32 FFMA kernel on every image, and
    # its labels against the generic decode kernel
    last_tokens = sets[(args.steps - 1) % n_sets][1]
    exact = ops.contraction(text, last_tokens, precision="fp32")
    logit_err = float(((logits - exact).abs().amax(dim=(1, 2, 3)) / exact.abs().max()).max().item())
    relabel = ops.decode_score(logits, (H, W), mode=_ffi.DECODE_GENERIC)
    label_mismatch = int((relabel != labels).sum().item())

    # ---- e2e: host (pinned) buffers through the C ABI's host entry, copies inside the timed region
    e2e = None
    if args.e2e_steps > 0:
        t_host = sets[0][1].cpu().pin_memory(); g_host = sets[0][2].cpu().pin_memory(); x_host = text.cpu().pin_memory()
        hist_host = np.zeros((Q, Q), np.int64)

        def e2e_step():
            _ffi.check(lib.zutis_semantic_eval_host(x_host.data_ptr(), t_host.data_ptr(), g_host.data_ptr(), _ffi.GT_I64, B, Q, D, h, w, H, W,
                                                    hist_host.ctypes.data, None, flags, local))
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(t_host.numel() * 4 + g_host.numel() * 8 + x_host.numel() * 4),
               "d2h_bytes_per_step": int(Q * Q * 8), "steps": args.e2e_steps, "timer": "wall clock around the synchronous host call, max over ranks",
               "api": "zutis_semantic_eval_host (C ABI, pinned host buffers in, int64 confusion matrix out)"}

    if rank != 0:
        return
    peak, peak_src = measured_peak_hbm()
    bytes_decode = B * (4 * Q * h * w + 8 * H * W + 2 * H * W)
    bytes_gemm = B * (4 * D * h * w + 4 * Q * h * w) + 4 * Q * D
    achieved = bytes_decode / (decode_ms * 1e-3) / 1e9
    cells_path = args.decode in ("auto", "cells") and H >= 4 * h and W >= 4 * w
    kname = "decode_cells_kernel" if cells_path else ("decode_tiled_kernel" if args.decode != "generic" else "decode_generic_kernel")
    launches_per_step = 2                       # contraction + decode
    line = {
        "metric": METRIC, "value": world * B * args.steps / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['desc']}", "tokens": {True: "iid", False: "model-like (x2-upsampled coarse features)", "segmented": "segmented (piecewise-constant label regions + noise)"}[token_mode(args)],
                   "gt_dtype": "int64", "images_per_gpu_per_step": B, "parallelism": f"dp{world} (images sharded, one int64 all-reduce at the end)",
                   "contraction": {0: "fp32 FFMA", 1: "tcgen05 3xTF32", 2: "tcgen05 TF32 single pass"}[flags & 3],
                   "decode": args.decode + (" (exact per-cell candidate pruning, decode_cells_kernel)" if cells_path else ""),
                   "l2_policy": f"{n_sets} rotating input sets of {tok_bytes / 1e6:.0f} MB tokens each (> 126 MB L2 between reuses)"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps + merges[0],
        "kernels_ms": {"contraction": gemm_ms, "decode_score": decode_ms, "hist_merge": merge_ms},
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(kname), "algorithmic_bytes_per_launch": bytes_decode, "peak_source": peak_src,
                     "contraction": {"achieved": bytes_gemm / (gemm_ms * 1e-3) / 1e9, "frac": bytes_gemm / (gemm_ms * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes_per_launch": bytes_gemm, "traffic": ncu_traffic("contraction")}},
        "check": {"mean_iou": float(scores["Mean IoU"]), "pixels_scored": total_px, "max_logit_err_vs_fp32_kernel": logit_err,
                  "labels_differing_from_generic_kernel": label_mismatch},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = time_cpu_baseline(cfg, min(cfg["B"], 64 if cfg["Q"] <= 128 else 2), 3, token_mode(args))
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if world > 1:
        from zutis_b200.distributed import init_distributed
        init_distributed("nccl")
    try:
        run_ours(args, cfg, rank, local, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
