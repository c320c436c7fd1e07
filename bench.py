#!/usr/bin/env python
"""bench.py — output Mpx/s of the tile -> model -> stitch hot path (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] = cunet/art scale 2 noise 3, tileSize 256, batchSize 8, fp16, synthetic
1920x1080 BGR frames -> 3840x2160 (60 tiles + 4 padding slots per frame).  One "step" = one frame.

  value : whole-job output Mpx/s with the input frames already resident in HBM (w2x_render_device), CUDA events on the
          engine's stream, max over ranks.
  e2e   : the same metric through the reference-facing call path with HOST buffers: pinned host frame -> H2D ->
          render -> D2H into a pinned host frame, pipelined (w2x_submit / w2x_wait), copies inside the timed region.
  roofline : the model stage (tcgen05 implicit-GEMM convolutions dominate it) against the measured dense bf16 peak.
  cpu_baseline / --impl reference : the reference has no CPU path and cannot be built here; this arm times
          the oracle port (PyTorch fp32 on the host cores + NumPy tiling restatement) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "waifu2x-tensorrt_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

FRAME_W, FRAME_H, TILE, BATCH, SCALE, BLEND = 1920, 1080, 256, 8, 2, 1.0 / 16.0
WORKLOAD = "cunet/art scale2 noise3 tile256 batch8 fp16, synthetic 1920x1080 -> 3840x2160 frames (BASELINE configs[1])"
METRIC, UNIT = "output Mpx/s", "Mpx/s"


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_sample(threads: int, repeats: int = 1):
    """Oracle port on the host cores: a 640x360 crop of the workload's frame (8 tiles of 256 -> 1280x720 output)."""
    import numpy as np
    import torch
    from oracle import tiling
    from oracle.models import make_model
    torch.set_num_threads(threads)
    model = make_model("cunet", SCALE, 1234)
    frame = tiling.synthetic_frame(FRAME_W, FRAME_H, 0)[:360, :640].copy()

    def fn(x):
        with torch.no_grad():
            return model(torch.from_numpy(np.ascontiguousarray(x))).numpy()

    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = tiling.render(frame, fn, TILE, 2 * TILE - 72, SCALE, BLEND, batch=1)
        times.append(time.perf_counter() - t0)
    mpx = out.shape[0] * out.shape[1] / 1e6
    return mpx, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    if args.warmup > 0:
        cpu_oracle_sample(threads, repeats=1)  # one warm-up pass at most
    mpx, times = cpu_oracle_sample(threads, repeats=args.steps)
    total = sum(times)
    value = mpx * args.steps / total
    sample = "640x360 crop of the 1920x1080 synthetic frame (8 tiles of 256) per step, oracle port: PyTorch fp32 CPU + NumPy tiling"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference (TensorRT + OpenCV-CUDA) has no CPU path and cannot be built in this image; this is the oracle port",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also print a per-layer profile to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__
    import w2x
    from oracle import tiling  # synthetic frame generator only; never on the measured path

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w2x.lib()

    # ---- build + load through the reference-shaped API ----
    tmp = tempfile.mkdtemp(prefix=f"w2x_bench_r{rank}_")
    _, onnx_path = __graft_entry__.make_synthetic_model(tmp, scale=SCALE, noise=3)
    eng = w2x.Img2Img()
    msgs = []
    eng.setMessageCallback(lambda sev, m: msgs.append((sev, m)))
    if not eng.build(onnx_path, w2x.BuildConfig.fixed(BATCH, TILE, device=local)):
        raise SystemExit(f"build failed: {msgs}")
    if not eng.load(onnx_path, w2x.RenderConfig(deviceId=local, batchSize=BATCH, height=TILE, width=TILE, scaling=SCALE, overlap=(BLEND, BLEND))):
        raise SystemExit(f"load failed: {msgs}")

    n_in = 4  # distinct input frames, rotated
    frames = [tiling.synthetic_frame(FRAME_W, FRAME_H, rank * 1000 + s) for s in range(n_in)]
    in_bytes, out_bytes = FRAME_W * FRAME_H * 3, FRAME_W * SCALE * FRAME_H * SCALE * 3
    out_mpx = FRAME_W * SCALE * FRAME_H * SCALE / 1e6
    d_in = [eng.device_alloc(in_bytes) for _ in range(n_in)]
    for d, f in zip(d_in, frames):
        eng.h2d(d, f)
    d_out = eng.device_alloc(out_bytes)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        eng.sync()

    # ---- device-resident throughput (`value`) ----
    for i in range(args.warmup):
        assert eng.render_device(d_in[i % n_in], FRAME_W, FRAME_H, d_out), eng.last_error
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    stage = {"unpack": 0.0, "model": 0.0, "stitch": 0.0}
    eng.timer_mark(0)
    for i in range(args.steps):
        assert eng.render_device(d_in[i % n_in], FRAME_W, FRAME_H, d_out), eng.last_error
    eng.timer_mark(1)
    eng.sync()
    torch.cuda.synchronize()
    ms_total = eng.timer_elapsed_ms(0, 1)
    launches = eng.launch_count - launches0
    # stage split of the last timed frame (events recorded inside the timed region)
    last = eng.last_stage_ms()
    for k in stage:
        stage[k] = last.get(k, 0.0)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    else:
        ms_max = ms_total
    value = world * args.steps * out_mpx / (ms_max / 1e3)

    # ---- end-to-end through host buffers (`e2e`) ----
    ring = 4
    pin_in = [w2x.PinnedArray((FRAME_H, FRAME_W, 3)) for _ in range(ring)]
    pin_out = [w2x.PinnedArray((FRAME_H * SCALE, FRAME_W * SCALE, 3)) for _ in range(ring)]
    for i in range(ring):
        pin_in[i].array[...] = frames[i % n_in]
    tickets = []

    def e2e_pass(n):
        tickets.clear()
        for i in range(n):
            if i >= ring:
                assert eng.wait(tickets[i - ring])  # dst buffer about to be reused
            t = eng.submit(pin_in[i % ring].ptr, FRAME_W, FRAME_H, pin_out[i % ring].ptr)
            assert t >= 0, eng.last_error
            tickets.append(t)
        for t in tickets[-ring:]:
            assert eng.wait(t)

    e2e_pass(max(args.warmup, 1))
    barrier()
    eng.timer_mark(2, 1)
    t0 = time.perf_counter()
    e2e_pass(args.steps)
    eng.timer_mark(3, 2)
    eng.sync()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = eng.timer_elapsed_ms(2, 3)
    if e2e_ms <= 0:
        e2e_ms = e2e_wall_ms
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * args.steps * out_mpx / (e2e_ms / 1e3)
    checksum = int(pin_out[(args.steps - 1) % ring].array[::97, ::89].astype(np.uint64).sum())  # result read on the host

    if rank == 0:
        peaks, peak_kind = _peaks()
        flops_frame = eng.flops_per_tile * 60  # 60 real tiles per 1080p frame (padding slots excluded)
        model_ms = stage["model"]
        achieved = flops_frame / (model_ms / 1e3) / 1e12 if model_ms > 0 else None
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("model_stage_dram_bytes_per_frame")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_rank": args.steps, "sharding": "frames round-robin, one engine per GPU, no collective",
                       "l2": "each step streams >1 GB of activations + 4 rotating input frames (> 126 MB L2)", "weights": "seeded synthetic (seed 1234)"},
            "fps": world * args.steps / (ms_max / 1e3),
            "stage_ms_last_frame": stage,
            "roofline": {"bound": "tensor", "kernel": "model stage (igemm_kernel tcgen05 convs + first-layer + SE kernels)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_kind})", "flops_per_frame": flops_frame, "traffic": traffic},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps, "host_checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            mpx, times = cpu_oracle_sample(threads, repeats=1)
            line["cpu_baseline"] = {"value": mpx / times[0], "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "one 640x360 crop (8 tiles of 256 -> 1280x720), PyTorch fp32 CPU oracle + NumPy tiling, 1 pass"}
        if args.layers:
            for name, ms, fl in eng.profile_layers(3):
                print(f"  {name:28s} {ms:8.3f} ms  {fl / ms / 1e9 if ms > 0 else 0:9.1f} TFLOP/s", file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
