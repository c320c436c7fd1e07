#!/usr/bin/env python
"""GPU comparator for SURVEY 8(d): the SAME graphs (oracle/models.py, seeded synthetic weights) run by PyTorch-CUDA in fp16
(cuDNN / cuBLAS kernels, channels_last, cudnn.benchmark) on the model stage only -- tiles already unpacked, no stitching --
timed with CUDA events.  This is a *library baseline, not TensorRT* (TensorRT is not available in this image) and it is not
part of the product or of bench.py's contract; it only says how the hand-written kernels compare with the stock library path
on the same box.  Usage: python scripts/bench_library_baseline.py [cunet|swin]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler  # noqa: E402
from oracle.models import make_model  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cunet"
    if which == "cunet":
        family, scale, batch, tiles, out_mpx = "cunet", 2, 8, 60, 3840 * 2160 / 1e6
    else:
        family, scale, batch, tiles, out_mpx = "swin_unet", 4, 4, 45, 7680 * 4320 / 1e6
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda:0")
    m = make_model(family, scale, 1234).to(dev).half().eval()
    if which == "cunet":
        m = m.to(memory_format=torch.channels_last)
    x = torch.rand(batch, 3, 256, 256, device=dev, dtype=torch.float16)
    if which == "cunet":
        x = x.contiguous(memory_format=torch.channels_last)
    last = tiles - (tiles // batch) * batch
    batches = [batch] * (tiles // batch) + ([last] if last else [])

    def frame():
        with torch.no_grad():
            for b in batches:
                m(x[:b])

    for _ in range(3):
        frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(5):  # nvidia-smi needs a few hundred ms to start
        frame()
    torch.cuda.synchronize()
    sampler.mark()
    e0.record()
    for _ in range(iters):
        frame()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    clocks = sampler.stop()
    print(json.dumps({"impl": "library baseline (PyTorch-CUDA fp16, cuDNN/cuBLAS; not TensorRT)", "workload": which, "model_stage_ms_per_frame": ms,
                      "equivalent_output_mpx_s": out_mpx / (ms / 1e3), "tiles": tiles, "batch": batch, "torch": torch.__version__,
                      "cudnn": torch.backends.cudnn.version(), "clocks": clocks, "note": "model stage only (no unpack / stitch / copies); padding slots skipped like the product"}))


if __name__ == "__main__":
    main()
