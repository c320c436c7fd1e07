"""Strong scaling of ONE large image over the GPUs of a box in row-band mode (w2x_render_banded): latency of a
7680x4320 -> 2x render (cunet/art, tile 256, batch 8) and a 3840x2160 -> 4x render (swin_unet/art, tile 256, batch 4) with
1 / 2 / 4 / 8 bands, host buffers, wall clock around the call (median of 5), plus the clocks seen.  One process drives all GPUs.
Usage: python scripts/bench_banded.py [max_gpus]"""
import json
import os
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__  # noqa: E402
import w2x  # noqa: E402
from bench import ClockSampler  # noqa: E402
from oracle import tiling  # noqa: E402

ngpu = min(int(sys.argv[1]) if len(sys.argv) > 1 else 8, torch.cuda.device_count())
CASES = [("cunet/art", 2, 256, 8, 7680, 4320), ("swin_unet/art", 4, 256, 4, 3840, 2160)]
for model, scale, tile, batch, W, H in CASES:
    tmp = tempfile.mkdtemp()
    _, onnx = __graft_entry__.make_synthetic_model(tmp, scale=scale, noise=3, model=model)
    src_pin = w2x.PinnedArray((H, W, 3))        # pinned host buffers: the per-band copies of all GPUs then run concurrently
    dst_pin = w2x.PinnedArray((H * scale, W * scale, 3))
    src_pin.array[...] = tiling.synthetic_frame(W, H, 1)
    src = src_pin.array
    engines = []
    for d in range(ngpu):
        e = w2x.Img2Img()
        assert e.build(onnx, w2x.BuildConfig.fixed(batch, tile, device=d)) and e.load(onnx, w2x.RenderConfig(deviceId=d, batchSize=batch, height=tile, width=tile, scaling=scale))
        engines.append(e)
    ref = None
    n = 1
    while n <= ngpu:
        out = w2x.render_banded(engines[:n], src, dst_pin.array)  # warm-up, allocations
        if ref is None:
            ref = out.copy()
        same = bool(np.array_equal(out, ref))
        sampler = ClockSampler(0)
        sampler.start()
        times = []
        for _ in range(5):
            t0 = time.perf_counter()
            out = w2x.render_banded(engines[:n], src, dst_pin.array)
            times.append(time.perf_counter() - t0)
        clocks = sampler.stop()
        ms = statistics.median(times) * 1e3
        print(json.dumps({"mode": "row-band single image", "model": model, "scale": scale, "input": [W, H], "n_gpus": n, "latency_ms": round(ms, 2),
                          "output_mpx_s": round(W * scale * H * scale / 1e6 / (ms / 1e3), 1), "byte_identical_to_1gpu": same, "clocks_gpu0": clocks}), flush=True)
        n *= 2
    for e in engines:
        e.close()
