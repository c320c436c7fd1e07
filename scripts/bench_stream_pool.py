"""BASELINE configs[4] (cfg5): swin_unet/photo scale 4, tile 256, batch 4, ONE ordered stream of synthetic frames sharded round-robin
over the GPUs of a box by ONE process through the engine pool (w2x_pool_*: one worker thread + engine per device, frames retired in
order).  Reports output Mpx/s and fps for 1 / 2 / 4 / 8 GPUs on 1920x1080 -> 7680x4320 and 960x540 -> 3840x2160 frames (SURVEY 8d
runs both readings of the config), pinned host buffers, H2D + D2H inside the timed region, wall clock from the first submit to the
last retired frame.  Usage: python scripts/bench_stream_pool.py [max_gpus] [frames]"""
import json
import os
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

max_gpus = min(int(sys.argv[1]) if len(sys.argv) > 1 else 8, torch.cuda.device_count())
frames_total = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tmp = tempfile.mkdtemp()
_, onnx = __graft_entry__.make_synthetic_model(tmp, scale=4, noise=3, model="swin_unet/photo")
for (W, H) in [(1920, 1080), (960, 540)]:
    srcs = [tiling.synthetic_frame(W, H, s) for s in range(4)]
    n = 1
    base = None
    while n <= max_gpus:
        pool = w2x.Img2ImgPool(list(range(n)))
        msgs = []
        pool.setMessageCallback(lambda s, m: msgs.append((s, m)))
        assert pool.build(onnx, w2x.BuildConfig.fixed(4, 256)), msgs
        assert pool.load(onnx, w2x.RenderConfig(batchSize=4, height=256, width=256, scaling=4)), msgs
        ring = 3 * n + 2
        pin_in = [w2x.PinnedArray((H, W, 3)) for _ in range(ring)]
        pin_out = [w2x.PinnedArray((H * 4, W * 4, 3)) for _ in range(ring)]
        for i in range(ring):
            pin_in[i].array[...] = srcs[i % 4]

        state = {"next": 0}

        def run(count):
            tickets = []
            for f in range(count):
                if f >= ring:
                    assert pool.wait(tickets[f - ring])       # in frame order: the writer's view
                tickets.append(pool.submit(pin_in[f % ring].ptr, W, H, pin_out[f % ring].ptr))
                assert tickets[-1] == state["next"], (tickets[-1], state["next"], msgs)   # tickets count frames since the pool was created
                state["next"] += 1
            for t in tickets[max(0, count - ring):]:
                assert pool.wait(t)

        run(2 * ring)  # warm-up: every engine allocated its buffers, every pinned buffer touched
        sampler = ClockSampler(0)
        sampler.start()
        sampler.mark()
        t0 = time.perf_counter()
        run(frames_total * n)
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        fps = frames_total * n / dt
        mpx = fps * W * 4 * H * 4 / 1e6
        base = base or mpx
        print(json.dumps({"workload": "swin_unet/photo scale4 tile256 batch4, ordered stream through the engine pool (one process)", "frame": [W, H], "n_gpus": n,
                          "frames": frames_total * n, "fps": round(fps, 2), "output_mpx_s": round(mpx, 1), "scaling_efficiency_vs_1gpu": round(mpx / (base * n), 3),
                          "launches_per_engine": pool.launch_counts(), "clocks_gpu0": clocks}), flush=True)
        pool.close()
        for p in pin_in + pin_out:
            p.free()
        n *= 2
