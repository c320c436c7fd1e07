"""Development: per-role cycle counters of conv3x3_patch_kernel (lib/libw2x_dev.so, W2X_PROF=1) on the big CUNet layer shapes,
optionally with parts of the kernel switched off (W2X_DBG), and the MMA-schedule probe with one / two commits per tile."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
import w2x  # noqa: E402

w2x.use_dev_lib()
what = sys.argv[1] if len(sys.argv) > 1 else "prof"
if what == "probe":
    l = w2x.dev_lib()
    names = {128: "commit/tile, no waits", 136: "no per-tile commit", 640: "two commits/tile, no waits", 129: "wait t-2 + commit", 641: "wait t-2 + two commits",
             385: "two issuing warps, wait t-2", 0: "plain unrolled"}
    for mode, nm in names.items():
        print(f"probe_mma_tiles mode={mode:4d} ({nm:28s}) {l.w2x_probe_mma_tiles(0, 2000, mode):6.1f} cycles/MMA", flush=True)
else:
    rng = np.random.default_rng(0)
    shapes = [(8, 444, 444, 64, 64), (8, 226, 226, 128, 64), (8, 236, 236, 64, 128), (8, 117, 117, 128, 256), (8, 126, 126, 64, 128)]
    for n, h, w_, cin, cout in shapes:
        x = (rng.uniform(-1, 1, size=(n, h, w_, cin))).astype(np.float16)
        wp = (rng.uniform(-1, 1, size=(cout, 9 * cin)) / np.sqrt(9 * cin)).astype(np.float16)
        b = np.zeros(cout, np.float32)
        for rep in range(2):
            w2x.run_conv_layer(0, x, wp, b, cout)
