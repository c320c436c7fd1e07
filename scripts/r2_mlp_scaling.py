"""Fused SwinUNet token kernels: launch time against the number of 128-token tiles per CTA (fixed cost vs per-tile cost)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import w2x  # noqa: E402
from test_gpu_swin_mlp import make_case  # noqa: E402

for c in (96, 192):
    for tiles_per_cta in (1, 2, 4, 8, 13):
        tokens = 148 * 128 * tiles_per_cta
        case = make_case(tokens, 3, c)
        _, ms = w2x.run_swin_mlp(*case, reps=30)
        print("c=%d tiles/CTA=%2d tokens=%7d  mlp %.4f ms" % (c, tiles_per_cta, tokens, ms), flush=True)
