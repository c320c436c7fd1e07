"""Generates tests/golden/tile_grid.json from the NumPy oracle (oracle/tiling.py).  The headline numbers of every
case (tile counts, scaled tile, overlaps, last output rect) were derived by hand from the reference formulas
(/root/reference/src/tensorrt/img2img_render.cpp:7-66) in SURVEY.md 8a and are asserted independently in
tests/test_tile_grid.py; this file pins the complete rect lists so the C restatement and the product can be diffed."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tiling  # noqa: E402

CASES = {
    # name: (W, H, tile, out_tile, scale, blend)
    "cfg1_cunet2x_t64_256x256": (256, 256, 64, 56, 2, 1 / 16),
    "cfg2_cunet2x_t256_1080p": (1920, 1080, 256, 440, 2, 1 / 16),
    "cfg3_cunet1x_t400_1080p": (1920, 1080, 400, 344, 1, 1 / 16),
    "cfg4_swin4x_t256_1080p": (1920, 1080, 256, 960, 4, 1 / 16),
    "cfg5_swin4x_t256_960x540": (960, 540, 256, 960, 4, 1 / 16),
    "q5_cunet2x_t400_blend32_drift": (1920, 1080, 400, 728, 2, 1 / 32),
    "blend0_cunet2x_t64": (200, 120, 64, 56, 2, 0.0),
    "blend8_cunet1x_t128": (197, 231, 128, 72, 1, 1 / 8),
    "single_tile": (20, 20, 64, 56, 2, 1 / 16),
}

out = {}
for name, (W, H, T, OT, S, B) in CASES.items():
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    out[name] = dict(args=[W, H, T, OT, S, B], count=g.count, nx=g.nx, ny=g.ny, scaled_in=list(g.scaled_in),
                     in_overlap=list(g.in_overlap), out_overlap=list(g.out_overlap),
                     in_rects=[list(r) for r in g.in_rects], out_rects=[list(r) for r in g.out_rects])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tile_grid.json"), "w") as f:
    json.dump(out, f)
print({k: v["count"] for k, v in out.items()})
