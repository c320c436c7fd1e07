"""Generates the golden vectors under tests/golden/ by running the REFERENCE'S OWN CODE: oracle/_ref/libw2xref.so is
/root/reference/src/tensorrt/*.cpp compiled unmodified against CPU mocks of the absent libraries (oracle/Makefile,
oracle/ref_shim/).  Run here (the container that has /root/reference); the outputs are committed so the GPU box and any
checkout without the reference can still check against reference-produced numbers.

  tile_grid.json      calculateTiles rect lists (img2img_render.cpp:7-66)
  ref_goldens.npz     createTileWeights ramps, Img2Img::render outputs for analytic models, stitch replays,
                      padRoi + applyAugmentation + blobFromImages tiles
  ref_hashes.json     getConfigHash / serializeConfig outputs
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import refcases  # noqa: E402
from oracle import ref, tiling  # noqa: E402

assert ref.available(), "build oracle/_ref first: make -C oracle"

# ---- tile grids ----
grids = {}
for name, (W, H, T, OT, S, B) in refcases.GRID_CASES.items():
    n, ins, outs = ref.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)  # only for the derived scalars (nx, ny, overlaps)
    assert n == len(ins) == g.nx * g.ny
    grids[name] = dict(args=[W, H, T, OT, S, B], count=n, nx=g.nx, ny=g.ny, scaled_in=list(g.scaled_in), in_overlap=list(g.in_overlap),
                       out_overlap=list(g.out_overlap), in_rects=[list(r) for r in ins], out_rects=[list(r) for r in outs],
                       source="oracle/_ref: calculateTiles of /root/reference/src/tensorrt/img2img_render.cpp compiled here")
json.dump(grids, open(os.path.join(HERE, "tile_grid.json"), "w"))

arrays = {}
# ---- blend weights: the whole images are rank-1, keep row/column profiles + a checksum of the full images ----
for ox, oy, size in refcases.WEIGHT_CASES:
    w = ref.create_tile_weights(ox, oy, size, size)
    assert np.array_equal(w[..., 0], w[..., 1]) and np.array_equal(w[..., 0], w[..., 2])
    key = f"weights_{ox}_{oy}_{size}"
    arrays[key + "_top_col0"] = w[0, :, 0, 0].copy()
    arrays[key + "_right_row0"] = w[1, 0, :, 0].copy()
    arrays[key + "_bottom_col0"] = w[2, :, 0, 0].copy()
    arrays[key + "_left_row0"] = w[3, 0, :, 0].copy()
    for i, nm in enumerate(("top", "right", "bottom", "left")):
        prof = w[i, :, 0, 0] if i in (0, 2) else w[i, 0, :, 0]
        full = np.broadcast_to(prof[:, None] if i in (0, 2) else prof[None, :], (size, size))
        assert np.array_equal(w[i, ..., 0], full), "weight image is not the outer broadcast of its profile"

# ---- full renders through Img2Img::build/load/render ----
for name, (W, H, T, OT, S, B, batch, tta, seed) in refcases.RENDER_CASES.items():
    src = refcases.frame(W, H, seed)
    out = ref.render(src, refcases.posdep_model(S, T, OT), T, OT, S, B, batch=batch, tta=tta)
    arrays["render_" + name] = out

# ---- stitch replays ----
for name, (W, H, T, OT, S, B, seed) in refcases.STITCH_CASES.items():
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    tiles = refcases.stitch_tiles(g.count, OT, seed)
    out = ref.render(np.zeros((H, W, 3), np.uint8), refcases.replay_model(tiles, OT), T, OT, S, B, batch=1)
    arrays["stitch_" + name] = out

# ---- unpack: padRoi -> applyAugmentation -> blobFromImages ----
for name, (W, H, T, OT, S, B, seed) in refcases.UNPACK_CASES.items():
    src = refcases.frame(W, H, seed)
    rgb = np.ascontiguousarray(src[..., ::-1])
    n, ins, _ = ref.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    tiles = np.stack([ref.apply_augmentation(ref.pad_roi(rgb, r), i % 8) for i, r in enumerate(ins)])
    ref.set_pitch_align(1)  # dense blob rows: the intended behaviour (q4 only bites when 3*T*T is not pitch-aligned)
    blob = ref.blob_from_images(tiles)
    ref.set_pitch_align(512)
    arrays["unpack_" + name] = blob.astype(np.float16)  # exact: the product stores RN-to-fp16 of the same f32 value
    arrays["unpack_" + name + "_f32_checksum"] = np.array([float(blob.astype(np.float64).sum())])
np.savez_compressed(os.path.join(HERE, "ref_goldens.npz"), **arrays)

# ---- config hash + sidecar ----
hashes = []
for dev, ints in refcases.HASH_CASES:
    ref.set_device_name(dev)
    cfg = ref.CBuild(*ints)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c.json")
        ref.serialize_config(p, cfg)
        hashes.append(dict(device=dev, cfg=list(ints), sha256=ref.config_hash(cfg), sidecar=open(p).read()))
ref.set_device_name("NVIDIA B200")
json.dump(hashes, open(os.path.join(HERE, "ref_hashes.json"), "w"), indent=1)
print("grids", {k: v["count"] for k, v in grids.items()})
print("arrays", {k: v.shape for k, v in arrays.items() if not k.startswith("weights")})
print(os.path.getsize(os.path.join(HERE, "ref_goldens.npz")), "bytes")
