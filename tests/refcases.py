"""Cases shared by tests/golden/make_ref_goldens.py (which runs them through oracle/_ref, the reference's own code compiled
here) and the parity tests (which replay them through the NumPy oracle and the product)."""
import numpy as np

# name: (W, H, tile, out_tile, scale, blend) -- BASELINE.json configs first
GRID_CASES = {
    "cfg1_cunet2x_t64_256x256": (256, 256, 64, 56, 2, 1 / 16),
    "cfg2_cunet2x_t256_1080p": (1920, 1080, 256, 440, 2, 1 / 16),
    "cfg3_cunet1x_t400_1080p": (1920, 1080, 400, 344, 1, 1 / 16),
    "cfg4_swin4x_t256_1080p": (1920, 1080, 256, 960, 4, 1 / 16),
    "cfg5_swin4x_t256_960x540": (960, 540, 256, 960, 4, 1 / 16),
    "q5_cunet2x_t400_blend32_drift": (1920, 1080, 400, 728, 2, 1 / 32),
    "blend0_cunet2x_t64": (200, 120, 64, 56, 2, 0.0),
    "blend8_cunet1x_t128": (197, 231, 128, 72, 1, 1 / 8),
    "single_tile": (20, 20, 64, 56, 2, 1 / 16),
    "8k_cunet2x_t640": (7680, 4320, 640, 1208, 2, 1 / 16),
    "swin2x_t112_ragged": (333, 201, 112, 192, 2, 1 / 16),
    "blend4_swin4x_t64": (131, 77, 64, 192, 4, 1 / 4),
}

# (overlap_x, overlap_y, tile size) for createTileWeights
WEIGHT_CASES = [(8, 8, 56), (32, 32, 440), (25, 25, 344), (64, 64, 960), (0, 0, 56), (3, 7, 40), (16, 4, 72)]

# name: (W, H, tile, out_tile, scale, blend, batch, tta, seed)
RENDER_CASES = {
    "cunet2x_t64_b2": (100, 76, 64, 56, 2, 1 / 16, 2, False, 0),
    "cunet2x_t64_blend0_b3": (90, 70, 64, 56, 2, 0.0, 3, False, 1),
    "cunet1x_t64_blend32_b1": (41, 37, 64, 8, 1, 1 / 32, 1, False, 2),
    "swin4x_t64_b4": (110, 60, 64, 192, 4, 1 / 16, 4, False, 3),
    "single_tile_b1": (20, 20, 64, 56, 2, 1 / 16, 1, False, 4),
    "cunet2x_t64_tta_b2": (70, 50, 64, 56, 2, 1 / 16, 2, True, 5),
    "cunet2x_t128_b8_padslots": (300, 160, 128, 184, 2, 1 / 16, 8, False, 6),
}


def frame(w, h, seed):
    return np.random.default_rng(1000 + seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def posdep_model(scale, tile, out_tile):
    """Analytic stand-in for the network: centre crop, nearest up-sampling, per-channel gain and a position-dependent offset
    (so it is NOT equivariant under flips / rotations: TTA order and de-augmentation errors change the result).  Pure
    elementwise IEEE f32, so every platform computes the same bits."""
    off = (tile * scale - out_tile) // (2 * scale)
    yy, xx = np.meshgrid(np.arange(out_tile), np.arange(out_tile), indexing="ij")
    ramp = (((yy * 3 + xx * 5) % 17).astype(np.float32) * np.float32(1.0 / 128.0)).astype(np.float32)
    gain = np.array([0.75, 0.8125, 0.875], np.float32)[None, :, None, None]

    def fn(x):
        n, c, h, w = x.shape
        o = x[:, :, off:h - off, off:w - off]
        o = np.repeat(np.repeat(o, scale, axis=2), scale, axis=3)
        assert o.shape[2] == out_tile, (o.shape, out_tile)
        return (o.astype(np.float32) * gain + ramp[None, None]).astype(np.float32)

    return fn


def replay_model(tiles_f16_nhwc, out_tile):
    """Ignores its input and returns the given tiles in call order ([count][outT][outT][3] fp16 values as f32): lets the
    reference's applyWeights + accumulate + convertTo run on chosen tile values (stitch golden)."""
    state = {"i": 0}

    def fn(x):
        n = x.shape[0]
        out = np.zeros((n, 3, out_tile, out_tile), np.float32)
        for b in range(n):
            if state["i"] < tiles_f16_nhwc.shape[0]:
                out[b] = tiles_f16_nhwc[state["i"]].astype(np.float32).transpose(2, 0, 1)
            state["i"] += 1
        return out

    return fn


# name: (W, H, tile, out_tile, scale, blend, seed) for the stitch replay goldens
STITCH_CASES = {
    "cfg1_like": (128, 96, 64, 56, 2, 1 / 16, 11),
    "blend8": (150, 90, 64, 56, 2, 1 / 8, 12),
    "blend0": (150, 90, 64, 56, 2, 0.0, 13),
    "cunet1x_t128": (333, 201, 128, 72, 1, 1 / 32, 14),
    "single": (20, 20, 64, 56, 2, 1 / 16, 15),
}


def stitch_tiles(count, out_tile, seed):
    return np.random.default_rng(2000 + seed).uniform(-0.1, 1.1, size=(count, out_tile, out_tile, 3)).astype(np.float16)


# (W, H, tile, seed) for padRoi + applyAugmentation + blobFromImages goldens: rects from the cfg-like grid, aug = i % 8
UNPACK_CASES = {
    "t64_ragged": (97, 61, 64, 56, 2, 1 / 8, 21),
    "t64_single": (20, 20, 64, 56, 2, 1 / 16, 22),
    "t128": (150, 131, 128, 72, 1, 1 / 32, 23),
}

# device name, BuildConfig ints (precision 1 = FP16) for getConfigHash / serializeConfig
HASH_CASES = [
    ("NVIDIA B200", (0, 1, 1, 1, 4, 3, 3, 3, 64, 256, 640, 64, 256, 640)),
    ("NVIDIA B200", (0, 1, 8, 8, 8, 3, 3, 3, 256, 256, 256, 256, 256, 256)),
    ("NVIDIA GeForce RTX 4090", (0, 0, 1, 2, 4, 3, 3, 3, 64, 128, 400, 64, 128, 400)),
    ("Tesla  T4", (0, 1, 1, 1, 1, 3, 3, 3, 64, 64, 64, 64, 64, 64)),
]
