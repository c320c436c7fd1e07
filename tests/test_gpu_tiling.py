"""-m gpu: the memory-bound tiling kernels through the C ABI vs the NumPy oracle.  Bit-exact (integer / index work and
IEEE fp32 sequences reproduced op-for-op)."""
import numpy as np
import pytest

from oracle import tiling

pytestmark = pytest.mark.gpu


def _frame(w, h, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("w,h,tile,out_tile,scale,blend", [
    (256, 256, 64, 56, 2, 1 / 16),      # cfg1
    (97, 61, 64, 56, 2, 1 / 8),         # ragged: frame smaller than a tile in y
    (20, 20, 64, 56, 2, 1 / 16),        # single tile, all four borders replicated
    (333, 201, 128, 72, 1, 1 / 32),
])
def test_unpack_matches_padroi_and_blob(w, h, tile, out_tile, scale, blend, built_lib):
    import w2x
    src = _frame(w, h, 1)
    g = tiling.calculate_tiles(w, h, w * scale, h * scale, tile, tile, out_tile, out_tile, scale, blend, blend)
    rects = g.in_rects
    augs = [i % 8 for i in range(len(rects))]
    got = w2x.unpack_tiles(src, rects, augs, tile)
    rgb = src[..., ::-1]
    for i, (r, k) in enumerate(zip(rects, augs)):
        t = tiling.augment(tiling.pad_roi(rgb, r), k)
        ref = tiling.normalize_u8(t).astype(np.float16)  # f32(u8) * f32(1/255) then RN to fp16
        assert np.array_equal(got[i, :, :, :3].view(np.uint16), np.ascontiguousarray(ref).view(np.uint16)), (i, k)
        assert (got[i, :, :, 3] == 0).all()


def _stitch_oracle(tiles_f16, g, cw, ch, out_tile):
    canvas = np.zeros((ch, cw, 3), np.float32)
    ov = g.out_overlap
    weights = tiling.create_tile_weights(ov[0], ov[1], out_tile, out_tile) if (ov[0] or ov[1]) else None
    for i, rect in enumerate(g.out_rects):
        o = tiles_f16[i, :, :, :3].astype(np.float32)
        if weights is not None:
            o = tiling.apply_weights(o, rect, cw, ch, weights)
        x, y, w, h = rect
        canvas[y:y + h, x:x + w] += o[:h, :w]
    return tiling.pack_u8(canvas)[..., ::-1]


@pytest.mark.parametrize("w,h,tile,out_tile,scale,blend", [
    (256, 256, 64, 56, 2, 1 / 16),
    (150, 90, 64, 56, 2, 1 / 8),
    (150, 90, 64, 56, 2, 0.0),
    (333, 201, 128, 72, 1, 1 / 32),
    (20, 20, 64, 56, 2, 1 / 16),
])
def test_stitch_is_bit_exact(w, h, tile, out_tile, scale, blend, built_lib):
    import w2x
    g = tiling.calculate_tiles(w, h, w * scale, h * scale, tile, tile, out_tile, out_tile, scale, blend, blend)
    rng = np.random.default_rng(2)
    tiles = rng.uniform(-0.1, 1.1, size=(g.count, out_tile, out_tile, 4)).astype(np.float16)
    got = w2x.stitch_tiles(tiles, g.nx, g.ny, g.out_overlap[0], g.out_overlap[1], w * scale, h * scale)
    ref = _stitch_oracle(tiles, g, w * scale, h * scale, out_tile)
    assert np.array_equal(got, ref)


def test_stitch_rounding_is_nearest_even(built_lib):
    """convertTo(CV_8UC3, 255) == rint (ties to even) + saturate, img2img_render.cpp:342."""
    import w2x
    vals = np.array([0.5 / 255, 1.5 / 255, 2.5 / 255, -3.0, 7.0, 254.5 / 255, 1.0, 0.0], np.float16)
    tiles = np.zeros((1, 8, 8, 4), np.float16)
    tiles[0, 0, :, 0] = vals
    got = w2x.stitch_tiles(tiles, 1, 1, 0, 0, 8, 8)
    ref = tiling.pack_u8(tiles[0, :, :, :3].astype(np.float32))[..., ::-1]
    assert np.array_equal(got, ref)


def test_tta_reduce_is_bit_exact(built_lib):
    import w2x
    rng = np.random.default_rng(3)
    outs = rng.uniform(0, 1, size=(3, 8, 24, 24, 4)).astype(np.float16)
    got = w2x.tta_reduce(outs)
    for t in range(3):
        acc = np.zeros((24, 24, 3), np.float32)
        for k in range(8):
            acc = acc + tiling.reverse_augment(outs[t, k, :, :, :3].astype(np.float32), k)
        ref = acc * np.float32(0.125)
        assert np.array_equal(got[t, :, :, :3], ref), t


# ---- against vectors produced by the reference's own code (oracle/_ref, see tests/golden/make_ref_goldens.py) ----------------
import os  # noqa: E402

import refcases  # noqa: E402

_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_goldens.npz"))


@pytest.mark.parametrize("name", sorted(refcases.UNPACK_CASES))
def test_unpack_matches_reference_golden(name, built_lib):
    """unpack kernel == padRoi -> applyAugmentation -> blobFromImages of the reference (f32 result rounded to fp16), bit-exact."""
    import w2x
    W, H, T, OT, S, B, seed = refcases.UNPACK_CASES[name]
    src = refcases.frame(W, H, seed)
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    gold = _GOLD["unpack_" + name]  # [n][3][T][T] fp16, from RGB tiles
    got = w2x.unpack_tiles(src, g.in_rects, [i % 8 for i in range(g.count)], T)
    assert got.shape[0] == gold.shape[0]
    assert np.array_equal(np.ascontiguousarray(got[..., :3].transpose(0, 3, 1, 2)).view(np.uint16), gold.view(np.uint16))
    assert (got[..., 3] == 0).all()


@pytest.mark.parametrize("name", sorted(refcases.STITCH_CASES))
def test_stitch_matches_reference_golden(name, built_lib):
    """stitch kernel == the reference's applyWeights + canvas add + convertTo(8U, 255) + RGB2BGR on the same tile values."""
    import w2x
    W, H, T, OT, S, B, seed = refcases.STITCH_CASES[name]
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    tiles = np.zeros((g.count, OT, OT, 4), np.float16)
    tiles[..., :3] = refcases.stitch_tiles(g.count, OT, seed)
    got = w2x.stitch_tiles(tiles, g.nx, g.ny, g.out_overlap[0], g.out_overlap[1], W * S, H * S)
    assert np.array_equal(got, _GOLD["stitch_" + name])
