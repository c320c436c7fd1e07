"""Tile grid / blend ramp: oracle (NumPy) vs oracle (C) vs product (C ABI) vs golden vectors.  Bit-exact integers.
Mirrors the host arithmetic of /root/reference/src/tensorrt/img2img_render.cpp:7-66 and img2img_load.cpp:29-52."""
import ctypes as C
import json
import os

import numpy as np
import pytest
from hypothesis import assume, given, settings, strategies as st

from oracle import tiling

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tile_grid.json")))


class _R(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]


def c_oracle_tiles(lib, W, H, T, OT, S, B):
    info = (C.c_int * 8)()
    lib.orc_calculate_tiles.argtypes = [C.c_int] * 9 + [C.c_double, C.c_double, C.POINTER(_R), C.POINTER(_R), C.c_int, C.POINTER(C.c_int)]
    n = lib.orc_calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B, None, None, 0, info)
    a, b = (_R * max(n, 1))(), (_R * max(n, 1))()
    lib.orc_calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B, a, b, n, info)
    return n, list(info), [(r.x, r.y, r.w, r.h) for r in a[:n]], [(r.x, r.y, r.w, r.h) for r in b[:n]]


# Hand-derived anchors, SURVEY.md 8a row a2 (count, nx, ny, sIn, iov, oov, last output rect)
ANCHORS = {
    "cfg1_cunet2x_t64_256x256": (121, 11, 11, 28, 4, 8, (480, 480, 32, 32)),
    "cfg2_cunet2x_t256_1080p": (60, 10, 6, 220, 16, 32, (3672, 2040, 168, 120)),
    "cfg3_cunet1x_t400_1080p": (24, 6, 4, 344, 25, 25, None),
    "cfg4_swin4x_t256_1080p": (45, 9, 5, 240, 16, 64, (7168, 3584, 512, 736)),
    "cfg5_swin4x_t256_960x540": (15, 5, 3, 240, 16, 64, None),
}


@pytest.mark.parametrize("name", sorted(ANCHORS))
def test_oracle_matches_hand_derived_anchors(name):
    W, H, T, OT, S, B = GOLDEN[name]["args"]
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    count, nx, ny, sin, iov, oov, last = ANCHORS[name]
    assert (g.count, g.nx, g.ny) == (count, nx, ny)
    assert g.scaled_in == (sin, sin) and g.in_overlap == (iov, iov) and g.out_overlap == (oov, oov)
    if last:
        assert g.out_rects[-1] == last
    border = (T - sin) // 2
    assert g.in_rects[0] == (-border, -border, T, T)


def test_q5_drift_is_replicated():
    """tile 400 / blend 1/32 / scale 2 -> iov = lround(12.5) = 13 but oov = 25 != 2*13."""
    g = GOLDEN["q5_cunet2x_t400_blend32_drift"]
    assert g["in_overlap"] == [13, 13] and g["out_overlap"] == [25, 25]
    assert g["in_rects"][g["ny"]][0] - g["in_rects"][0][0] == 364 - 13
    assert g["out_rects"][g["ny"]][0] == 728 - 25  # 703, not 2 * 351 = 702


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_golden_three_way(name, built_lib, oracle_c):
    import w2x
    W, H, T, OT, S, B = GOLDEN[name]["args"]
    gold = GOLDEN[name]
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    assert g.count == gold["count"] and [list(r) for r in g.in_rects] == gold["in_rects"] and [list(r) for r in g.out_rects] == gold["out_rects"]
    n, info, ir, orr = c_oracle_tiles(oracle_c, W, H, T, OT, S, B)
    assert n == gold["count"] and [list(r) for r in ir] == gold["in_rects"] and [list(r) for r in orr] == gold["out_rects"]
    n2, grid, ir2, or2 = w2x.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    assert n2 == gold["count"] and [list(r) for r in ir2] == gold["in_rects"] and [list(r) for r in or2] == gold["out_rects"]
    assert grid == [gold["nx"], gold["ny"], *gold["scaled_in"], *gold["in_overlap"], *gold["out_overlap"]] == info


@settings(max_examples=200, deadline=None)
@given(W=st.integers(16, 4000), H=st.integers(16, 2500), T=st.sampled_from([64, 128, 256, 400, 640]),
       S=st.sampled_from([1, 2, 4]), B=st.sampled_from([0.0, 1 / 32, 1 / 16, 1 / 8]), fam=st.sampled_from(["cunet", "swin"]))
def test_product_equals_oracle_on_random_frames(W, H, T, S, B, fam, built_lib):
    import w2x
    if fam == "cunet":
        if S == 4:
            S = 2
        OT = T - 56 if S == 1 else 2 * T - 72
    else:
        OT = (T - 16) * S
    # the reference divides by (scaledInputTile - inputOverlap): configurations where that is <= 0 are invalid there too
    assume(tiling.lround(OT / (T * S) * T) - tiling.lround(T * B) > 0)
    g = tiling.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    n, grid, ir, orr = w2x.calculate_tiles(W, H, W * S, H * S, T, T, OT, OT, S, B, B)
    assert n == g.count and ir == g.in_rects and orr == g.out_rects


def test_every_output_pixel_is_covered_and_weights_sum_to_one():
    """Closed-form property the blend relies on: overlapping ramp pairs sum to 1 (createTileWeights), so a constant
    model output reproduces the constant: checks grid + weights + predicates together."""
    for name in ("cfg1_cunet2x_t64_256x256", "blend8_cunet1x_t128", "blend0_cunet2x_t64"):
        W, H, T, OT, S, B = GOLDEN[name]["args"]
        out = tiling.render(np.zeros((H, W, 3), np.uint8), lambda x: np.full((x.shape[0], 3, OT, OT), 0.5, np.float32), T, OT, S, B)
        assert out.min() >= 127 and out.max() <= 128, name


@pytest.mark.parametrize("ov", [0, 1, 8, 25, 32, 64])
def test_blend_ramp_three_way(ov, built_lib, oracle_c):
    import w2x
    w = tiling.create_tile_weights(ov, ov, 128, 128)
    ramp = w2x.blend_ramp(ov)
    assert ramp.shape[0] == ov
    buf = (C.c_float * max(ov, 1))()
    oracle_c.orc_blend_ramp.argtypes = [C.c_int, C.POINTER(C.c_float)]
    oracle_c.orc_blend_ramp(ov, buf)
    for r in range(ov):
        assert w[0][r, 0] == ramp[r] == buf[r] == np.float32((r + 1) / (ov + 1))
        assert w[3][0, r] == ramp[r] and w[2][127 - r, 0] == ramp[r] and w[1][0, 127 - r] == ramp[r]
    assert (w[0][ov:] == 1).all() and (w[3][:, ov:] == 1).all()


def test_augment_closed_forms_match_numpy():
    """The index remaps the CUDA kernels implement == flips/rot90 (D4 ops of img2img_render.cpp:123-222)."""
    n = 7
    a = np.arange(n * n).reshape(n, n)
    for k in range(8):
        aug = tiling.augment(a, k)
        back = tiling.reverse_augment(aug, k)
        assert (back == a).all(), k
        for r in range(n):
            for c in range(n):
                rr, cc = tiling.augment_src_index(k, r, c, n)
                assert aug[r, c] == a[rr, cc]
                rr, cc = tiling.reverse_src_index(k, r, c, n)
                assert back[r, c] == aug[rr, cc]
