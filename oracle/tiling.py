"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the tile -> model -> stitch path.

NumPy restatement of the reference's host/OpenCV-CUDA tiling arithmetic.  Nothing
on the product path may import this module: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs use it, as the checker.

Parity status: this restatement is PINNED against the reference's own code: oracle/_ref/libw2xref.so is
/root/reference/src/tensorrt/*.cpp compiled unmodified against CPU mocks of the absent libraries (oracle/Makefile,
oracle/ref_shim/), and tests/test_ref_parity.py checks every function below, and the whole render() loop with analytic
stand-in networks, bit-for-bit against it (live when the library is built, else against tests/golden/ref_goldens.npz and
tile_grid.json that tests/golden/make_ref_goldens.py generated from it).  Values that pass through the NEURAL NETWORK stay
"parity unpinned" against the reference (TensorRT engine + nunif ONNX files are unobtainable): they are
checked against the fp32 PyTorch graphs in oracle/models.py with BASELINE.json's tolerance.

Reference citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Sequence, Tuple

import numpy as np

Rect = Tuple[int, int, int, int]  # x, y, w, h  (cv::Rect2i)


def lround(x: float) -> int:
    """std::lround: round half away from zero (src/tensorrt/img2img_render.cpp:17-33)."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def c_div(a: int, b: int) -> int:
    """C++ integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


@dataclass
class TileGrid:
    count: int
    nx: int
    ny: int
    scaled_in: Tuple[int, int]      # scaledInputTileSize (w, h)
    in_overlap: Tuple[int, int]     # inputOverlap (x, y)
    out_overlap: Tuple[int, int]    # scaledOutputOverlap (x, y)
    in_rects: List[Rect]
    out_rects: List[Rect]


def calculate_tiles(in_w: int, in_h: int, out_w: int, out_h: int,
                    tile_w: int, tile_h: int, out_tile_w: int, out_tile_h: int,
                    scaling: int, overlap_x: float, overlap_y: float) -> TileGrid:
    """calculateTiles, src/tensorrt/img2img_render.cpp:7-66.

    Note the reference uses inputTileSize.width for BOTH dims of
    scaledOutputTileSize (:11-14); kept.  Outer loop x, inner loop y (:43-44).
    """
    sot_w = tile_w * scaling
    sot_h = tile_w * scaling                      # sic (:13)
    sin_w = lround(out_tile_w / sot_w * tile_w)   # :17
    sin_h = lround(out_tile_h / sot_h * tile_h)   # :18
    iov_x = lround(tile_w * overlap_x)            # :22
    iov_y = lround(tile_h * overlap_y)            # :23
    oov_x = lround(sot_w * overlap_x)             # :27
    oov_y = lround(sot_h * overlap_y)             # :28
    nx = lround(math.ceil((in_w - iov_x) / (sin_w - iov_x)))  # :32
    ny = lround(math.ceil((in_h - iov_y) / (sin_h - iov_y)))  # :33
    in_rects: List[Rect] = []
    out_rects: List[Rect] = []
    for i in range(nx):
        for j in range(ny):
            in_rects.append((
                -c_div(tile_w - sin_w, 2) + i * sin_w - i * iov_x,   # :47
                -c_div(tile_h - sin_h, 2) + j * sin_h - j * iov_y,   # :48
                tile_w, tile_h))
            x = i * out_tile_w - i * oov_x                           # :54
            y = j * out_tile_h - j * oov_y                           # :55
            out_rects.append((
                x, y,
                out_w - x if x + out_tile_w > out_w else out_tile_w,   # :59
                out_h - y if y + out_tile_h > out_h else out_tile_h))  # :60
    return TileGrid(nx * ny, nx, ny, (sin_w, sin_h), (iov_x, iov_y), (oov_x, oov_y),
                    in_rects, out_rects)


def pad_roi(img: np.ndarray, rect: Rect) -> np.ndarray:
    """padRoi, src/tensorrt/img2img_render.cpp:68-105: ROI view, or
    copyMakeBorder(BORDER_REPLICATE) of the in-bounds part == clamp indexing."""
    x, y, w, h = rect
    H, W = img.shape[:2]
    ys = np.clip(np.arange(y, y + h), 0, H - 1)
    xs = np.clip(np.arange(x, x + w), 0, W - 1)
    return img[ys][:, xs]


def create_tile_weights(oov_x: int, oov_y: int, out_tile_w: int, out_tile_h: int):
    """createTileWeights, src/tensorrt/img2img_load.cpp:29-52 (called :262-269).

    Returns [top, right, bottom, left] as float32 [H, W] images (the reference's
    3 channels are identical).  alpha is computed in double and stored as f32."""
    top = np.ones((out_tile_h, out_tile_w), np.float32)
    left = np.ones((out_tile_h, out_tile_w), np.float32)
    hh = oov_y + 1
    for i in range(1, hh):
        top[i - 1, :] = np.float32(float(i) / hh)      # :35-38
    ww = oov_x + 1
    for i in range(1, ww):
        left[:, i - 1] = np.float32(float(i) / ww)     # :42-45
    bottom = top[::-1, :].copy()                       # :48 flip code 0
    right = left[:, ::-1].copy()                       # :51 flip code 1
    return [top, right, bottom, left]


def apply_weights(tile: np.ndarray, rect: Rect, canvas_w: int, canvas_h: int, weights) -> np.ndarray:
    """applyWeights, src/tensorrt/img2img_render.cpp:107-121.  `rect` is the CLIPPED
    output rect; the canvas rect origin is (0,0).  Sequential in-place f32 multiplies in
    the order left, top, right, bottom."""
    x, y, w, h = rect
    t = tile.astype(np.float32, copy=True)
    if x > 0:
        t = t * weights[3][..., None]
    if y > 0:
        t = t * weights[0][..., None]
    if x + w < canvas_w:
        t = t * weights[1][..., None]
    if y + h < canvas_h:
        t = t * weights[2][..., None]
    return t


# --- D4 augmentations -------------------------------------------------------------------
# src/tensorrt/img2img_render.cpp:123-222.  Defined by OpenCV flip CODE (enum names in the
# reference are swapped): code 0 reverses rows, code 1 reverses columns.
# cv::cuda::rotate == nppiRotate about (0,0): x' = c*x + s*y + shiftX, y' = -s*x + c*y + shiftY
# => angle 90 with shift (0, H-1) is a counter-clockwise quarter turn == np.rot90(a, 1).

def augment(a: np.ndarray, k: int) -> np.ndarray:
    """applyAugmentation (:134-177) on an HWC (or HW) array."""
    if k == 0:
        return a
    if k == 1:
        return a[::-1]
    if k == 2:
        return a[:, ::-1]
    if k == 3:
        return np.rot90(a, 1)
    if k == 4:
        return np.rot90(a, 2)
    if k == 5:
        return np.rot90(a, 3)
    if k == 6:
        return np.rot90(a[::-1], 1)
    if k == 7:
        return np.rot90(a[:, ::-1], 1)
    raise ValueError(k)


def reverse_augment(a: np.ndarray, k: int) -> np.ndarray:
    """reverseAugmentation (:179-222) with the aliasing bug fixed: the true inverse."""
    if k == 0:
        return a
    if k == 1:
        return a[::-1]
    if k == 2:
        return a[:, ::-1]
    if k == 3:
        return np.rot90(a, 3)
    if k == 4:
        return np.rot90(a, 2)
    if k == 5:
        return np.rot90(a, 1)
    if k == 6:
        return np.rot90(a, 3)[::-1]
    if k == 7:
        return np.rot90(a, 3)[:, ::-1]
    raise ValueError(k)


def augment_src_index(k: int, r: int, c: int, n: int) -> Tuple[int, int]:
    """For an n x n tile: augmented[r][c] == original[rr][cc]; returns (rr, cc).
    Closed form the CUDA unpack kernel implements (checked against augment() in tests)."""
    m = n - 1
    return {
        0: (r, c),
        1: (m - r, c),
        2: (r, m - c),
        3: (c, m - r),
        4: (m - r, m - c),
        5: (m - c, r),
        6: (m - c, m - r),
        7: (c, r),
    }[k]


def reverse_src_index(k: int, r: int, c: int, n: int) -> Tuple[int, int]:
    """For an n x n model output: deaugmented[r][c] == model_out[rr][cc]; returns (rr, cc)."""
    m = n - 1
    return {
        0: (r, c),
        1: (m - r, c),
        2: (r, m - c),
        3: (m - c, r),
        4: (m - r, m - c),
        5: (c, m - r),
        6: (m - c, m - r),
        7: (c, r),
    }[k]


def normalize_u8(tile_u8: np.ndarray) -> np.ndarray:
    """blobFromImages, src/tensorrt/img2img_infer.cpp:19: f32(u8) * f32(1/255)."""
    return tile_u8.astype(np.float32) * np.float32(1.0 / 255.0)


def pack_u8(canvas: np.ndarray) -> np.ndarray:
    """output.convertTo(CV_8UC3, 255.0), src/tensorrt/img2img_render.cpp:342: OpenCV CUDA
    saturate_cast<uchar>(float) == round-to-nearest-even then clamp."""
    v = canvas.astype(np.float32) * np.float32(255.0)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


ModelFn = Callable[[np.ndarray], np.ndarray]  # [B,3,T,T] f32 -> [B,3,outT,outT] f32


def render(src_bgr: np.ndarray, model: ModelFn, tile: int, out_tile: int, scaling: int,
           overlap: float, batch: int = 1, tta: bool = False,
           model_batch: int | None = None, tta_mode: str = "mean") -> np.ndarray:
    """trt::Img2Img::render, src/tensorrt/img2img_render.cpp:224-348.  `batch` only groups model calls; padding slots
    (:281) are zero tiles whose outputs are discarded (:298-299).

    tta_mode "mean" (default, what the product implements): the 8-way average the code computes at :305-314.
    tta_mode "reference_q1": what the reference actually adds to the canvas -- `outputTile = &tmpOutputMat` (:315), i.e.
    the LAST de-augmented tile, the average being dropped; used to pin this restatement against oracle/_ref."""
    H, W = src_bgr.shape[:2]
    rgb = src_bgr[..., ::-1]                                        # :227
    oh, ow = H * scaling, W * scaling
    canvas = np.zeros((oh, ow, 3), np.float32)                      # :228-229
    g = calculate_tiles(W, H, ow, oh, tile, tile, out_tile, out_tile, scaling, overlap, overlap)
    overlapping = overlap != 0
    weights = create_tile_weights(g.out_overlap[0], g.out_overlap[1], out_tile, out_tile) if overlapping else None
    steps_per_tile = 8 if tta else 1
    batch_count = lround(math.ceil(g.count * steps_per_tile / batch))   # :249
    step_count = batch_count * batch
    pending: List[Tuple[int, int]] = []
    inputs: List[np.ndarray] = []
    acc = None
    for step in range(step_count):
        ti, aug = step // steps_per_tile, step % steps_per_tile
        pending.append((ti, aug))
        if ti < g.count:
            t = pad_roi(rgb, g.in_rects[ti])
            if tta and aug != 0:
                t = augment(t, aug)
            inputs.append(normalize_u8(t).transpose(2, 0, 1))       # infer.cpp:10-19 NCHW planar
        else:
            inputs.append(np.zeros((3, tile, tile), np.float32))    # :281
        if step % batch != batch - 1:
            continue
        outs = model(np.stack(inputs))
        for b in range(batch):
            ti, aug = pending[b]
            if ti >= g.count:
                break
            o = outs[b].transpose(1, 2, 0).astype(np.float32)       # infer.cpp:30-36
            if tta:
                if aug == 0:
                    acc = np.zeros_like(o)
                    acc = acc + o                                   # :307-308
                else:
                    acc = acc + reverse_augment(o, aug)             # :310-312
                if aug != 7:
                    continue
                if tta_mode == "reference_q1":
                    o = reverse_augment(o, aug)                     # :315 tmpOutputMat, the last de-augmented tile
                else:
                    o = acc * np.float32(1.0 / 8.0)                 # :314 (mean; q1)
            rect = g.out_rects[ti]
            if overlapping:
                o = apply_weights(o, rect, ow, oh, weights)         # :325-326
            x, y, w, h = rect
            canvas[y:y + h, x:x + w] += o[:h, :w]                   # :329-330
        pending.clear()
        inputs.clear()
    return pack_u8(canvas)[..., ::-1].copy()                        # :342-343


def synthetic_frame(w: int, h: int, seed: int) -> np.ndarray:
    """SURVEY 8d synthetic input: low-frequency field + N(0,8) grain, u8 BGR HWC."""
    rng = np.random.default_rng(seed)
    lw, lh = max(2, w // 16 + 2), max(2, h // 16 + 2)
    low = rng.uniform(0, 255, size=(lh, lw, 3)).astype(np.float32)
    # separable linear upsample of the 1/16-res field (cheap stand-in for bicubic)
    ys = np.linspace(0, lh - 1.001, h)
    xs = np.linspace(0, lw - 1.001, w)
    y0 = ys.astype(int); fy = (ys - y0)[:, None, None].astype(np.float32)
    x0 = xs.astype(int); fx = (xs - x0)[None, :, None].astype(np.float32)
    rows = low[y0] * (1 - fy) + low[y0 + 1] * fy
    img = rows[:, x0] * (1 - fx) + rows[:, x0 + 1] * fx
    img = img + rng.normal(0, 8, size=img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
