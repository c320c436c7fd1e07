"""ctypes binding of oracle/_ref/libw2xref.so (TEST INFRASTRUCTURE ONLY).

libw2xref.so is the reference's OWN src/tensorrt/*.cpp compiled unmodified from /root/reference (recipe: oracle/Makefile)
against CPU mocks of the libraries that are absent from the image (oracle/ref_shim/: OpenCV-CUDA subset, TensorRT, CUDA
runtime, nlohmann-json).  The neural network is a callback; everything else -- calculateTiles, padRoi, applyWeights,
apply/reverseAugmentation, createTileWeights, blobFromImages, getConfigHash, serializeConfig, getEnginePath,
Img2Img::build/load/render -- is reference code executing here.  It pins the NumPy restatement (oracle/tiling.py) and the
product's host logic; it is never imported by the product.

`available()` is False when the library has not been built (no /root/reference and no prebuilt copy): callers then fall
back to the committed vectors under tests/golden/ that tests/golden/make_ref_goldens.py generated from it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libw2xref.so")

MODEL_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_void_p)


class CBuild(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("deviceId", "precision", "minB", "optB", "maxB", "minC", "optC", "maxC", "minW", "optW", "maxW", "minH", "optH", "maxH")]


class CRender(C.Structure):
    _fields_ = [("deviceId", C.c_int), ("precision", C.c_int), ("batchSize", C.c_int), ("channels", C.c_int), ("height", C.c_int), ("width", C.c_int),
                ("scaling", C.c_int), ("overlapX", C.c_double), ("overlapY", C.c_double), ("tta", C.c_int)]


def build_config(batch=(1, 1, 4), channels=(3, 3, 3), width=(64, 256, 640), height=(64, 256, 640), precision=1, device=0) -> CBuild:
    """trt::BuildConfig with the reference's defaults (config.h:12-31); precision 1 = FP16, 0 = TF32."""
    return CBuild(device, precision, *batch, *channels, *width, *height)


def render_config(batch=1, tile=256, scaling=4, overlap=(0.0625, 0.0625), tta=False, precision=1, device=0, channels=3, height=None) -> CRender:
    return CRender(device, precision, batch, channels, tile if height is None else height, tile, scaling, overlap[0], overlap[1], int(tta))


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} is not built (run `make -C oracle`; needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        ip, fp, u8p = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        L.ref_shim_set_model.argtypes = [C.c_int, C.c_int, MODEL_FN, C.c_void_p]
        L.ref_shim_set_device_name.argtypes = [C.c_char_p]
        L.ref_shim_set_pitch_align.argtypes = [C.c_int]
        L.ref_calculate_tiles.argtypes = [C.c_int] * 9 + [C.c_double, C.c_double, ip, ip, C.c_int]
        L.ref_pad_roi.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.ref_apply_augmentation.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.ref_reverse_augmentation.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
        L.ref_create_tile_weights.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, fp]
        L.ref_apply_weights.argtypes = [fp] + [C.c_int] * 9
        L.ref_blob_from_images.argtypes = [u8p, C.c_int, C.c_int, fp]
        L.ref_config_hash.argtypes = [C.POINTER(CBuild), C.c_char_p]
        L.ref_serialize_config.argtypes = [C.c_char_p, C.POINTER(CBuild)]
        L.ref_is_compatible.argtypes = [C.POINTER(CRender), C.POINTER(CBuild)]
        L.ref_is_optimized.argtypes = [C.POINTER(CRender), C.POINTER(CBuild)]
        L.ref_get_engine_path.argtypes = [C.c_char_p, C.POINTER(CRender), C.c_char_p, C.c_size_t]
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_log.argtypes = [C.c_void_p]
        L.ref_log.restype = C.c_char_p
        L.ref_build.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(CBuild)]
        L.ref_load.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(CRender)]
        L.ref_render.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, u8p, C.c_size_t, ip, ip]
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def set_device_name(name: str) -> None:
    lib().ref_shim_set_device_name(name.encode())


def set_pitch_align(nbytes: int) -> None:
    """Row pitch alignment of fresh mock GpuMat allocations (cudaMallocPitch: 512 on current GPUs); 1 = dense."""
    lib().ref_shim_set_pitch_align(nbytes)


def calculate_tiles(in_w, in_h, out_w, out_h, tile_w, tile_h, out_tile_w, out_tile_h, scaling, ov_x, ov_y, cap=1 << 12):
    """calculateTiles (img2img_render.cpp:7-66) -> (tileCount, in_rects, out_rects) as lists of (x, y, w, h)."""
    sentinel = -(1 << 31)
    a = (C.c_int * (4 * cap))(*([sentinel] * (4 * cap)))
    b = (C.c_int * (4 * cap))()
    n = lib().ref_calculate_tiles(in_w, in_h, out_w, out_h, tile_w, tile_h, out_tile_w, out_tile_h, scaling, ov_x, ov_y, a, b, cap)
    m = 0  # rects actually emitted (tileCount can disagree with the vectors' size for degenerate grids)
    while m < cap and a[4 * m + 2] != sentinel:
        m += 1
    if n > cap and m == cap:
        return calculate_tiles(in_w, in_h, out_w, out_h, tile_w, tile_h, out_tile_w, out_tile_h, scaling, ov_x, ov_y, cap=n)
    return n, [tuple(a[4 * i:4 * i + 4]) for i in range(m)], [tuple(b[4 * i:4 * i + 4]) for i in range(m)]


def pad_roi(img: np.ndarray, rect) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    x, y, w, h = rect
    out = np.zeros((h, w, 3), np.uint8)
    assert lib().ref_pad_roi(_u8(img), img.shape[1], img.shape[0], x, y, w, h, _u8(out))
    return out


def apply_augmentation(tile: np.ndarray, k: int) -> np.ndarray:
    tile = np.ascontiguousarray(tile, np.uint8)
    out = np.zeros_like(tile)
    assert lib().ref_apply_augmentation(_u8(tile), tile.shape[0], k, _u8(out))
    return out


def reverse_augmentation(tile: np.ndarray, k: int, aliased: bool = False) -> np.ndarray:
    tile = np.ascontiguousarray(tile, np.float32)
    out = np.zeros_like(tile)
    assert lib().ref_reverse_augmentation(_f32(tile), tile.shape[0], k, int(aliased), _f32(out))
    return out


def create_tile_weights(ov_x: int, ov_y: int, w: int, h: int) -> np.ndarray:
    """[4][h][w][3] f32 in the reference's array order: top, right, bottom, left."""
    out = np.zeros((4, h, w, 3), np.float32)
    assert lib().ref_create_tile_weights(ov_x, ov_y, w, h, _f32(out))
    return out


def apply_weights(tile: np.ndarray, ov_x: int, ov_y: int, rect, canvas_w: int, canvas_h: int) -> np.ndarray:
    t = np.ascontiguousarray(tile, np.float32).copy()
    assert lib().ref_apply_weights(_f32(t), t.shape[0], ov_x, ov_y, *rect, canvas_w, canvas_h)
    return t


def blob_from_images(tiles: np.ndarray) -> np.ndarray:
    """[n][T][T][3] u8 -> [n][3][T][T] f32 (blobFromImages + the linear copy of infer())."""
    tiles = np.ascontiguousarray(tiles, np.uint8)
    n, t = tiles.shape[0], tiles.shape[1]
    out = np.zeros((n, 3, t, t), np.float32)
    assert lib().ref_blob_from_images(_u8(tiles), n, t, _f32(out))
    return out


def config_hash(cfg: CBuild) -> str:
    buf = C.create_string_buffer(65)
    lib().ref_config_hash(C.byref(cfg), buf)
    return buf.value.decode()


def serialize_config(path: str, cfg: CBuild) -> None:
    assert lib().ref_serialize_config(path.encode(), C.byref(cfg))


def is_compatible(r: CRender, b: CBuild) -> bool:
    return bool(lib().ref_is_compatible(C.byref(r), C.byref(b)))


def is_optimized(r: CRender, b: CBuild) -> bool:
    return bool(lib().ref_is_optimized(C.byref(r), C.byref(b)))


def get_engine_path(model_path: str, r: CRender) -> Tuple[bool, str]:
    buf = C.create_string_buffer(4096)
    ok = lib().ref_get_engine_path(model_path.encode(), C.byref(r), buf, 4096)
    return bool(ok), buf.value.decode()


class Img2Img:
    """trt::Img2Img (img2img.h:14-50) with the network replaced by `model(x[B,3,T,T] f32) -> [B,3,outT,outT] f32`."""

    def __init__(self, model: Callable[[np.ndarray], np.ndarray], scale: int, out_minus: int):
        self._model = model
        self.calls: List[Tuple[int, ...]] = []

        def tramp(pin, n, c, h, w, pout, oh, ow, _user):
            x = np.ctypeslib.as_array(pin, shape=(n, c, h, w))
            y = np.ascontiguousarray(self._model(x.copy()), np.float32)
            assert y.shape == (n, c, oh, ow), (y.shape, (n, c, oh, ow))
            C.memmove(pout, y.ctypes.data, y.nbytes)

        self._tramp = MODEL_FN(tramp)
        self._scale, self._out_minus = scale, out_minus
        self._h = lib().ref_create()

    def _bind(self):
        lib().ref_shim_set_model(self._scale, self._out_minus, self._tramp, None)

    @property
    def log(self) -> str:
        return lib().ref_log(self._h).decode()

    def build(self, onnx_path: str, cfg: CBuild) -> bool:
        self._bind()
        return bool(lib().ref_build(self._h, onnx_path.encode(), C.byref(cfg)))

    def load(self, onnx_path: str, cfg: CRender) -> bool:
        self._bind()
        self._scaling = cfg.scaling
        return bool(lib().ref_load(self._h, onnx_path.encode(), C.byref(cfg)))

    def render(self, src_bgr: np.ndarray) -> Optional[np.ndarray]:
        self._bind()
        src = np.ascontiguousarray(src_bgr, np.uint8)
        h, w = src.shape[:2]
        s = self._scaling
        dst = np.zeros((h * s, w * s, 3), np.uint8)
        ow, oh = C.c_int(), C.c_int()
        if not lib().ref_render(self._h, _u8(src), w, h, _u8(dst), dst.nbytes, C.byref(ow), C.byref(oh)):
            return None
        assert (ow.value, oh.value) == (w * s, h * s)
        return dst

    def close(self):
        if self._h:
            lib().ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(src_bgr: np.ndarray, model, tile: int, out_tile: int, scaling: int, overlap: float, batch: int = 1, tta: bool = False,
           workdir: Optional[str] = None) -> np.ndarray:
    """Img2Img::build + load + render of the reference with `model` as the engine; same signature as oracle.tiling.render."""
    import tempfile
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        onnx = os.path.join(d, "model.onnx")
        open(onnx, "wb").write(b"not a real onnx file: the mock parser only checks that it exists")
        eng = Img2Img(model, scaling, tile * scaling - out_tile)
        try:
            assert eng.build(onnx, build_config(batch=(batch, batch, batch), width=(tile, tile, tile), height=(tile, tile, tile))), eng.log
            assert eng.load(onnx, render_config(batch=batch, tile=tile, scaling=scaling, overlap=(overlap, overlap), tta=tta)), eng.log
            out = eng.render(src_bgr)
            assert out is not None, eng.log
            return out
        finally:
            eng.close()
