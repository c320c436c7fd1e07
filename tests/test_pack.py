"""ONNX import + weight packing + engine-file naming (host-only; no GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import onnx_io
from oracle.models import make_model


def test_onnx_writer_reader_roundtrip():
    m = make_model("cunet", 2)
    blob = onnx_io.export_cunet(m)
    nodes, inits = onnx_io.read_model(blob)
    convs = [n for n in nodes if n[0] in ("Conv", "ConvTranspose")]
    assert len(convs) == 30
    sd = m.state_dict()
    for k, v in sd.items():
        assert np.array_equal(inits[k], v.numpy()), k


def test_emitted_onnx_runs_in_opencv_dnn():
    """The emitted file is a valid ONNX graph: OpenCV DNN (the only ONNX runtime in the image) reproduces torch."""
    cv2 = pytest.importorskip("cv2")
    import tempfile
    import torch
    m = make_model("cunet", 2)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.onnx")
        onnx_io.export_cunet(m, p)
        net = cv2.dnn.readNetFromONNX(p)
        x = np.random.default_rng(0).random((1, 3, 64, 64), dtype=np.float32)
        net.setInput(x)
        y = net.forward()
        with torch.no_grad():
            yt = m(torch.from_numpy(x)).numpy()
        assert y.shape == (1, 3, 56, 56)
        assert np.abs(y - yt).max() < 1e-4


@pytest.mark.parametrize("scale,arch,offset", [(1, 1, 28), (2, 2, 36)])
def test_pack_onnx(scale, arch, offset, built_lib, models_dir, tmp_path):
    import w2x
    _, path = models_dir[1][scale]
    out = str(tmp_path / "m.w2x")
    w2x.pack_onnx(path, out)
    info = w2x.pack_info(out)
    assert info == dict(arch=arch, scale=scale, offset=offset, layers=22)


@pytest.mark.parametrize("scale,layers", [(1, 105), (2, 105), (4, 106)])
def test_pack_swin_unet(scale, layers, built_lib, tmp_path):
    """swin_unet/*: 2 patch convs + 14 blocks x 7 records + 2 PatchDown + 2 (3 for 4x) PatchUp + ToImage."""
    import w2x
    from oracle.models import make_model
    m = make_model("swin_unet", scale)
    p = str(tmp_path / "swin.onnx")
    onnx_io.export_swin(m, p)
    nodes, inits = onnx_io.read_model(open(p, "rb").read())
    assert sum(1 for n in nodes if n[0] == "LayerNormalization") == 28
    assert inits["swin1.block.0.attn.relative_position_bias"].shape == (1, 6, 36, 36)
    out = str(tmp_path / "swin.w2x")
    w2x.pack_onnx(p, out)
    assert w2x.pack_info(out) == dict(arch=3, scale=scale, offset=8 * scale, layers=layers)


def test_pack_rejects_garbage(built_lib, tmp_path):
    import w2x
    p = tmp_path / "bad.onnx"
    p.write_bytes(b"\x08\x07\x12\x03abc")
    with pytest.raises(RuntimeError):
        w2x.pack_onnx(str(p), str(tmp_path / "o.w2x"))
    # a valid graph that is not a cunet
    blob = onnx_io.model([onnx_io.node("Relu", ["x"], ["y"])], [], [onnx_io.value_info("x", [1, 3, 8, 8])], [onnx_io.value_info("y", [1, 3, 8, 8])])
    p.write_bytes(blob)
    with pytest.raises(RuntimeError, match="template"):
        w2x.pack_onnx(str(p), str(tmp_path / "o.w2x"))


def test_config_hash_matches_reference_format(built_lib):
    """getConfigHash, img2img_build.cpp:8-27: sha256 of '<NameNoSpaces>.<FP16|TF32>.minB.optB.maxB.minC...maxH'."""
    import w2x
    cfg = w2x.BuildConfig.fixed(8, 256)
    s = "NVIDIAB200.FP16.8.8.8.3.3.3.256.256.256.256.256.256"
    assert w2x.config_hash("NVIDIA B200", cfg) == hashlib.sha256(s.encode()).hexdigest()
    cfg2 = w2x.BuildConfig(precision=w2x.PRECISION_TF32)
    s2 = "NVIDIAGeForceRTX4090.TF32.1.1.4.3.3.3.64.256.640.64.256.640"
    assert w2x.config_hash("NVIDIA GeForce RTX 4090", cfg2) == hashlib.sha256(s2.encode()).hexdigest()


def test_model_path_rule():
    """src/main.cpp:201-204, including the trailing underscore for scale 1."""
    import w2x
    assert w2x.model_path("models", "cunet/art", 3, 2) == "models/cunet/art/noise3_scale2x.onnx"
    assert w2x.model_path("models", "cunet/art", -1, 2) == "models/cunet/art/scale2x.onnx"
    assert w2x.model_path("models", "swin_unet/photo", 0, 1) == "models/swin_unet/photo/noise0_.onnx"


# ---- importer tolerance / strictness (ADVICE: "same conv shapes, different topology packs silently") ----------
def test_pack_accepts_equivalent_cunet_exports(built_lib, tmp_path):
    """Crops written as negative Pad nodes and initializers in another order describe the same network: identical pack."""
    import w2x
    from oracle import onnx_io
    from oracle.models import make_model
    m = make_model("cunet", 2, 7)
    blobs = {}
    for mut in ("", "pad", "shuffle"):
        p, q = str(tmp_path / f"m_{mut}.onnx"), str(tmp_path / f"m_{mut}.w2x")
        onnx_io.export_cunet(m, p, mutate=mut)
        w2x.pack_onnx(p, q)
        blobs[mut] = open(q, "rb").read()
    assert blobs[""] == blobs["pad"] == blobs["shuffle"]


@pytest.mark.parametrize("mut,needle", [("alpha", "alpha"), ("crop", "crops"), ("noclip", "Clip")])
def test_pack_rejects_cunet_with_different_topology(mut, needle, built_lib, tmp_path):
    """Same weight shapes, different network (LeakyReLU slope, a skip crop, the final clamp): the build must fail loudly."""
    import w2x
    from oracle import onnx_io
    from oracle.models import make_model
    p = str(tmp_path / "m.onnx")
    onnx_io.export_cunet(make_model("cunet", 2, 7), p, mutate=mut)
    with pytest.raises(RuntimeError) as ei:
        w2x.pack_onnx(p, str(tmp_path / "m.w2x"))
    assert needle in str(ei.value), str(ei.value)


def test_pack_rejects_grouped_and_dilated_convs(built_lib, tmp_path):
    import w2x
    from oracle import onnx_io
    from oracle.models import make_model
    m = make_model("cunet", 1, 3)
    blob = onnx_io.export_cunet(m)
    # flip `group` 1 -> 2 in the first Conv node: attribute "group" is written as name(field 1) + i(field 3) + type(field 20)
    needle = b"\x0a\x05group\x18\x01"
    assert needle in blob
    bad = blob.replace(needle, b"\x0a\x05group\x18\x02", 1)
    p = str(tmp_path / "g.onnx")
    open(p, "wb").write(bad)
    with pytest.raises(RuntimeError) as ei:
        w2x.pack_onnx(p, str(tmp_path / "g.w2x"))
    assert "group" in str(ei.value)


@pytest.mark.parametrize("scale", [2, 4])
def test_pack_swin_decomposed_export_equals_fused_export(scale, built_lib, tmp_path):
    """LayerNorm as primitive ops, Linear as Gemm(transB=1), shuffled initializers, opset 13 -> byte-identical packed weights."""
    import w2x
    from oracle import onnx_io
    from oracle.models import make_model
    m = make_model("swin_unet", scale, 5)
    out = {}
    for variant in ("", "decomposed"):
        p, q = str(tmp_path / f"s_{variant}.onnx"), str(tmp_path / f"s_{variant}.w2x")
        onnx_io.export_swin(m, p, variant=variant)
        w2x.pack_onnx(p, q)
        out[variant] = open(q, "rb").read()
    assert w2x.pack_info(str(tmp_path / "s_.w2x")) == w2x.pack_info(str(tmp_path / "s_decomposed.w2x"))
    assert out[""] == out["decomposed"]
