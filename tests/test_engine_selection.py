"""Engine artefact discovery (img2img_load.cpp:79-114 with img2img_build.cpp's naming): which
"<stem>_<sha256(cfg)[:16]>.w2x" + ".json" pair a `load` picks.  Host logic only -- runs without a GPU."""
import json
import os

import pytest

GPU = "NVIDIA B200"


def _artefact(dirpath, stem, cfg, device=GPU, sidecar=True, ext=".w2x"):
    import w2x
    h = w2x.config_hash(device, cfg)[:16]
    base = os.path.join(dirpath, f"{stem}_{h}")
    open(base + ext, "wb").write(b"placeholder")  # selection never opens the plan itself
    if sidecar:
        d = {"deviceName": device, "precision": "FP16" if cfg.precision == w2x.PRECISION_FP16 else "TF32"}
        for k in ("minBatchSize", "optBatchSize", "maxBatchSize", "minChannels", "optChannels", "maxChannels", "minWidth", "optWidth",
                  "maxWidth", "minHeight", "optHeight", "maxHeight"):
            d[k] = getattr(cfg, k)
        json.dump(d, open(base + ".json", "w"), indent=4)       # the reference writes it with nlohmann::json dump(4)
    return base + ext


def test_optimized_engine_wins_over_compatible(built_lib, tmp_path):
    import w2x
    d = str(tmp_path)
    model = os.path.join(d, "noise3_scale2x.onnx")
    open(model, "wb").write(b"onnx")
    wide = w2x.BuildConfig(minBatchSize=1, optBatchSize=4, maxBatchSize=8, minWidth=64, optWidth=128, maxWidth=640, minHeight=64, optHeight=128, maxHeight=640)
    exact = w2x.BuildConfig.fixed(8, 256)
    other = w2x.BuildConfig.fixed(4, 256)
    p_wide, p_exact, _ = _artefact(d, "noise3_scale2x", wide), _artefact(d, "noise3_scale2x", exact), _artefact(d, "noise3_scale2x", other)
    rc = w2x.RenderConfig(batchSize=8, height=256, width=256, scaling=2)
    assert w2x.select_engine(model, rc, GPU) == p_exact                      # opt == requested beats a merely compatible range
    rc6 = w2x.RenderConfig(batchSize=6, height=256, width=256, scaling=2)
    assert w2x.select_engine(model, rc6, GPU) == p_wide                      # only the wide profile admits batch 6
    with pytest.raises(RuntimeError, match="could not satisfy render configuration"):
        w2x.select_engine(model, w2x.RenderConfig(batchSize=16, height=256, width=256, scaling=2), GPU)


def test_device_name_precision_and_sidecar_rules(built_lib, tmp_path):
    import w2x
    d = str(tmp_path)
    model = os.path.join(d, "scale2x.onnx")
    open(model, "wb").write(b"onnx")
    cfg = w2x.BuildConfig.fixed(2, 64)
    rc = w2x.RenderConfig(batchSize=2, height=64, width=64, scaling=2)
    _artefact(d, "scale2x", cfg, device="NVIDIA H100")                       # built for another GPU model
    _artefact(d, "scale2x", w2x.BuildConfig.fixed(2, 128), sidecar=False)    # no sidecar: skipped (img2img_load.cpp:95-98)
    with pytest.raises(RuntimeError, match="could not satisfy"):
        w2x.select_engine(model, rc, GPU)
    good = _artefact(d, "scale2x", cfg)
    assert w2x.select_engine(model, rc, GPU) == good
    assert w2x.select_engine(model, w2x.RenderConfig(deviceId=5, batchSize=2, height=64, width=64, scaling=2), GPU) == good  # any device with that name
    tf32 = w2x.RenderConfig(precision=w2x.PRECISION_TF32, batchSize=2, height=64, width=64, scaling=2)
    with pytest.raises(RuntimeError, match="could not satisfy"):
        w2x.select_engine(model, tf32, GPU)                                  # precision must match (isCompatible, :9-18)


def test_name_must_be_stem_plus_hash(built_lib, tmp_path):
    """the reference's prefix match lets model `noise0` load `noise0_scale2x_<hash>`; here the name is exact."""
    import w2x
    d = str(tmp_path)
    model = os.path.join(d, "noise0_.onnx")                                   # scale-1 model path rule (main.cpp:201-204)
    open(model, "wb").write(b"onnx")
    cfg = w2x.BuildConfig.fixed(1, 64)
    _artefact(d, "noise0_scale2x", cfg)                                       # another model's engine sharing the prefix
    _artefact(d, "noise0_", cfg, ext=".trt")                                  # a TensorRT plan of the reference: ignored
    rc = w2x.RenderConfig(batchSize=1, height=64, width=64, scaling=1)
    with pytest.raises(RuntimeError, match="could not satisfy"):
        w2x.select_engine(model, rc, GPU)
    mine = _artefact(d, "noise0_", cfg)
    assert w2x.select_engine(model, rc, GPU) == mine
    with pytest.raises(RuntimeError, match="model file does not exist"):
        w2x.select_engine(os.path.join(d, "missing.onnx"), rc, GPU)
