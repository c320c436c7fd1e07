"""The C-ABI library loads without a GPU and exports every symbol include/w2x.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "w2x.h")).read()
    return sorted(set(re.findall(r"W2X_API[^;]*?\b(w2x_[a-z0-9_]+)\s*\(", hdr, re.S)))


def test_header_and_binding_agree(built_lib):
    import w2x
    decl = declared_symbols()
    assert len(decl) >= 30
    assert decl == w2x.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol(built_lib):
    import w2x
    raw = C.CDLL(w2x.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), name


def test_defaults_match_reference_config(built_lib):
    """src/tensorrt/config.h:12-43 defaults."""
    import w2x
    b = w2x._BuildConfig()
    built_lib.w2x_default_build_config(C.byref(b))
    assert [getattr(b, f[0]) for f in b._fields_] == [0, 1, 1, 1, 4, 3, 3, 3, 64, 256, 640, 64, 256, 640]
    r = w2x._RenderConfig()
    built_lib.w2x_default_render_config(C.byref(r))
    assert (r.deviceId, r.precision, r.batchSize, r.channels, r.height, r.width, r.scaling, r.overlapX, r.overlapY, r.tta) == \
        (0, 1, 1, 3, 256, 256, 4, 0.0625, 0.0625, 0)
    d = w2x.BuildConfig()
    assert [getattr(d, f[0]) for f in b._fields_] == [getattr(b, f[0]) for f in b._fields_]


def test_errors_do_not_cross_the_boundary(built_lib):
    """Like the reference's function-try-blocks: failure == false + a message, never an exception/abort."""
    import w2x
    e = w2x.Img2Img()
    msgs = []
    e.setMessageCallback(lambda sev, m: msgs.append((sev, m)))
    assert e.load("/nonexistent/model.onnx", w2x.RenderConfig()) is False
    assert msgs and msgs[-1][0] == 1 and "Failed to find engine file" in msgs[-1][1] and "model file does not exist" in msgs[-1][1]
    assert msgs[-1][1].startswith("[load@")  # "[func@line] msg", logger.cpp:19-22
    assert e.render(__import__("numpy").zeros((8, 8, 3), "uint8")) is None
    assert "no engine loaded" in e.last_error
    e.close()
