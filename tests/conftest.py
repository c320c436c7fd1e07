import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "waifu2x-tensorrt_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libw2x.so, built in-tree.  Built on demand so `pytest -m "not gpu"` works on a fresh checkout."""
    import w2x
    if not os.path.exists(w2x.LIB_PATH) or not os.path.exists(os.path.join(PKG, "bin", "waifu2x-b200")):
        import __graft_entry__
        __graft_entry__.build()
    return w2x.lib()


@pytest.fixture(scope="session")
def oracle_c():
    """The plain-C restatement of the tile grid (oracle/tiling_c.c), via ctypes."""
    import ctypes as C
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_tiling.so")
    if not os.path.exists(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "oracle", "tiling_c.c"), "-o", so, "-lm"])
    return C.CDLL(so)


@pytest.fixture(scope="session")
def models_dir(tmp_path_factory):
    """Synthetic cunet/art ONNX files (scale 1 and 2) laid out like the reference's models/ directory."""
    import __graft_entry__
    d = str(tmp_path_factory.mktemp("models"))
    out = {}
    for scale in (1, 2):
        m, path = __graft_entry__.make_synthetic_model(d, scale=scale, noise=0)
        out[scale] = (m, path)
    return d, out
