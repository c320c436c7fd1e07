"""Development: role counters of the fused first-layer kernel (lib/libw2x_dev.so, W2X_PROF=1) inside a real model run."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
import w2x
w2x.use_dev_lib()
import __graft_entry__
tmp = tempfile.mkdtemp()
_, onnx = __graft_entry__.make_synthetic_model(tmp, scale=2, noise=3, model="cunet/art")
eng = w2x.Img2Img()
assert eng.build(onnx, w2x.BuildConfig.fixed(8, 256)) and eng.load(onnx, w2x.RenderConfig(batchSize=8, height=256, width=256, scaling=2))
eng.profile_layers(1)
