"""Development: board power / SM clock of conv3x3_patch_kernel at the unet2.conv5 shape with parts of the kernel switched off
(lib/libw2x_dev.so, W2X_DBG) -- a power decomposition of the layer.  Each variant runs ~3 s of back-to-back launches."""
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
import w2x  # noqa: E402

w2x.use_dev_lib()
rows = []
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw.instant", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)


def pump():
    for line in p.stdout:
        try:
            c, w = [float(x) for x in line.split(",")]
            rows.append((time.perf_counter(), c, w))
        except Exception:
            pass


threading.Thread(target=pump, daemon=True).start()
n, h, w_, cin, cout = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (8, 444, 444, 64, 64))]
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, size=(n, h, w_, cin)).astype(np.float16)
wp = (rng.uniform(-1, 1, size=(cout, 9 * cin)) / np.sqrt(9 * cin)).astype(np.float16)
b = np.zeros(cout, np.float32)
w2x.run_conv_layer(0, x, wp, b, cout)
reps = int(os.environ.get("W2X_REPEAT", "1"))
t0 = time.perf_counter()
w2x.run_conv_layer(0, x, wp, b, cout)
t1 = time.perf_counter()
sel = [(c, pw) for t, c, pw in rows if t0 + 1.0 <= t <= t1 - 0.3]
med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")
flops = 2.0 * n * (h - 2) * (w_ - 2) * cout * 9 * cin * reps
print(f"W2X_DBG={os.environ.get('W2X_DBG', '0'):>3s} reps {reps}: wall {t1 - t0:5.2f} s (incl. copies)  ~{flops / (t1 - t0 - 0.35) / 1e12:6.0f} TFLOP/s  sm clock median {med([c for c, _ in sel]):6.0f} MHz  power median {med([pw for _, pw in sel]):6.0f} W  ({len(sel)} samples)")
p.terminate()
