"""Development: steady-state SM clock and board power while one kernel family runs back to back for a few seconds
(nvidia-smi sampled every 20 ms), for the conv3x3 patch kernel at the conv5 shape, the UMMA issue-rate probe and a cuBLAS bf16 GEMM
(the workload MEASURED_PEAKS.json's tensor peak comes from).  Shows whether a kernel's clock is set by the power cap."""
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
import torch  # noqa: E402
import w2x  # noqa: E402

Q = "clocks.sm,power.draw.instant,power.draw.average,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu,enforced.power.limit"


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        sel = [r for t, r in self.rows if t0 <= t <= t1]
        def col(i):
            v = []
            for r in sel:
                try:
                    v.append(float(r[i]))
                except Exception:
                    pass
            return v
        clk, pin, pav = col(0), col(1), col(2)
        cap = sum(1 for r in sel if r[3].lower().startswith("active"))
        med = lambda v: sorted(v)[len(v) // 2] if v else None
        return dict(samples=len(sel), sm_mhz_median=med(clk), sm_mhz_min=min(clk) if clk else None, power_instant_median=med(pin), power_instant_max=max(pin) if pin else None,
                    power_avg_median=med(pav), sw_power_cap_samples=cap, temp=med(col(6)), limit=med(col(7)))


def run_for(seconds, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        fn()
        n += 1
        if n % 8 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return t0, time.perf_counter(), n


s = Sampler()
time.sleep(0.5)
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    torch.matmul(a, b)
t0, t1, n = run_for(3.0, lambda: torch.matmul(a, b))
r = s.window(t0 + 0.5, t1)
print(f"cuBLAS bf16 8192^3: {2 * 8192 ** 3 * n / (t1 - t0) / 1e12:7.1f} TFLOP/s  {r}", flush=True)
time.sleep(1.0)

w2x.use_dev_lib()
l = w2x.dev_lib()
for nn in (64, 128, 256):
    t0, t1, n = run_for(2.0, lambda: l.w2x_probe_mma_rate(0, nn, 20000, 1024))
    r = s.window(t0 + 0.5, t1)
    print(f"UMMA issue probe N={nn}: {r}", flush=True)
    time.sleep(0.5)

# conv5 shape through the product kernel, back to back
rng = np.random.default_rng(0)
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402
import tempfile
tmp = tempfile.mkdtemp()
_, onnx = __graft_entry__.make_synthetic_model(tmp, scale=2, noise=3, model="cunet/art")
eng = w2x.Img2Img()
assert eng.build(onnx, w2x.BuildConfig.fixed(8, 256)) and eng.load(onnx, w2x.RenderConfig(batchSize=8, height=256, width=256, scaling=2))
t0 = time.perf_counter()
prof = None
while time.perf_counter() - t0 < 4.0:
    prof = eng.profile_layers(20)
t1 = time.perf_counter()
r = s.window(t0 + 0.5, t1)
print(f"cunet layers back to back (profile_layers x20 per layer): {r}")
for name, ms, fl in prof:
    if ms > 0.02:
        print(f"   {name:28s} {ms:7.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s")
