#!/bin/bash
# fused SwinUNet attention kernel: unit parity vs torchvision fp32, its standalone time
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_swin_attn.py -x -q -m gpu 2>&1 | tail -15
timeout -s KILL 120 python - <<'PY'
import sys
sys.path.insert(0, 'waifu2x-tensorrt_b200'); sys.path.insert(0, 'tests')
import w2x
from test_gpu_swin_attn import make_case
for c, n, h, w in [(96, 4, 240, 240), (192, 4, 120, 120), (192, 4, 60, 60)]:
    case = make_case(n, h, w, 3, c)
    for shift in (0, 3):
        out, ms = w2x.run_swin_attn(*case, shift=shift, reps=50)
        print('fused attention c=%d shift=%d, %d tokens: %.4f ms per launch' % (c, shift, n * h * w, ms), flush=True)
PY
