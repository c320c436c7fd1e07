#!/bin/bash
# fused SwinUNet MLP kernel: unit parity vs torch fp32, its standalone time, the swin model tests, the swin bench line
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_swin_mlp.py -x -q -m gpu 2>&1 | tail -15
timeout -s KILL 90 python - <<'PY'
import sys, os
sys.path.insert(0, 'waifu2x-tensorrt_b200'); sys.path.insert(0, 'tests')
import numpy as np, w2x
from test_gpu_swin_mlp import make_case
for c, tokens, variant in [(96, 4 * 240 * 240, 0), (96, 4 * 240 * 240, 1), (192, 4 * 120 * 120, 0)]:
    case = make_case(tokens, 3, c)
    for _ in range(2):
        out, ms = w2x.run_swin_mlp(*case, reps=50, variant=variant)
        print('fused mlp c=%d variant=%d, %d tokens: %.4f ms per launch' % (c, variant, tokens, ms))
PY
timeout -s KILL 300 python -m pytest tests/test_gpu_swin.py tests/test_gpu_banded.py -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 200 python bench.py --only --workload swin --no-cpu-baseline --steps 16 2>/dev/null | tee gpurun_out/bench_swin_mlp.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('swin', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
