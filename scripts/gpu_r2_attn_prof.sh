#!/bin/bash
# fused attention in the engine: swin model tests, swin bench line + per-layer table, ncu --set full of the kernel
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_swin.py tests/test_gpu_swin_attn.py -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 300 python bench.py --only --workload swin --no-cpu-baseline --steps 16 --layers > gpurun_out/bench_swin_attn.json 2> gpurun_out/layers_swin_attn.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_swin_attn.json') if l.startswith('{')][0]); print('swin', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
cat > /tmp/attn_run.py <<'PY'
import sys
sys.path.insert(0, 'waifu2x-tensorrt_b200'); sys.path.insert(0, 'tests')
import w2x
from test_gpu_swin_attn import make_case
out, ms = w2x.run_swin_attn(*make_case(4, 240, 240, 3), shift=3, reps=3)
PY
timeout -s KILL 600 ncu --clock-control none --set full --import-source on -k regex:swin_attn_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_swin_attn -f python /tmp/attn_run.py > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_swin_attn.ncu-rep
