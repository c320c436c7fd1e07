#!/bin/bash
# swin model tests + swin bench line with the per-layer table
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_swin_attn.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 300 python bench.py --only --workload swin --no-cpu-baseline --steps 16 --layers > gpurun_out/bench_swin_attn.json 2> gpurun_out/layers_swin_attn.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_swin_attn.json') if l.startswith('{')][0]); print('swin', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'], d['roofline']['model_stage']['frac'], d['clocks'], d['gpu_launches'])"
grep -E "up0|to_image|up1|patch" gpurun_out/layers_swin_attn.txt
