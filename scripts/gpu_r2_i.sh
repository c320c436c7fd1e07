#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_banded.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2i_tests.txt
timeout 600 python bench.py --only --layers --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_layers.txt
grep -E "conv1.conv.2|conv5 |conv3.conv.2" gpurun_out/r2i_layers.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['stage_ms_last_frame'], d['clocks'])
PY
