#!/bin/bash
# round 2: tiling kernel rewrite check + the new bench line (all three workloads)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiling.py tests/test_gpu_model.py tests/test_gpu_swin.py tests/test_cli_videoio.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2f_tests.txt
timeout 900 python bench.py --layers > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_layers.txt
tail -c 600 gpurun_out/r2f_layers.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
def show(name, x):
    rt = x.get('roofline_tiling') or {}
    print(name, 'value', round(x['value'],1), 'ms', round(x['ms_per_step'],3), 'e2e', round(x['e2e']['value'],1), 'e2e_sync', round(x['e2e_sync']['value'],1),
          'roofline', round(x['roofline']['achieved'],1), round(x['roofline']['frac'],3), 'model_stage', round(x['roofline']['model_stage']['achieved'],1),
          'stage', {k: round(v,3) for k,v in x['stage_ms_last_frame'].items()},
          'tiling', {k: (round(v['achieved']), round(v['frac'],3)) for k,v in rt.items() if isinstance(v, dict)}, x['clocks'])
show('cunet', d)
for k,v in d.get('workloads',{}).items(): show(k, v)
print(d.get('cpu_baseline'))
PY
