#!/bin/bash
# full GPU suite + the three-workload bench line (evidence refresh after a hot-path change)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests.txt
timeout 900 python bench.py --steps 32 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_full.json'))
def row(n, x):
    print(n, round(x['value'], 1), round(x['ms_per_step'], 3), 'e2e', round(x['e2e']['value'], 1), 'frac', round(x['roofline']['frac'], 3), x.get('roofline_tiling', {}).get('unpack'), x.get('clocks'))
row('cunet', d)
for k, v in d.get('workloads', {}).items():
    row(k, v)
print(d.get('cpu_baseline'))
PY
