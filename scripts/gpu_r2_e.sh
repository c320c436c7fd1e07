#!/bin/bash
# quick check after a hot-path change: the model/tiling/banded parity tests, then two bench lines of the default workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_tiling.py tests/test_gpu_banded.py tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -5
for i in 1 2; do
  timeout 300 python bench.py --only --no-cpu-baseline --steps 48 2>/dev/null | tee gpurun_out/bench_e_$i.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), d['clocks'], d['gpu_launches'])"
done
