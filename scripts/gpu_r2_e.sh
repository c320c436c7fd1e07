#!/bin/bash
mkdir -p gpurun_out
for b in 8 4 2 1; do
  timeout 300 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --batch $b > gpurun_out/r2e_b$b.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2e_b$b.json')); print('batch $b', round(d['value'],1), 'Mpx/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value'],1), d['stage_ms_last_frame'], d['clocks'])" | tee -a gpurun_out/r2e_batch.txt
done
