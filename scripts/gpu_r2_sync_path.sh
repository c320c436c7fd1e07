#!/bin/bash
# progressive stitch + band download in the synchronous drop-in render(): parity tests, then e2e_sync of both headline workloads
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_swin.py tests/test_gpu_tiling.py tests/test_gpu_banded.py -x -q -m gpu 2>&1 | tail -5
for wl in cunet swin; do
timeout -s KILL 300 python bench.py --only --workload $wl --no-cpu-baseline --steps 16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_sync', round(d['e2e_sync']['value'],1), round(d['e2e_sync']['ms_per_step'],2), 'ms', d['stage_ms_last_frame'])"
done
