#!/bin/bash
# round 2: validate a patch-kernel change (layer tests + model parity), role counters, per-layer table + bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_banded.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2b_tests.txt
W2X_PROF=1 timeout 120 python scripts/r2_prof_patch.py prof 2>&1 | grep "w2x prof" | tee gpurun_out/r2b_prof.txt
W2X_VERBOSE=1 timeout 600 python bench.py --steps 24 --warmup 4 --layers --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_layers.txt
cat gpurun_out/r2b_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['stage_ms_last_frame'], d['clocks'])"
grep -v "^\[w2x\]" gpurun_out/r2b_layers.txt | tail -30
