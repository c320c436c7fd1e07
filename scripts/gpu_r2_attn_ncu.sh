#!/bin/bash
# ncu --set full of the fused attention kernel alone (level-1 geometry of a batch of four 256-pixel tiles, shifted block)
mkdir -p gpurun_out
cat > /tmp/attn_run.py <<'PY'
import sys
sys.path.insert(0, 'waifu2x-tensorrt_b200'); sys.path.insert(0, 'tests')
import w2x
from test_gpu_swin_attn import make_case
out, ms = w2x.run_swin_attn(*make_case(4, 240, 240, 3), shift=3, reps=3)
PY
timeout -s KILL 600 ncu --clock-control none --set full --import-source on -k regex:swin_attn_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_swin_attn_v2 -f python /tmp/attn_run.py > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_swin_attn_v2.ncu-rep
