#!/bin/bash
# per-layer table of the swin workload + ncu --set full of the fused MLP kernel
mkdir -p gpurun_out
timeout 600 python bench.py --only --workload swin --no-cpu-baseline --steps 8 --layers > gpurun_out/bench_swin_layers.json 2> gpurun_out/layers_swin.txt
grep -c ms gpurun_out/layers_swin.txt
cat > /tmp/mlp_run.py <<'PY'
import sys
sys.path.insert(0, 'waifu2x-tensorrt_b200'); sys.path.insert(0, 'tests')
import w2x
from test_gpu_swin_mlp import make_case
out, ms = w2x.run_swin_mlp(*make_case(4 * 240 * 240, 3), reps=3)
PY
timeout 600 ncu --clock-control none --set full --import-source on -k regex:swin_mlp_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_swin_mlp -f python /tmp/mlp_run.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
