#!/bin/bash
# round 2 evidence run (one GPU): GPU test suite, the full bench line, per-layer tables, ncu launch lists and --set full captures
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt
timeout -s KILL 900 python bench.py --layers > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_layers_all.txt
head -c 400 gpurun_out/r02_bench_full.json; echo
NCU="ncu --clock-control none"
B="python bench.py --only --steps 2 --warmup 3 --no-cpu-baseline"
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/r02_launches_cunet.csv $B > /dev/null 2>&1
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file gpurun_out/r02_launches_swin.csv $B --workload swin > /dev/null 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:conv3x3_patch_kernel -s 33 -c 11 -o gpurun_out/r02_ncu_patch_batch -f $B > /dev/null 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:"swin_attn_kernel|swin_mlp" -s 32 -c 32 -o gpurun_out/r02_ncu_swin_fused -f $B --workload swin > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
