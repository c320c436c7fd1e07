#!/bin/bash
# round 2 evidence run (one GPU): [GPU test suite,] the full bench line, per-layer tables, ncu launch lists and --set full captures.
# gpurun copies back at most 64 MiB: the .ncu-rep files are summarised on the box and only the small ones are kept.
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt; fi
timeout -s KILL 900 python bench.py --layers > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_layers_all.txt
head -c 300 gpurun_out/r02_bench_full.json; echo
NCU="ncu --clock-control none"
B="python bench.py --only --steps 2 --warmup 3 --no-cpu-baseline"
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches_cunet.csv $B > /dev/null 2>&1
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches_swin.csv $B --workload swin > /dev/null 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:conv3x3_patch_kernel -s 33 -c 11 -o gpurun_out/r02_ncu_patch_batch -f $B > /dev/null 2>&1
timeout -s KILL 600 $NCU --set full -k regex:"swin_attn_kernel|swin_mlp" -s 32 -c 32 -o gpurun_out/r02_ncu_swin_fused -f $B --workload swin > /dev/null 2>&1
for f in r02_ncu_patch_batch r02_ncu_swin_fused; do python scripts/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.txt 2>&1; done
find gpurun_out -name '*.ncu-rep' -size +30M -delete
ls -la gpurun_out
