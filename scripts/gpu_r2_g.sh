#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"unpack_kernel|stitch_kernel|tta_reduce_kernel" -s 8 -c 5 -o gpurun_out/r2g_tiling -f \
  python bench.py --workload cunet_tta --only --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
tail -3 gpurun_out/r2g_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stitch_kernel|unpack_kernel" -s 16 -c 3 -o gpurun_out/r2g_tiling_cunet -f \
  python bench.py --workload cunet --only --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
