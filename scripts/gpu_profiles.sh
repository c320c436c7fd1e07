# ncu evidence for profiles/ (run under gpurun, one GPU).  Numbers printed by these runs are never bench values.
mkdir -p gpurun_out
# 1. launch lists (device time per launch, cold-cache + serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 320 --csv --log-file gpurun_out/launches_cunet.csv \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/launches_cunet.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/launches_swin.csv \
    python bench.py --workload swin --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/launches_swin.log 2>&1
# 2. full capture of the dominant kernel family: one batch worth of conv3x3_patch launches (11 per batch)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_patch -s 44 -c 11 -o gpurun_out/prof_patch_batch \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_patch_batch.log 2>&1
# 3. memory-bound tiling kernels
timeout 600 ncu --set full --clock-control none -k regex:"unpack_kernel|stitch_kernel" -s 10 -c 3 -o gpurun_out/prof_tiling \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_tiling.log 2>&1
# 4. image head kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_head|conv_up4" -s 18 -c 2 -o gpurun_out/prof_head \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_head.log 2>&1
ls -la gpurun_out
