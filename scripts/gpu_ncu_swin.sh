mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 1900 -c 6 -o gpurun_out/prof_swin_lin \
  python bench.py --workload swin --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_swin_lin.log 2>&1
tail -2 gpurun_out/prof_swin_lin.log | cut -c1-200
