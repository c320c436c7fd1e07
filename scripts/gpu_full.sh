mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json
timeout 900 python bench.py --workload swin --steps 8 --warmup 2 > gpurun_out/bench_swin.json 2> gpurun_out/bench_swin.err; tail -c 600 gpurun_out/bench_swin.json
