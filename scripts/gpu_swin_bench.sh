mkdir -p gpurun_out
W2X_VERBOSE=1 timeout 900 python bench.py --workload swin --steps 6 --warmup 2 --layers --no-cpu-baseline > gpurun_out/bench_swin.json 2> gpurun_out/bench_swin.err
tail -c 1300 gpurun_out/bench_swin.json
grep -v "^\[w2x\]" gpurun_out/bench_swin.err | awk '{s+=$2} END{print "sum ms/batch:", s}'
grep -v "^\[w2x\]" gpurun_out/bench_swin.err | head -24
