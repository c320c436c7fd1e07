mkdir -p gpurun_out
echo "=== conv selftests ==="
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/t_conv.log
echo "=== model tests ==="
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/t_model.log
echo "=== bench ==="
W2X_VERBOSE=1 timeout 600 python bench.py --steps 16 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
tail -c 1500 gpurun_out/bench_v2.json; grep -v "^\[w2x\]" gpurun_out/bench_v2.err | tail -30; grep "^\[w2x\]" gpurun_out/bench_v2.err | sort -u | head -30
