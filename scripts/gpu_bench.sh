# usage: bash scripts/gpu_bench.sh [tag]   (run under gpurun; writes gpurun_out/*)
TAG=${1:-r01}
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 900 python bench.py --steps 16 --warmup 3 --layers > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -40 gpurun_out/bench_$TAG.err
