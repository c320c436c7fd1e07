# N-GPU weak-scaling bench + multi-GPU tests (run under gpurun --gpus N); N from $1
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 16 --warmup 4 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 900 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
timeout 600 python -m pytest tests/test_gpu_banded.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/t_banded_${N}gpu.log
