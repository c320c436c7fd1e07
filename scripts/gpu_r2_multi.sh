#!/bin/bash
# 8-GPU box: the single-process engine pool on an ordered frame stream (BASELINE configs[4]) and row-band strong scaling
mkdir -p gpurun_out
timeout 600 python scripts/bench_stream_pool.py 8 32 > gpurun_out/stream_pool_8gpu.txt 2>&1
tail -12 gpurun_out/stream_pool_8gpu.txt
timeout 600 python scripts/bench_banded.py 8 > gpurun_out/banded_strong_scaling.txt 2>&1
tail -14 gpurun_out/banded_strong_scaling.txt
