#!/bin/bash
# end-of-round check on one GPU: the whole GPU suite, smoke(), and the driver's bench line with the per-layer tables
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 900 python bench.py --layers > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_layers_all.txt
head -c 300 gpurun_out/r02_bench_full.json; echo
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 | head -c 600; echo
