mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/t_swin.log
