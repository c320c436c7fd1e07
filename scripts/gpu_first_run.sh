mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== tiling + model with direct convs ===" 
W2X_CONV_IMPL=direct timeout 600 python -m pytest tests/test_gpu_tiling.py tests/test_gpu_model.py -x -q -m gpu -k "not flops" 2>&1 | tail -25 | tee gpurun_out/t_direct.log
echo "=== conv selftests (tcgen05) ==="
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/t_conv.log
echo "=== model with igemm ==="
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/t_igemm.log
