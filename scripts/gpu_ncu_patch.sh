mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/t_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_patch -c 1 -o gpurun_out/prof_patch_conv5 \
  python -c "
import sys; sys.path.insert(0,'waifu2x-tensorrt_b200'); import w2x; print(w2x.selftest_conv(0,8,444,444,64,64))" > gpurun_out/ncu_patch.log 2>&1
tail -3 gpurun_out/ncu_patch.log
