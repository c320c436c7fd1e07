#!/bin/bash
# compute-sanitizer memcheck over the fused token kernels (small geometries: ragged tiles, shifted blocks, both widths)
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_swin_attn.py tests/test_gpu_swin_mlp.py -q -m gpu -k "1-6-6 or 2-12-30 or 1-18-6 or independent or (129 and not 148)" 2>&1 | tail -15 > gpurun_out/r02_memcheck_swin.txt
echo "exit $?" >> gpurun_out/r02_memcheck_swin.txt
cat gpurun_out/r02_memcheck_swin.txt
# engine level: a swin render (fused attention at both widths, composed head, row-major tile order, band-wise stitch + download) and the
# cunet band-wise download test
timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_swin.py tests/test_gpu_model.py -q -m gpu -k "swin_render_matches_oracle or tile_112 or bandwise" 2>&1 | tail -8 > gpurun_out/r02_memcheck_engine.txt
echo "exit $?" >> gpurun_out/r02_memcheck_engine.txt
cat gpurun_out/r02_memcheck_engine.txt
