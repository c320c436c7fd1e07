#!/bin/bash
# round 2, first GPU session: full GPU test suite + patch-kernel cycle counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.txt
cat gpurun_out/r2a_tests.txt
timeout 120 python scripts/r2_prof_patch.py probe 2>&1 | tee gpurun_out/r2a_probe.txt
for dbg in 0 1 16 19 3; do
  echo "== W2X_DBG=$dbg" | tee -a gpurun_out/r2a_prof.txt
  W2X_DBG=$dbg W2X_PROF=1 timeout 120 python scripts/r2_prof_patch.py prof 2>&1 | grep "w2x prof" | tee -a gpurun_out/r2a_prof.txt
done
