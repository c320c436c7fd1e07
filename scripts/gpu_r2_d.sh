#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3 16 19 23; do
  W2X_DBG=$dbg W2X_REPEAT=25000 timeout 120 python scripts/r2_power_dbg.py 2>&1 | grep W2X_DBG | tee -a gpurun_out/r2d_power_dbg.txt
done
