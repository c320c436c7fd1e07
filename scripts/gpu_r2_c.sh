#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r2_power_probe.py 2>&1 | grep -v Warning | tee gpurun_out/r2c_power.txt
