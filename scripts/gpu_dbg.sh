mkdir -p gpurun_out
for d in 0 1 2 3 4 5 7; do
  echo "=== W2X_DBG=$d ==="
  W2X_DBG=$d timeout 300 python bench.py --steps 4 --warmup 2 --layers --no-cpu-baseline 2>&1 >/dev/null | grep "conv1.conv.2\|conv2.conv.2\|conv3.conv.0\|conv4.conv.0\|conv5 \|unet1.conv3"
done
