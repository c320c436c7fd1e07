mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/mma_shift.log
import sys, ctypes as C; sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
print("# UMMA (M=128, N=64, K=16) issue rate when the A descriptor starts `rows` 128-byte rows into the swizzled tile (SBO = 1280 B, the patch kernel's tap views)")
res = (C.c_float * 2)()
for rows in (0, 1, 2, 3, 4, 7, 8, 10, 11, 12, 20, 21, 22):
    rc = l.w2x_probe_mma_rate_stream(0, 64, 20000, -(rows + 1), res)
    print(f"rows={rows:2d}  rc={rc}  {res[0]:6.1f} cycles/MMA")
PY
