mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/mma_stream.log
import sys, ctypes as C; sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
print("# UMMA (M=128, K=16, fp16, operands resident in smem) issue rate with a concurrent cp.async.bulk stream into another smem region")
print("# columns: N, stream copy size, SM cycles per MMA, streamed bytes per SM cycle, operand bytes per cycle (A 4 KB + B N*32 B per MMA)")
res = (C.c_float * 2)()
for n in (64, 128, 256):
    for sb in (0, 4096, 16384):
        rc = l.w2x_probe_mma_rate_stream(0, n, 20000, sb, res)
        op = (4096 + n * 32) / res[0] if res[0] > 0 else 0
        print(f"N={n:3d} stream={sb:5d} B  rc={rc}  {res[0]:6.1f} cycles/MMA  {res[1]:6.1f} B/clk streamed  {op:6.1f} B/clk operands")
PY
