mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/l2_stream.log
import sys; sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
print("# L2 -> SM stream probe: 148 SMs each copy the same L2-resident buffer into smem with cp.async.bulk, 4 copies in flight per SM; 1.965 GHz assumed")
for kb in (4, 8, 16, 32, 48):
    iters = 200000 // kb
    ms = l.w2x_probe_l2_stream(0, kb * 1024, iters)
    tot = 148 * iters * kb * 1024
    print(f"copy={kb:2d} KiB  {ms:8.3f} ms  {tot / (ms * 1e-3) / 1e12:6.2f} TB/s aggregate  {iters * kb * 1024 / (ms * 1e-3 * 1.965e9):6.1f} B/clk/SM")
PY
