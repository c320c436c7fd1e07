python - <<'PY' 2>&1 | tee gpurun_out/mma_rate.log
import sys; sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
iters = 20000
print("# UMMA issue-rate probe: M=128, K=16 fp16, 148 SMs, 4*iters MMAs per SM; clock assumed 1.965 GHz (boost, short run)")
for sbo in (1024, 1280):
    for n in (16, 32, 64, 96, 128, 192, 256):
        ms = l.w2x_probe_mma_rate(0, n, iters, sbo)
        cyc = ms * 1e-3 * 1.965e9 / (4 * iters)
        tf = 148 * 4 * iters * 128 * n * 16 * 2 / (ms * 1e-3) / 1e12
        print(f"sboA={sbo:5d} N={n:4d}  {ms:8.3f} ms  {cyc:7.1f} cycles/MMA  {tf:8.1f} TFLOP/s")
PY
