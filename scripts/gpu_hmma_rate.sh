mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/hmma_rate.log
import sys; sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
iters = 20000
print("# legacy mma.sync.m16n8k16 (fp16 -> fp32) issue-rate probe: operands in registers, 148 SMs; clock assumed 1.965 GHz")
for warps in (1, 4, 8, 16):
    for chains in (1, 2, 4, 8):
        ms = l.w2x_probe_hmma_rate(0, warps, chains, iters)
        n = warps * chains * iters
        cyc_sm = ms * 1e-3 * 1.965e9 / n          # SM cycles per HMMA
        tf = 148 * n * 16 * 8 * 16 * 2 / (ms * 1e-3) / 1e12
        lat = ms * 1e-3 * 1.965e9 / iters / chains if warps == 1 and chains == 1 else None
        print(f"warps/SM={warps:2d} chains={chains}  {ms:8.3f} ms  {cyc_sm:6.2f} SM-cycles/HMMA  {tf:7.1f} TFLOP/s" + (f"  (dependent-issue latency {lat:.1f} cycles)" if lat else ""))
PY
