mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/mma_tiles.log
import sys; sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
print("# patch-kernel MMA schedule in isolation: 36 UMMAs (M=128, N=64, K=16) per tile, alternating accumulators; SM cycles per MMA (48.0 = operand-stream bound)")
names = {1: "wait tile t-2", 2: "same B for all taps", 4: "unshifted A", 8: "single commit", 16: "one accumulator", 32: "always accumulate", 64: "rolled tap loops", 128: "ky rolled, kx/ks unrolled", 256: "two issuing warps"}
for mode in (0, 129, 256, 257, 384, 385):
    c = l.w2x_probe_mma_tiles(0, 2000, mode)
    desc = ", ".join(v for k, v in names.items() if mode & k) or "plain"
    print(f"mode={mode:2d} ({desc:50s}) {c:6.1f} cycles/MMA")
PY
