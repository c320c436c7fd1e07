mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/probe.log
import sys, ctypes as C
sys.path.insert(0, "waifu2x-tensorrt_b200")
import w2x
l = w2x.dev_lib()  # probes live in lib/libw2x_dev.so (-DW2X_DEV)
for pitch in (16, 10, 24):
    for mode in (0, 1):
        err = (C.c_float * 9)()
        rc = l.w2x_probe_umma(0, mode, pitch, err)
        print(f"pitch {pitch:2d} mode {mode} rc {rc} err/tap:", " ".join(f"{e:8.4f}" for e in err), flush=True)
PY
# ncu full capture of the v1 igemm kernel on a conv5-like shape (8 x 444x444x64 -> 64)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -o gpurun_out/prof_v1_conv5 \
  python -c "
import sys; sys.path.insert(0,'waifu2x-tensorrt_b200'); import w2x; print(w2x.selftest_conv(0,8,444,444,64,64))" > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_v1.log
