mkdir -p gpurun_out
# launch list of one bench run (warm-up launches skipped): per-launch device time, cold-cache and serialised -> compare SHARES
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0][-60:]
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={v[0]:4d} total={v[1]/1e3:9.1f} us share={100*v[1]/tot:5.1f}%  avg={v[1]/v[0]/1e3:8.2f} us")
PY
