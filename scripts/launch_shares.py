#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel time shares."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("w2x::", "").replace("<unnamed>::", "")[-58:]
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)-1} launches, {tot/1e6:.3f} ms total device time (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={v[0]:4d} total={v[1]/1e3:9.1f} us share={100*v[1]/tot:5.1f}%  avg={v[1]/v[0]/1e3:8.2f} us")
