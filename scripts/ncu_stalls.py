#!/usr/bin/env python
"""Top stall sites of one kernel in a .ncu-rep (source page, SASS view).  Usage: ncu_stalls.py file.ncu-rep <launch id> [top N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[hdr.index("# Samples")].isdigit():
        break  # a second view (with its own header) follows the first
    data.append(r)
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for c in stallcols:
        agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stallcols), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):6d} {100.0 * int(r[isamp]) / max(tot, 1):5.1f}% ex={r[iex]:>8s}  {r[isrc].strip()[:64]:64s} {st}")
