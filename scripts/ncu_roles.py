#!/usr/bin/env python
"""Warp-stall samples of one kernel in a .ncu-rep, split at given SASS instruction indices (the role branches of a warp-specialised
kernel).  Usage: ncu_roles.py file.ncu-rep [idx ...]   (no indices: print every instruction with >= 0.4 % of the samples)"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
cuts = [int(x) for x in sys.argv[2:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data), "executed warp instructions", sum(int(r[iex]) for r in data))
if cuts:
    edges = [0] + cuts + [len(data)]
    for lo, hi in zip(edges, edges[1:]):
        agg = collections.Counter()
        for r in data[lo:hi]:
            for c in stallcols:
                agg[hdr[c].replace("stall_", "")] += int(r[c])
        print(f"[{lo:5d},{hi:5d}) samples {sum(int(r[isamp]) for r in data[lo:hi]):6d}  executed {sum(int(r[iex]) for r in data[lo:hi]):10d}  {agg.most_common(6)}")
else:
    for i, r in enumerate(data):
        if int(r[isamp]) * 250 >= tot:
            st = sorted(((int(r[c]), hdr[c].replace("stall_", "")) for c in stallcols), reverse=True)[:2]
            print(f"{i:5d} {int(r[isamp]):6d} {100.0 * int(r[isamp]) / tot:5.1f}% ex={r[iex]:>9s}  {r[isrc].strip()[:64]:64s} {st}")
