#!/usr/bin/env python
"""Summarise an ncu `--page source --csv` dump: hottest SASS lines by stall samples / executed instructions."""
import csv
import sys

path = sys.argv[1]
top = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
        continue
    try:
        data.append((int(r[idx["# Samples"]] or 0), int(r[idx["Instructions Executed"]] or 0),
                     float(r[idx["Avg. Threads Executed"]] or 0), r[idx["Address"]][-4:], r[idx["Source"]].strip(),
                     {k[6:]: int(r[idx[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and int(r[idx[k]] or 0)}))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
toti = sum(d[1] for d in data)
print("total samples", tot, "total warp-inst", toti, "lines", len(data))
for i, d in enumerate(data):
    if d[0] >= tot * top or d[1] >= toti * top * 1.5:
        st = ",".join(f"{k}:{v}" for k, v in sorted(d[5].items(), key=lambda kv: -kv[1])[:3])
        print(f"{i:4d} {d[3]} smp={d[0]:6d} ({100.0*d[0]/max(tot,1):4.1f}%) inst={d[1]:9d} thr={d[2]:4.1f}  {d[4][:70]:70s} {st}")
