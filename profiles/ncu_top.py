#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/ncu_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][:2])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
body = rows[2:]
total_samples = 0
total_inst = 0
for r in body:
    if len(r) < len(hdr):
        continue
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
    total_samples += int(r[col["# Samples"]] or 0)
    total_inst += int(r[col["Instructions Executed"]] or 0)
print("samples", total_samples, "warp-instructions", total_inst)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {s:28s} {v:9d} {100.0 * v / max(total_samples, 1):5.1f}%")
print("hottest SASS lines (samples, executed, top stall):")
hot = sorted([r for r in body if len(r) >= len(hdr)], key=lambda r: -int(r[col["# Samples"]] or 0))[:topn]
for r in hot:
    st = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f"  {int(r[col['# Samples']]):7d} {int(r[col['Instructions Executed']]):11d}  {st:18s} {r[col['Source']].strip()[:90]}")
