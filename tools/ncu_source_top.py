#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` export: python tools/ncu_source_top.py file.csv [n] [lo hi]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"{len(body)} instructions, {tot} samples")
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 // max(tot, 1) >= 1))
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i, r in enumerate(body[lo:hi], lo):
        s = int(r[ix["# Samples"]] or 0)
        top = max(stalls, key=lambda k: int(r[ix[k]] or 0))
        print(f"{i:6d} {s:6d} {r[ix['Instructions Executed']]:>9} {top[6:]:12s} {r[ix['Source']].strip()[:90]}")
else:
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
    for i in sorted(order):
        r = body[i]
        s = int(r[ix["# Samples"]] or 0)
        top = max(stalls, key=lambda k: int(r[ix[k]] or 0))
        print(f"{i:6d} {s:6d} {s * 100.0 / tot:5.1f}% {r[ix['Instructions Executed']]:>9} {top[6:]:12s} {r[ix['Source']].strip()[:90]}")
