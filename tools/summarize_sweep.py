#!/usr/bin/env python
"""Pivot tools/kernel_sweep.py output (JSON lines) into per-kernel tables: rows = resolution x channels, columns = batch."""
import collections
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
tables = collections.OrderedDict()
for r in rows:
    unit = "TFLOPs" if "TFLOPs" in r else "GBs"
    name = r["kernel"] + (f" {r['dtype']}" if "dtype" in r else "") + f"  [{'TFLOP/s' if unit == 'TFLOPs' else 'GB/s'}]"
    res = r.get("res", r.get("res_in", r.get("res_out")))
    ch = f"{r['Cin']}->{r['Cout']}" if "Cin" in r else str(r["C"])
    lab = {"res_in": "in "}.get(next((k for k in ("res_in",) if k in r), ""), "") + f"{res}^2 x {ch}"
    tables.setdefault(name, collections.OrderedDict()).setdefault(lab, {})[r["B"]] = (r[unit], r["us"])
for name, t in tables.items():
    bs = sorted({b for v in t.values() for b in v})
    print(name)
    print(f"  {'shape':<24}" + "".join(f"{'B=' + str(b):>10}" for b in bs) + f"{'us @B=' + str(bs[-1]):>14}")
    for lab, v in t.items():
        print(f"  {lab:<24}" + "".join(f"{v[b][0]:10.0f}" if b in v else f"{'-':>10}" for b in bs) + f"{v[bs[-1]][1]:14.1f}" if bs[-1] in v else "")
    print()
