#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (per kernel + sequence)."""
import collections
import csv
import sys


def main(path, show_seq=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        t = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        t = t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)
        seq.append((row["Kernel Name"], t, row["Grid Size"], row["Block Size"]))
    agg = collections.OrderedDict()
    for name, t, _, _ in seq:
        k = name.split("(")[0][-70:]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{len(seq)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} us {v[0]:4d}x {100 * v[1] / tot:5.1f}%  {k}")
    if show_seq:
        for i, (name, t, g, b) in enumerate(seq):
            print(f"{i:4d} {t:9.1f} us grid={g:>18} {name.split('(')[0][-50:]}")


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)
