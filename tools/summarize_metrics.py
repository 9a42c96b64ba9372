#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
sm__pipe_tensor_cycles_active...` pass over one step, and profiles/igemm_traffic.json (DRAM bytes per
igemm launch: the `roofline.traffic` figure bench.py reports).
usage: summarize_metrics.py <metrics.csv> [--write-traffic profiles/igemm_traffic.json]"""
import collections
import csv
import json
import sys


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return num(v) * scale.get(unit, 1)


def to_us(v, unit):
    scale = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}
    return num(v) * scale.get(unit, 1)


def main(path, traffic_out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = launches.setdefault(row["ID"], {"name": row["Kernel Name"].split("(")[0][-60:]})
        m, v, u = row["Metric Name"], row["Metric Value"], row["Metric Unit"]
        if m.startswith("dram__bytes_read"):
            d["rd"] = to_bytes(v, u)
        elif m.startswith("dram__bytes_write"):
            d["wr"] = to_bytes(v, u)
        elif m.startswith("gpu__time_duration"):
            d["us"] = to_us(v, u)
        elif m.startswith("sm__pipe_tensor_cycles_active"):
            d["tc"] = num(v)
    agg = collections.OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], dict(n=0, us=0.0, rd=0.0, wr=0.0, tc_us=0.0))
        a["n"] += 1
        a["us"] += d.get("us", 0.0)
        a["rd"] += d.get("rd", 0.0)
        a["wr"] += d.get("wr", 0.0)
        a["tc_us"] += d.get("tc", 0.0) * d.get("us", 0.0) / 100.0
    tot = sum(a["us"] for a in agg.values())
    print(f"{len(launches)} launches, {tot:.1f} us total (cold-cache, serialised under ncu: compare shares)")
    print(f"{'us':>10} {'n':>4} {'share':>6} {'DRAM rd MB':>11} {'DRAM wr MB':>11} {'GB/s':>7} {'tensor%':>8}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0
        tc = 100 * a["tc_us"] / a["us"] if a["us"] else 0
        print(f"{a['us']:10.1f} {a['n']:4d} {100 * a['us'] / tot:5.1f}% {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:11.1f} {gbs:7.0f} {tc:8.1f}  {k}")
    if traffic_out:
        ig = [d for d in launches.values() if "igemm_conv_kernel" in d["name"]]
        total = sum(d.get("rd", 0) + d.get("wr", 0) for d in ig)
        json.dump({"bytes_per_launch": total / max(len(ig), 1), "launches": len(ig), "dram_bytes_total": total,
                   "source": f"bytes per igemm_conv_kernel launch: dram__bytes_read.sum + dram__bytes_write.sum over the {len(ig)} "
                             f"igemm launches of one step (ncu pass {path})"}, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--write-traffic") + 1] if "--write-traffic" in sys.argv else None
    main(sys.argv[1], out)
