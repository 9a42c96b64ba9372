#!/usr/bin/env python
"""Copy the output of tools/final_profile_r02.sh (gpurun_out/r02) into profiles/ and write the summaries."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out", "r02"), os.path.join(ROOT, "profiles")
for src, dst in [("bench.json", "r02_bench_final.json"), ("bench_partition_off.json", "r02_bench_final_partition_off.json"),
                 ("bench_reference_arm.json", "r02_bench_final_reference_arm.json"), ("timeline.txt", "r02_timeline_final.txt"),
                 ("microbench.jsonl", "r02_microbench.jsonl"), ("kernel_sweep.jsonl", "r02_kernel_sweep.jsonl"),
                 ("launches_step.csv", "r02_launches_step.csv"), ("step_metrics.csv", "r02_step_metrics.csv"),
                 ("prof_enc.txt", "r02_prof_enc.txt"), ("prof_up.txt", "r02_prof_up.txt")]:
    shutil.copy(os.path.join(O, src), os.path.join(P, dst))


def run(args, out):
    open(os.path.join(P, out), "w").write(subprocess.run([sys.executable] + args, capture_output=True, text=True, cwd=ROOT).stdout)


run(["tools/summarize_launches.py", os.path.join(O, "launches_step.csv")], "r02_launches_step_summary.txt")
run(["tools/summarize_metrics.py", os.path.join(P, "r02_step_metrics.csv"), "--write-traffic", os.path.join(P, "igemm_traffic.json")],
    "r02_step_metrics_summary.txt")
run(["tools/summarize_sweep.py", os.path.join(O, "kernel_sweep.jsonl")], "r02_kernel_sweep_summary.txt")
tj = os.path.join(P, "igemm_traffic.json")
t = json.load(open(tj))
t["source"] = t["source"].replace(P + os.sep, "profiles/").replace(ROOT + os.sep, "")
json.dump(t, open(tj, "w"), indent=1)
keys = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size", "launch__cluster_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for name in ["wgrad64", "igemm64_pair", "up64", "lat2", "blur256"]:
    rows = list(csv.reader(open(os.path.join(O, f"{name}_raw.csv"))))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    with open(os.path.join(P, f"r02_{name}_full_summary.txt"), "w") as f:
        for k in keys:
            if k in d:
                f.write(f"{k} = {d[k][0]} {d[k][1]}\n")
        f.write("\n--- source-level stall sampling (ncu --page source --csv; tools/ncu_source_top.py): top sites ---\n")
        f.write(subprocess.run([sys.executable, "tools/ncu_source_top.py", os.path.join(O, f"{name}_source.csv"), "25"],
                               capture_output=True, text=True, cwd=ROOT).stdout)
    print(name, d["gpu__time_duration.sum"][0], "us  tensor", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", ("", ""))[0],
          " dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("", ""))[0])
for f in ["bench.json", "bench_partition_off.json", "bench_reference_arm.json"]:
    d = json.loads(open(os.path.join(O, f)).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d.get("clocks", {}).get("sm_mhz"), "roofline", r.get("achieved"), r.get("frac"),
          "unweighted", r.get("frac_unweighted"), "full_chip", (r.get("full_chip") or {}).get("frac"))
    for k, v in (r.get("by_network") or {}).items():
        print("   ", k, round(v["achieved"]), round(v["frac"], 3), round(v["sm_share"], 3), round(v["kernel_ms_per_step"], 2), round(v["tflops_on_its_sms"]))
    if "hbm_kernels" in d:
        print("   ", {k: (round(v["achieved"]), round(v["frac"], 3)) for k, v in d["hbm_kernels"].items() if isinstance(v, dict)})
