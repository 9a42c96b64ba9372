#!/usr/bin/env python
"""Copy the output of tools/final_profile_r02b.sh (gpurun_out/r02b) into profiles/ and write the summaries."""
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out", "r02b"), os.path.join(ROOT, "profiles")
for src, dst in [("bench.json", "r02b_bench.json"), ("bench_reference_arm.json", "r02b_bench_reference_arm.json"),
                 ("microbench_ops.jsonl", "r02b_microbench_ops.jsonl"), ("kernel_sweep.jsonl", "r02b_kernel_sweep.jsonl"),
                 ("upfirdn_widths.txt", "r02b_upfirdn_widths.txt")]:
    shutil.copy(os.path.join(O, src), os.path.join(P, dst))
open(os.path.join(P, "r02b_pytest_gpu.txt"), "w").write(
    "".join(open(os.path.join(O, "pytest_gpu.txt")).readlines()[-1:]) + open(os.path.join(O, "smoke.txt")).read())
out = subprocess.run([sys.executable, "tools/summarize_sweep.py", os.path.join(O, "kernel_sweep.jsonl")], capture_output=True,
                     text=True, cwd=ROOT).stdout
open(os.path.join(P, "r02b_kernel_sweep_summary.txt"), "w").write(out)
keys = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for name in ["ufs_f32_65", "ufs_f32_257", "ufs_bf16_257"]:
    rows = list(csv.reader(open(os.path.join(O, name + "_raw.csv"))))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    with open(os.path.join(P, f"r02b_{name}_full_summary.txt"), "w") as f:
        for h, u, v in zip(hdr, units, vals):
            if h in keys:
                f.write(f"{h} = {v} {u}\n")
print("ok")
