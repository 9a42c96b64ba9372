#!/bin/bash
# Round-2 evidence pass (one B200): bench lines, launch list + per-launch metrics of one step, full captures (source-level
# stall sampling exported as CSV: .ncu-rep files stay on the box) of the kernels this round changed, timeline,
# micro-benchmarks, kernel sweep.
set -x
O=gpurun_out/r02; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err
FM3D_PARTITION=0 python bench.py --no-cpu-baseline > $O/bench_partition_off.json 2> $O/bench_partition_off.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
FM3D_GRAPH=0 FM3D_STREAMS=0 python tools/timeline.py -v > $O/timeline.txt 2>&1
python tools/microbench.py ops igemm synth upconv train > $O/microbench.jsonl 2> $O/microbench.err
python tools/prof_enc.py all > $O/prof_enc.txt 2>&1
for s in "64 512 256" "32 512 512" "16 512 512"; do python tools/prof_up.py $s; done > $O/prof_up.txt 2>&1
export FM3D_GRAPH=0 FM3D_STREAMS=0 FM3D_PARTITION_SERIAL=1
python tools/profile_step.py && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/profile_step.py > $O/ncu_launches.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/step_metrics.csv python tools/profile_step.py > $O/ncu_metrics.log 2>&1
unset FM3D_GRAPH FM3D_STREAMS FM3D_PARTITION_SERIAL
cap() { k=$1; kr=$2; shift 2; "$@" > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$kr -s 4 -c 1 -o /tmp/$k -f "$@" > $O/ncu_$k.log 2>&1; ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null; ncu -i /tmp/$k.ncu-rep --page source --csv > $O/${k}_source.csv 2>/dev/null; }
cap wgrad64 wgrad_kernel python tools/prof_wgrad.py 64 512 512
cap igemm64_pair igemm_conv python tools/prof_conv.py 64 512 512 rgb
cap up64 igemm_conv python tools/prof_up.py 64 512 256
cap lat2 igemm_conv python tools/prof_enc.py lat2
cap blur256 blur_act python tools/prof_blur.py
python tools/kernel_sweep.py --batches 1,32 --quick > $O/kernel_sweep.jsonl 2> $O/kernel_sweep.err
ls -la $O
