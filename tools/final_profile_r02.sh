#!/bin/bash
# Round-2 evidence pass (one B200): bench lines, launch list + per-launch metrics of one step, full captures of the new
# kernels (wgrad, merged-phase up-conv) and of the top conv, timeline, micro-benchmarks, kernel sweep.
set -x
O=gpurun_out/r02; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
FM3D_GRAPH=0 FM3D_STREAMS=0 python tools/timeline.py -v > $O/timeline.txt 2>&1
python tools/microbench.py ops igemm synth upconv train > $O/microbench.jsonl 2> $O/microbench.err
export FM3D_GRAPH=0 FM3D_STREAMS=0
python tools/profile_step.py && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/profile_step.py > $O/ncu_launches.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/step_metrics.csv python tools/profile_step.py > $O/ncu_metrics.log 2>&1
unset FM3D_GRAPH FM3D_STREAMS
python tools/prof_wgrad.py 64 512 512 && ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 4 -c 1 -o $O/wgrad64_full -f python tools/prof_wgrad.py 64 512 512 > $O/ncu_full_wgrad.log 2>&1
python tools/prof_wgrad.py 256 128 128 && ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 4 -c 1 -o $O/wgrad256_full -f python tools/prof_wgrad.py 256 128 128 > $O/ncu_full_wgrad256.log 2>&1
python tools/prof_conv.py 64 512 512 rgb && ncu --set full --clock-control none --import-source on -k regex:igemm_conv -s 4 -c 1 -o $O/igemm64_pair_full -f python tools/prof_conv.py 64 512 512 rgb > $O/ncu_full.log 2>&1
for f in wgrad64_full wgrad256_full igemm64_pair_full; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
python tools/kernel_sweep.py --batches 1,32 --quick > $O/kernel_sweep.jsonl 2> $O/kernel_sweep.err
ls -la $O
