#!/bin/bash
# Round-2 second evidence pass (one B200), after the bandwidth-kernel work: GPU test suite, smoke, bench lines (both arms),
# micro-benchmarks and kernel sweep of the ops, full captures of upfirdn2d_stream / bias_act exported as raw CSV.
set -x
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-400 $O/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python tools/microbench.py ops > $O/microbench_ops.jsonl 2> $O/microbench_ops.err
python tools/kernel_sweep.py --batches 1,32 --quick > $O/kernel_sweep.jsonl 2> $O/kernel_sweep.err
bash tools/prof_ufs_r02b.sh > /dev/null 2>&1; cp gpurun_out/ufs/times.txt $O/upfirdn_widths.txt
cap() { k=$1; kr=$2; shift 2; ncu --set full --clock-control none --import-source on -k regex:$kr --launch-skip 3 --launch-count 1 -o /tmp/$k -f "$@" > $O/ncu_$k.log 2>&1; ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null; }
cap ufs_f32_65 upfirdn2d_stream python tools/prof_upfirdn_w.py 65 16384 f32
cap ufs_f32_257 upfirdn2d_stream python tools/prof_upfirdn_w.py 257 4096 f32
cap ufs_bf16_257 upfirdn2d_stream python tools/prof_upfirdn_w.py 257 4096 bf16
ls -la $O
