# upfirdn2d_stream A/B: timings at the generator's blur widths (fp32 / bf16), optional ncu captures (NCU=1)
mkdir -p gpurun_out/ufs
rm -f gpurun_out/ufs/times.txt
for cfg in "65 16384 f32" "129 8192 f32" "257 4096 f32" "65 16384 bf16" "129 8192 bf16" "257 4096 bf16" "33 32768 f32" "33 32768 bf16"; do
  set -- $cfg
  python tools/prof_upfirdn_w.py $1 $2 $3 2>&1 | tail -1 >> gpurun_out/ufs/times.txt
done
if [ "${NCU:-0}" = "1" ]; then
i=0
for cfg in "129 8192 f32" "257 4096 bf16"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:upfirdn2d_stream --launch-skip 3 --launch-count 1 -o gpurun_out/ufs/cap$i python tools/prof_upfirdn_w.py $1 $2 $3 > gpurun_out/ufs/ncu$i.log 2>&1
  ncu -i gpurun_out/ufs/cap$i.ncu-rep --page raw --csv > gpurun_out/ufs/raw$i.csv 2>/dev/null
  ncu -i gpurun_out/ufs/cap$i.ncu-rep --page source --csv > gpurun_out/ufs/src$i.csv 2>/dev/null
  rm -f gpurun_out/ufs/cap$i.ncu-rep
  i=$((i+1))
done
fi
cat gpurun_out/ufs/times.txt
