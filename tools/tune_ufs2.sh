# sweep of the streaming kernel's geometry knobs (strip height cap, ring slot size, row split) per blur width / dtype
out=gpurun_out/ufs/sweep.txt; mkdir -p gpurun_out/ufs; rm -f $out
run() { # W planes dt RMAX KB RS
  echo -n "rmax $4 kb $5 rs $6: " >> $out
  FM3D_UFS_RMAX=$4 FM3D_UFS_SLOT_KB=$5 FM3D_UFS_RS=$6 python tools/prof_upfirdn_w.py $1 $2 $3 2>&1 | tail -1 >> $out
}
for k in "40 36 4" "64 36 4" "64 36 2" "64 36 1" "96 50 4" "96 50 2" "128 72 4"; do run 257 4096 bf16 $k; done
for k in "40 36 4" "64 36 4" "64 36 1" "96 50 4" "128 72 4"; do run 129 8192 f32 $k; done
for k in "40 36 4" "40 36 2" "40 36 1" "40 72 4" "40 72 2"; do run 129 8192 bf16 $k; done
for k in "40 36 4" "40 36 2" "40 36 1" "40 72 2"; do run 65 16384 bf16 $k; done
for k in "40 36 4" "40 36 2" "40 36 1"; do run 65 16384 f32 $k; done
for k in "40 36 4" "96 50 4" "128 72 4"; do run 257 4096 f32 $k; done
cat $out
