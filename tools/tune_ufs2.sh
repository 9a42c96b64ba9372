# sweep of the streaming kernel's row-split threshold (threads that must have a task before rows stop being split)
out=gpurun_out/ufs/sweep.txt; mkdir -p gpurun_out/ufs; rm -f $out
run() { # W planes dt BUSY
  echo -n "busy $4: " >> $out
  FM3D_UFS_BUSY=$4 python tools/prof_upfirdn_w.py $1 $2 $3 2>&1 | tail -1 >> $out
}
for b in 128 192; do run 129 8192 f32 $b; run 257 4096 bf16 $b; run 65 16384 bf16 $b; run 65 16384 f32 $b; run 33 32768 f32 $b; run 33 32768 bf16 $b; done
run 129 8192 bf16 256
cat $out
