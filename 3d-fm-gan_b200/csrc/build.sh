#!/bin/bash
# Build libfm3d.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../fm3d/libfm3d.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/../fm3d" "$HERE/obj"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
       -I"$HERE/../../include" -I"$HERE" --use_fast_math -Xptxas -v "$@")
pids=()
for f in runtime bias_act upfirdn2d igemm wgrad synth encoder_ops; do
  if [ ! -f "$HERE/obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/obj/$f.o" ] || [ "$HERE/common.cuh" -nt "$HERE/obj/$f.o" ] \
     || [ "$HERE/../../include/fm3d.h" -nt "$HERE/obj/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" > "$HERE/obj/$f.log" 2>&1 &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$HERE"/obj/*.log; exit 1; fi
"$NVCC" -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -cudart static \
  "$HERE"/obj/{runtime,bias_act,upfirdn2d,igemm,wgrad,synth,encoder_ops}.o -o "$OUT"
echo "built $OUT"
