#!/bin/bash
# SM-partition sweep: bench value for ResNet / W+ encoder / generator grid caps (FM3D_PARTITION=r,p,g) and batches in flight.
O=gpurun_out/part; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "repeats", [round(v) for v in d["repeats"]["values"]], flush=True)
except Exception as e:
    print(sys.argv[2], "FAILED", e, flush=True)
PY
}
for cfg in ${CFGS:-"0,0,0" "16,116,0" "12,124,0" "10,128,0" "20,108,0" "16,116,132" "14,120,0"}; do
  for f in ${IFS_:-3}; do
    run p${cfg//,/_}_if$f FM3D_PARTITION=$cfg FM3D_BENCH_INFLIGHT=$f
  done
done
