#!/bin/bash
# SM-partition experiment: bench value for ResNet / pSp / generator grid caps and batches in flight (one B200).
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
run base FM3D_X=0
for cfg in "16 116 116" "12 124 124" "16 116 148" "16 100 100" "16 74 74" "12 62 62" "0 74 74" "16 58 58" "16 116 74" "16 74 116" "20 108 108" "16 132 132"; do
  set -- $cfg
  for f in 2 3; do
    run r$1_p$2_g$3_if$f FM3D_RESNET_CTAS=$1 FM3D_MAIN_CTAS=$2 FM3D_GEN_CTAS=$3 FM3D_BENCH_INFLIGHT=$f
  done
done
run r16_p74_g74_if4 FM3D_RESNET_CTAS=16 FM3D_MAIN_CTAS=74 FM3D_GEN_CTAS=74 FM3D_BENCH_INFLIGHT=4
run r16_p58_g58_if4 FM3D_RESNET_CTAS=16 FM3D_MAIN_CTAS=58 FM3D_GEN_CTAS=58 FM3D_BENCH_INFLIGHT=4
