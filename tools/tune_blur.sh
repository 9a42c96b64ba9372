for r in 0 128 64 32 16; do echo "== FM3D_BLUR_ROWS=$r"; FM3D_BLUR_ROWS=$r python tools/microbench.py synth 2>&1 | grep blur_act | cut -c1-110; done
