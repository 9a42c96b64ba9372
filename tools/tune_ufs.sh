for c in 1 2; do for dt in f32 bf16; do echo -n "cols $c: "; FM3D_UFS_COLS=$c python tools/prof_upfirdn.py $dt | tail -1; done; done
FM3D_UFS_COLS=2 timeout 200 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -2
