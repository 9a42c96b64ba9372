"""Host-side logic that needs no GPU: engine slots, split-K workspace scoping, and the bench.py output contract."""
import json
import os
import subprocess
import sys
import threading

from conftest import ROOT


def test_engine_slot_is_scoped_and_thread_local():
    from fm3d import ops
    assert ops.current_slot() == 0
    with ops.engine_slot(3):
        assert ops.current_slot() == 3
        with ops.engine_slot(1):
            assert ops.current_slot() == 1
        assert ops.current_slot() == 3
        seen = []
        t = threading.Thread(target=lambda: seen.append(ops.current_slot()))
        t.start(); t.join()
        assert seen == [0]                      # another thread keeps the default slot
    assert ops.current_slot() == 0
    try:
        with ops.engine_slot(2):
            raise ValueError
    except ValueError:
        pass
    assert ops.current_slot() == 0              # restored on exceptions too


def test_splitk_scope_restores_owner():
    from fm3d import ops

    class Owner:
        pass
    a, b = Owner(), Owner()
    assert ops._SCOPE.owner is None
    with ops.splitk_scope(a):
        assert ops._SCOPE.owner is a
        with ops.splitk_scope(b):
            assert ops._SCOPE.owner is b
        assert ops._SCOPE.owner is a
    assert ops._SCOPE.owner is None


def test_bench_reference_arm_prints_one_json_line():
    """The driver parses ONE JSON line from stdout; everything else (library banners, warnings) must go to stderr."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["gpu_launches"] == 0 and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
