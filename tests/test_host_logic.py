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


def test_fastdiv_magic_is_exact():
    """The multiply-high division csrc/common.cuh hands to the kernels (FastDivU32: round-up magic with the add-back step),
    restated with Python integers: exact for every 32-bit dividend, for the divisors the launchers can pass."""
    import random

    def make(d):
        L = 0
        while (1 << L) < d:
            L += 1
        return ((1 << 32) * ((1 << L) - d)) // d + 1, min(L, 1), max(L - 1, 0)

    def div(n, f):
        m, s1, s2 = f
        t = (m * n) >> 32
        return (t + (((n - t) & 0xFFFFFFFF) >> s1)) >> s2

    rng = random.Random(0)
    divisors = [1, 2, 3, 5, 7, 8, 9, 63, 64, 65, 129, 255, 257, 1000, 4225, 16641, 65535, 65536, 65537, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1]
    divisors += [rng.randrange(1, 2 ** 32) >> rng.randrange(0, 31) or 1 for _ in range(300)]
    for d in divisors:
        f = make(d)
        assert f[0] < 2 ** 32
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 2, 2 ** 32 - 1] + [rng.randrange(2 ** 32) for _ in range(200)]:
            n &= 0xFFFFFFFF
            assert div(n, f) == n // d, (n, d)


def test_up2_polyphase_index_mapping_matches_the_oracle():
    """The index arithmetic of upfirdn2d_up2_kernel (a thread's 2 x 4 outputs from a 3 x 4 source patch, tap / patch
    indices fixed by the parities of the two leading pads), restated in numpy, against the oracle's upfirdn2d."""
    import numpy as np
    import torch
    from oracle import fm_oracle as orc

    def emulate(x, k, px0, px1, py0, py1):
        kh, kw = k.shape
        ih, iw = x.shape
        oh, ow = ih * 2 + py0 + py1 - kh + 1, iw * 2 + px0 + px1 - kw + 1
        kf = np.zeros((4, 4))
        kf[:kh, :kw] = k[::-1, ::-1]
        PY, PX = py0 & 1, px0 & 1
        yoff, xoff = (PY - py0) >> 1, (PX - px0) >> 1
        out = np.zeros((oh, ow))
        for rp in range((oh + 1) // 2):
            for cq in range((ow + 3) // 4):
                v = np.zeros((3, 4))
                for r in range(3):
                    for c in range(4):
                        iy, ix = rp + yoff + r, 2 * cq + xoff + c
                        if 0 <= iy < ih and 0 <= ix < iw:
                            v[r, c] = x[iy, ix]
                for tr in range(2):
                    if 2 * rp + tr >= oh:
                        break
                    ay, rb = (PY, 0) if tr == 0 else (1 - PY, 1 - PY)
                    for j in range(4):
                        if 4 * cq + j < ow:
                            ax = (PX - j) & 1
                            cb = (j + ax - PX) >> 1
                            out[2 * rp + tr, 4 * cq + j] = (v[rb, cb] * kf[ay, ax] + v[rb, cb + 1] * kf[ay, ax + 2]
                                                             + v[rb + 1, cb] * kf[ay + 2, ax] + v[rb + 1, cb + 1] * kf[ay + 2, ax + 2])
        return out

    rng = np.random.default_rng(0)
    for ih, iw, kh, kw, pads in [(8, 8, 4, 4, (2, 1, 2, 1)), (5, 7, 4, 4, (2, 1, 2, 1)), (6, 5, 3, 4, (1, 2, 0, 3)), (4, 9, 4, 2, (3, 0, 1, 1)),
                                 (7, 7, 4, 4, (0, 0, 0, 0)), (6, 6, 4, 4, (-1, 2, 2, -1)), (3, 3, 2, 2, (1, 1, 1, 1))]:
        x, k = rng.standard_normal((ih, iw)), rng.standard_normal((kh, kw))
        ref = orc.upfirdn2d_ref(torch.from_numpy(x)[None, None].float(), torch.from_numpy(k).float(), 2, 2, 1, 1, *pads)[0, 0].numpy()
        got = emulate(x, k, *pads)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-5
