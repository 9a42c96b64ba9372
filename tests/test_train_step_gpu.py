"""BASELINE config 4 smoke: the reference's four update functions (D, D-R1, G, G-path-length;
train_3_encoder.py:448-596) run end to end on the mirrored modules, with a tiny batch so the
test stays in seconds.  Checks that every loss is finite and that parameters actually moved."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_iteration_runs(cuda):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_step.py"), "--size", "256", "--batch", "2",
                          "--iters", "4", "--warmup", "1", "--d-reg-every", "2", "--g-reg-every", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["losses_finite"], line
    assert set(line["losses"]) >= {"d", "r1", "g", "l1", "path"}
    assert line["value"] > 0
