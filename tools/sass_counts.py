#!/usr/bin/env python
"""Per-kernel SASS instruction counts of fm3d/libfm3d.so (cuobjdump -sass): the mnemonics that prove tcgen05 / TMA / bulk
copies / packed fp32 are in the built code.  usage: sass_counts.py > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "3d-fm-gan_b200", "fm3d", "libfm3d.so")
cols = [("UTCHMMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("UTMALDG (TMA load)", r"\bUTMALDG"),
        ("UTMASTG (TMA store)", r"\bUTMASTG"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
        ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("STG.E.ENL2.256", r"\bSTG\.E\.ENL2\.256"),
        ("FFMA2/FMUL2/FADD2 (packed fp32)", r"\bF(FMA|MUL|ADD)2\b")]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts, order, cur = collections.defaultdict(lambda: [0] * len(cols)), [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        if cur not in order:
            order.append(cur)
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
        for i, (_, pat) in enumerate(cols):
            if re.search(pat, line):
                counts[cur][i] += 1
print("# SASS instruction counts per kernel of fm3d/libfm3d.so (cuobjdump -sass, sm_100a)")
print("# kernel | " + " | ".join(c for c, _ in cols))
for k in sorted(order):
    if any(counts[k]):                      # kernels with none of these (plain elementwise code) are left out
        print(k + " | " + " | ".join(str(v) for v in counts[k]))
