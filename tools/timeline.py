#!/usr/bin/env python
"""Warm per-call device times of one bench step: wraps every libfm3d entry point with CUDA events
(steady state, caches warm -- complements the cold-cache ncu launch list)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from fm3d import _lib  # noqa: E402
from Util.network_util import Forward_Inference_3_Encoder  # noqa: E402


class Proxy:
    def __init__(self, real):
        self._real = real
        self.records = []
        self.enabled = False

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if not name.startswith("fm_") or name in ("fm_last_error", "fm_launch_count", "fm_version"):
            return fn

        def wrapped(*a):
            if not self.enabled:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            info = ""
            if name == "fm_conv_igemm":
                d = a[0]._obj
                fl = 2.0 * d.B * d.OH * d.OW * d.Cin * d.Cout * d.ntaps
                info = (f"B{d.B} {d.H}x{d.W} {d.Cin}->{d.Cout} taps{d.ntaps} s{max(d.stride, d.stride_y)} out{d.OH}x{d.OW} "
                        f"g{max(d.groups, 1)} epi{'R' if d.rgb else ''}{'S' if d.residual else ''}{'B' if d.border_tab else ''}", fl)
            self.records.append((name, e0, e1, info))
            return r
        return wrapped


def main():
    B = int(os.environ.get("FM3D_PROFILE_BATCH", "32"))
    dev = torch.device("cuda:0")
    real = _lib.lib()
    proxy = Proxy(real)
    _lib._lib = proxy
    e_tsr, e_w, e_wp, g = bench.build_models(dev)
    p, r = [t.to(dev) for t in bench.synthetic_batch(B, 1)]
    with torch.no_grad():
        for _ in range(3):
            Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
        torch.cuda.synchronize()
        proxy.enabled = True
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
        t1.record()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    tot = 0.0
    for i, (name, e0, e1, info) in enumerate(proxy.records):
        us = e0.elapsed_time(e1) * 1e3
        if info:
            name = f"{name} {info[0]}  {info[1] / us / 1e6:7.1f} TFLOP/s"
        tot += us
        key = name.split(" ")[0]
        agg.setdefault(key, [0, 0.0])
        agg[key][0] += 1
        agg[key][1] += us
        if "-v" in sys.argv:
            print(f"{i:4d} {us:9.1f} us {name}")
    print(f"step {t0.elapsed_time(t1) * 1e3:.1f} us, sum of libfm3d calls {tot:.1f} us, {len(proxy.records)} calls")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} us {v[0]:4d}x {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
