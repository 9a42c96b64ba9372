#!/usr/bin/env python
"""One steady-state step of the bench workload inside a cudaProfilerStart/Stop range.
  plain run:  python tools/profile_step.py
  launch list: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
               --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402  (adds the package paths)

from Util.network_util import Forward_Inference_3_Encoder  # noqa: E402


def main():
    B = int(os.environ.get("FM3D_PROFILE_BATCH", "32"))
    dev = torch.device("cuda:0")
    e_tsr, e_w, e_wp, g = bench.build_models(dev)
    p, r = [t.to(dev) for t in bench.synthetic_batch(B, 1)]
    with torch.no_grad():
        for _ in range(3):
            Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profile_step ok")


if __name__ == "__main__":
    main()
