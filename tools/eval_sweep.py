#!/usr/bin/env python
"""BASELINE config 5: batch-sharded inference, visual_eval-style (photo, render) pair sweep.
512 pairs are split over the ranks (64 per GPU at 8 GPUs), each rank runs the 3-encoder forward at B = 64 and
converts the images with the device-side tensor2im (Evaluation/visual_eval.py:24-38 -> one kernel per batch);
no collective.  Reports aggregate images/s excluding and including the H2D copy of the inputs and the D2H copy of
the uint8 images (max over ranks, CUDA events).
  one GPU : python tools/eval_sweep.py [--pairs 512] [--batch 64]
  N GPUs  : python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/eval_sweep.py"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402  (adds the package paths)
from fm3d import ops  # noqa: E402
from Util.network_util import Forward_Inference_3_Encoder  # noqa: E402
from Evaluation.visual_eval import tensor2im_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5, help="timed sweeps over the pair set")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    per_rank = args.pairs // world
    nb = max(1, per_rank // B)
    e_tsr, e_w, e_wp, g = bench.build_models(dev, seed=0)
    host = [tuple(t.pin_memory() for t in bench.synthetic_batch(B, 2000 + 31 * rank + i)) for i in range(nb)]
    dev_in = [(p.to(dev), r.to(dev)) for p, r in host]
    out_host = [torch.empty(B, 256, 256, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
    main_s = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(dev) for _ in range(2)]

    def sweep(copies):
        """nb batches; batch i on stream i % 2 (engine slot i % 2); copies: inputs from / images to pinned host memory."""
        for s in streams:
            s.wait_stream(main_s)
        for i in range(nb):
            k = i % 2
            with torch.cuda.stream(streams[k]), ops.engine_slot(k):
                p, r = (host[i][0].to(dev, non_blocking=True), host[i][1].to(dev, non_blocking=True)) if copies else dev_in[i]
                img = Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
                u8 = tensor2im_batch(img)
                if copies:
                    out_host[k].copy_(u8, non_blocking=True)
        for s in streams:
            main_s.wait_stream(s)

    def timed(copies):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            sweep(copies)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        for _ in range(4):                       # plans: two eager calls, capture, replay
            sweep(False)
        sweep(True)
        # ---- correctness before timing: the replayed bf16 engine vs the fp32 composition of the same modules
        # (FM3D_ENGINE=0: cuDNN fp32, TF32 off) on the first 8 pairs, and the uint8 conversion vs the numpy recipe
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(7)
        img = Forward_Inference_3_Encoder(*dev_in[0], e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
        os.environ["FM3D_ENGINE"] = "0"
        try:
            torch.manual_seed(7)                 # same per-layer noise draws (NoiseInjection order, stylegan2.py:308-310)
            lat_t = e_tsr(dev_in[0][1][:8]); lat_w = e_w(dev_in[0][1][:8]); lat_wp = e_wp(dev_in[0][0][:8])
        finally:
            del os.environ["FM3D_ENGINE"]
        import numpy as np
        u8 = tensor2im_batch(img[:2]).cpu().numpy()
        ref8 = ((np.transpose(np.clip(img[:2].cpu().float().numpy(), -1, 1), (0, 2, 3, 1)) + 1.0) * (255.0 / 2.0)).astype(np.uint8)
        if not np.array_equal(u8, ref8):
            raise SystemExit("eval_sweep: tensor2im_batch differs from the numpy recipe of visual_eval.py:24-38")
        with ops.engine_slot(5):
            e_t2, e_w2, e_wp2 = e_tsr(dev_in[0][1][:8]), e_w(dev_in[0][1][:8]), e_wp(dev_in[0][0][:8])
        for nm, a, b in (("E_Tsr", e_t2, lat_t), ("E_W", e_w2, lat_w), ("E_W_Plus", e_wp2, lat_wp)):
            err = float((a - b).abs().max() / b.abs().max())
            if not err < 3e-2:
                raise SystemExit(f"eval_sweep: {nm} engine output differs from the fp32 modules (rel err {err:.3e})")
        if not bool(torch.isfinite(img).all()):
            raise SystemExit("eval_sweep: non-finite image")
        ms_dev = timed(False)
        ms_e2e = timed(True)
    n = world * nb * B * args.reps
    if rank == 0:
        print(json.dumps({"config": "5: batch-sharded 3-encoder inference, uint8 images", "n_gpus": world, "pairs": world * nb * B,
                          "batch_per_gpu": B, "images_per_s_device_resident": n / (ms_dev * 1e-3),
                          "images_per_s_incl_h2d_and_uint8_d2h": n / (ms_e2e * 1e-3),
                          "h2d_bytes_per_image": 2 * 3 * 256 * 256 * 4, "d2h_bytes_per_image": 3 * 256 * 256}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
