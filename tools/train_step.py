#!/usr/bin/env python
"""BASELINE config 4: the reference's OWN training-step functions, loaded from its source (train_3_encoder.py lines
1-879, the valid prefix: SURVEY 0.6) and run unchanged on the mirrored modules, one process per GPU:

    Module_To_Train_Setup :311-364      Optimizer_Initilization :405-445
    D_Loss_BackProp :448-477            D_Reg_BackProp :479-493 (R1, double backward)
    G_Loss_BackProp :495-558            G_Reg_BackProp :561-596 (path length, double backward)
    accumulate :195-200 (EMA through .data)

The iteration below is the body of ``train()`` (:786-815) without its data loaders, logging and checkpointing.
``nn.DataParallel`` (:355-362) is replaced by one process per GPU: the optimizers the reference's functions call
``.step()`` on are wrapped so that the bucketed, overlapped all-reduce of Miscellaneous/distributed.py completes first.

  python tools/train_step.py [--batch 32] [--iters 16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py

Needs the reference source: /root/reference in the build container, baseline/_ref (tools/stage_reference.sh) on a GPU
box.  Loss networks whose pretrained blobs are not in the tree (LPIPS-VGG backbone, ArcFace)
run as random-init networks of the reference's own architecture (LPIPS-VGG16 with the shipped linear heads, ArcFace
ResNet-18) on the native conv path; the heat-map term stays off (face_alignment is not installed; hmap_iter_thres = inf
in the reference's config as well).
Synthetic data, random-init weights.  Every convolution of G and D -- forward, dgrad, wgrad, and the second-order passes
of R1 / path length -- runs on the tcgen05 kernels (fm3d/convgrad.py); ``FM3D_NATIVE_GRAD=0`` runs them on ATen.
Prints one JSON line."""
import argparse
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
# an iteration allocates and frees ~50 GB of activations in blocks of very different sizes: without expandable segments the
# caching allocator fragments and falls back to synchronous cudaFree / cudaMalloc (iteration time varied 385 .. 740 ms)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import ref_env  # noqa: E402


class SteppedOptimizer:
    """What the reference's step functions see as ``*_optim``: ``step()`` first completes the gradient all-reduce of
    this parameter group (hooks launched it bucket by bucket during backward), then steps the real optimizer."""

    def __init__(self, optim, reducer):
        self.optim, self.reducer = optim, reducer

    def step(self, *a, **k):
        self.reducer.finish()
        return self.optim.step(*a, **k)

    def __getattr__(self, name):
        return getattr(self.optim, name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="images per GPU (rec_batch)")
    ap.add_argument("--iters", type=int, default=16, help="timed iterations (a multiple of d_reg_every covers both regularisers)")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--d-reg-every", type=int, default=16)
    ap.add_argument("--g-reg-every", type=int, default=4)
    ap.add_argument("--path-batch-shrink", type=int, default=2)
    ap.add_argument("--bucket-mb", type=int, default=32)
    ap.add_argument("--profile", default="", help="write a per-kernel CUDA time table of one full regulariser cycle to this file")
    ap.add_argument("--loss-nets", default="real", choices=["real", "zero"],
                    help="LPIPS-VGG16 / ArcFace-ResNet18 as real (random-init) networks on the native conv path, or zero stubs")
    args_cli = ap.parse_args()

    ts = ref_env.load_train_script()                 # the reference's functions, bound to the mirrored classes
    import numpy as np
    import torch
    import stylegan2
    import train_3_encoder_hyperparams as hp          # the reference's config module
    from Miscellaneous import distributed as D_
    assert ts.Generator is stylegan2.Generator and ts.D_Loss_BackProp.__code__.co_filename.endswith("train_3_encoder.py")

    rank, world, local_rank = D_.init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)                              # identical initial weights on every rank
    np.random.seed(0)
    warnings.filterwarnings("ignore", message=".*add_.*deprecated.*")     # accumulate(): add_(scalar, tensor) overload (:200)

    args = argparse.Namespace(
        device=str(dev), size=args_cli.size, latent=hp.latent, n_mlp=hp.n_mlp, channel_multiplier=hp.channel_multiplier,
        w_plus_encoder_layer_num=hp.w_plus_encoder_layer_num, rec_dataset_type=hp.rec_dataset_type,
        ds_dataset_type=hp.ds_dataset_type, use_separate_D=False, ckpt=None, load_train_state=False,
        distributed=False, gpu_device_ids=[local_rank],          # no nn.DataParallel: one process per GPU
        lr=hp.init_lr, g_reg_every=args_cli.g_reg_every, d_reg_every=args_cli.d_reg_every, use_g_reg=hp.use_g_reg,
        tsr_train=hp.tsr_train, w_train=hp.w_train, w_plus_train=hp.w_plus_train,
        tsr_encode='Render Image', w_plus_sliced_layer=hp.w_plus_sliced_layer, use_tanh=hp.use_tanh,
        r1=hp.discriminator_r1, generator_path_reg_weight=hp.generator_path_reg_weight,
        path_reg_batch_shrink=args_cli.path_batch_shrink, rec_batch=args_cli.batch,
        lpips_loss_lambda=hp.lpips_loss_lambda, l1_loss_lambda=hp.l1_loss_lambda,
        ep_lpips_l1_weight_shrink=hp.ep_lpips_l1_weight_shrink, face_id_loss_lambda=hp.face_id_loss_lambda,
        face_id_loss_type=hp.face_id_loss_type, hmap_loss_lambda=hp.hmap_loss_lambda, hmap_iter_thres=hp.hmap_iter_thres,
        rec_face_reg_loss_lambda=hp.rec_face_reg_loss_lambda, ds_face_reg_loss_lambda=hp.ds_face_reg_loss_lambda,
        ep_face_reg_loss_lambda=hp.ep_face_reg_loss_lambda, ds_freq=hp.ds_freq, ex_ds_freq=hp.ex_ds_freq)

    G, E_Tsr, E_W, E_W_Plus, D, D_edit, g_ema, ckpt = ts.Module_To_Train_Setup(args)         # reference :311-364
    for m in (E_Tsr, E_W, E_W_Plus):
        m.eval()          # what the reference ends up with from its first sampling step on (:678-683)
    g_enc_optim, d_optim, _ = ts.Optimizer_Initilization(args, G, E_Tsr, E_W, E_W_Plus, D, D_edit, ckpt)   # :405-445
    g_params = [p for grp in g_enc_optim.param_groups for p in grp["params"]]
    g_red = D_.GradBucketReducer(g_params, bucket_mb=args_cli.bucket_mb)
    d_red = D_.GradBucketReducer(list(D.parameters()), bucket_mb=args_cli.bucket_mb)
    g_red.enable_timing(); d_red.enable_timing()
    g_step, d_step = SteppedOptimizer(g_enc_optim, g_red), SteppedOptimizer(d_optim, d_red)

    # Frozen loss networks: the reference's own LPIPS (lpips/networks_basic.py:36-101: VGG16 backbone +
    # the linear heads shipped in lpips/weights/v0.1) and ArcFace ResNet-18 (Util/arcface_pytorch), both random-init where
    # their pretrained blobs are not in the tree (the VGG16 backbone is a download, resnet18_arcfacenet.pth is in
    # .MISSING_LARGE_BLOBS).  They sit on the gradient path of the generated image: their convolutions run forward and
    # dgrad on the tcgen05 kernels (fm3d.convgrad.use_native_convs).  --loss-nets zero: stub networks that return 0.
    fa_model = None
    if args_cli.loss_nets == "real":
        import lpips
        from fm3d.convgrad import use_native_convs
        from Util.arcface_pytorch.resnet_face_recognition import resnet_face18
        lpips_model = lpips.PerceptualLoss(model='net-lin', net='vgg', use_gpu=True, gpu_ids=[local_rank])   # :391-392
        face_rec_model = resnet_face18(use_se=False).to(dev)
        ts.requires_grad(face_rec_model, False)
        face_rec_model.eval()
        n_native = use_native_convs(lpips_model.model.net) + use_native_convs(face_rec_model)
    else:
        class ZeroLPIPS(torch.nn.Module):
            def forward(self, a, b):
                return (a[:, :1, :1, :1] * 0).flatten()

        class ZeroFaceNet(torch.nn.Module):
            def forward(self, x):
                return x.mean(dim=(1, 2, 3), keepdim=False)[:, None] * 0
        lpips_model, face_rec_model, n_native = ZeroLPIPS(), ZeroFaceNet(), 0

    B = args_cli.batch
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    batches = [tuple((torch.rand(B, 3, args.size, args.size, generator=gen) * 2 - 1).to(dev) for _ in range(3)) for _ in range(2)]
    state = dict(r1=torch.tensor(0.0, device=dev), path=torch.tensor(0.0, device=dev), mean_path_length=0, ds_count=0)
    loss_dict = {}
    accum = 0.5 ** (32 / (10 * 1000))

    def iteration(iter_idx):                          # body of train(), reference :786-815
        if (iter_idx % args.ds_freq) == (args.ds_freq - 1):
            ds_flag = True
            extreme_ds_flag = ((state["ds_count"] % args.ex_ds_freq) == (args.ex_ds_freq - 1))
            state["ds_count"] += 1
        else:
            ds_flag, extreme_ds_flag = False, False
        g_input, r_input, g_ref = (t.clone() for t in batches[iter_idx % 2])       # stands in for Data_Loading (:361-413)
        ts.D_Loss_BackProp(G, E_Tsr, E_W, E_W_Plus, D, g_input, r_input, g_ref, args, loss_dict, d_step)
        if iter_idx % args.d_reg_every == 0:
            state["r1"] = ts.D_Reg_BackProp(g_ref, D, args, d_step)
        loss_dict['r1'] = state["r1"]
        ts.G_Loss_BackProp(G, E_Tsr, E_W, E_W_Plus, D, g_input, r_input, g_ref, args, loss_dict, g_step, lpips_model,
                           face_rec_model, fa_model, iter_idx, extreme_ds_flag, ds_flag)
        if (iter_idx % args.g_reg_every) == 0 and args.use_g_reg:
            state["path"], _, state["mean_path_length"] = ts.G_Reg_BackProp(G, E_Tsr, E_W, E_W_Plus, g_input, r_input, args,
                                                                            state["mean_path_length"], g_step)
        loss_dict['g_reg'] = state["path"]
        ts.accumulate(g_ema, G, accum)

    for i in range(args_cli.warmup):
        iteration(i)
    if args_cli.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        n_prof = args_cli.g_reg_every
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(n_prof):
                iteration(i)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        with open(args_cli.profile, "w") as f:
            f.write(f"# {n_prof} iterations (D + R1 + G + path length at i = 0), batch {B}/GPU: CUDA kernel time {tot / n_prof / 1e3:.1f} ms/iter\n")
            for e in rows[:60]:
                f.write(f"{e.device_time_total / n_prof / 1e3:9.3f} ms/iter {100 * e.device_time_total / tot:5.1f}% {e.count // n_prof:5d}x  {e.key[:150]}\n")
    g_red.enable_timing(); d_red.enable_timing()
    D_.synchronize()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args_cli.iters):
        iteration(i)
    e1.record()
    D_.synchronize()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    # the EMA generator is what evaluation samples from: it must run on the engine with the weights accumulate() wrote
    with torch.no_grad():
        sample = ts.Forward_Inference_3_Encoder(batches[0][0][:2], batches[0][1][:2], E_Tsr, E_W, E_W_Plus, g_ema,
                                                args.tsr_encode, args.w_plus_sliced_layer, args.use_tanh)
    keys = sorted(k for k, v in loss_dict.items() if torch.is_tensor(v))
    red = D_.reduce_loss_dict({k: loss_dict[k].detach().float().reshape(()) for k in keys})
    exposed = {"g_enc": g_red.exposed_times_ms(), "d": d_red.exposed_times_ms()}
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        finite = all(bool(torch.isfinite(v).all()) for v in red.values()) and bool(torch.isfinite(sample).all())
        ex_total = sum(sum(v) for v in exposed.values())
        print(json.dumps({
            "metric": "train_3_encoder iterations/s (D + R1/%d + G + path-length/%d; the reference's own step functions)"
                      % (args.d_reg_every, args.g_reg_every),
            "value": args_cli.iters / sec, "unit": "it/s",
            "images_per_s": args_cli.iters * B * world / sec, "n_gpus": world, "batch_per_gpu": B, "iters": args_cli.iters,
            "ms_per_iter": sec * 1e3 / args_cli.iters, "size": args.size, "losses_finite": finite,
            "losses": {k: float(v) for k, v in red.items()},
            "native_grad": os.environ.get("FM3D_NATIVE_GRAD", "1") != "0",
            "loss_nets": args_cli.loss_nets, "loss_net_convs_on_native_path": n_native,
            "allreduce": {"world": world, "bucket_mb": args_cli.bucket_mb,
                          "bytes_per_step": {"g_enc": sum(g_red.bucket_bytes()), "d": sum(d_red.bucket_bytes())},
                          "buckets": {"g_enc": len(g_red.buckets), "d": len(d_red.buckets)},
                          "exposed_ms_per_iter": ex_total / args_cli.iters if world > 1 else 0.0,
                          "exposed_fraction_of_iteration": (ex_total / args_cli.iters) / (sec * 1e3 / args_cli.iters) if world > 1 else 0.0,
                          "note": "exposed = time the compute stream waited in finish() for collectives still running when "
                                  "backward ended; the rest of the all-reduce overlapped backward"},
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "source": ts.D_Loss_BackProp.__code__.co_filename}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
