#!/usr/bin/env python
"""BASELINE config 4: the four update functions of the reference's training iteration
(train_3_encoder.py: D_Loss_BackProp :448-477, D_Reg_BackProp :479-493, G_Loss_BackProp :495-558,
G_Reg_BackProp :561-596) run on the mirrored modules, one process per GPU, gradients averaged with the
bucketed overlapped all-reduce of Miscellaneous/distributed.py (replaces nn.DataParallel,
train_3_encoder.py:355-362).

  python tools/train_step.py [--batch 8] [--iters 16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py

Losses: GAN (non-saturating / logistic), R1 (every d_reg_every), path length (every g_reg_every), L1;
LPIPS / face-identity / heat-map terms need pretrained blobs that are not in the tree and are 0.
Synthetic data, random-init weights.  Prints one JSON line (iterations/s averaged over a cycle that
contains both regularisers).  The gradient-free generator forward inside the D step runs on the fused
bf16 engine; everything under autograd runs the fp32 differentiable composition on the libfm3d ops."""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import resnet_encoder as rn  # noqa: E402
import stylegan2  # noqa: E402
from Miscellaneous import distributed as D_  # noqa: E402
from psp_encoder_model.encoders import psp_encoders as psp  # noqa: E402
from Util.network_util import Forward_Inference_3_Encoder  # noqa: E402


def requires_grad(model, flag=True):          # train_3_encoder.py:190-193
    for p in model.parameters():
        p.requires_grad = flag


def d_logistic_loss(real_pred, fake_pred):    # Util/training_util.py:38-43
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):           # Util/training_util.py:46-52
    grad_real, = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_nonsaturating_loss(fake_pred):          # Util/training_util.py:55-58
    return F.softplus(-fake_pred).mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="images per GPU (rec_batch)")
    ap.add_argument("--iters", type=int, default=16, help="timed iterations (a multiple of d_reg_every covers both regularisers)")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--d-reg-every", type=int, default=16)
    ap.add_argument("--g-reg-every", type=int, default=4)
    ap.add_argument("--path-batch-shrink", type=int, default=2)
    args = ap.parse_args()

    rank, world, local_rank = D_.init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)                         # identical initial weights on every rank
    G = stylegan2.Generator(args.size, 512, 8, channel_multiplier=2).to(dev)
    Dn = stylegan2.Discriminator(args.size, channel_multiplier=2).to(dev)
    E_Tsr = rn.resnet18(tensor_encoding=True).to(dev)
    E_W = rn.resnet18(tensor_encoding=False).to(dev)
    E_WP = psp.GradualStyleEncoder(18, 'ir_se', types.SimpleNamespace(input_nc=3, n_styles=G.n_latent)).to(dev)
    for m in (E_Tsr, E_W, E_WP):
        m.eval()                                 # frozen BN statistics, as the reference ends up with

    g_ratio = args.g_reg_every / (args.g_reg_every + 1)
    d_ratio = args.d_reg_every / (args.d_reg_every + 1)
    g_params = list(G.parameters()) + list(E_Tsr.parameters()) + list(E_W.parameters()) + list(E_WP.parameters())
    g_optim = torch.optim.Adam(g_params, lr=0.002 * g_ratio, betas=(0 ** g_ratio, 0.99 ** g_ratio))      # :405-433
    d_optim = torch.optim.Adam(Dn.parameters(), lr=0.002 * d_ratio, betas=(0 ** d_ratio, 0.99 ** d_ratio))
    g_red = D_.GradBucketReducer(g_params)
    d_red = D_.GradBucketReducer(list(Dn.parameters()))

    B = args.batch
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    batches = [tuple((torch.rand(B, 3, args.size, args.size, generator=gen) * 2 - 1).to(dev) for _ in range(3)) for _ in range(2)]
    mean_path_length = torch.zeros((), device=dev)
    losses = {}

    def fwd(p, r, **kw):
        return Forward_Inference_3_Encoder(p, r, E_Tsr, E_W, E_WP, G, 'Render Image', None, False, **kw)

    def iteration(i):
        nonlocal mean_path_length
        g_input, r_input, g_ref = batches[i % 2]
        # ---- D step (:448-477): generator + encoders frozen -> fused bf16 engine forward
        for m in (G, E_Tsr, E_W, E_WP):
            requires_grad(m, False)
        requires_grad(Dn, True)
        with torch.no_grad():
            fake = fwd(g_input, r_input)
        d_loss = d_logistic_loss(Dn(g_ref), Dn(fake))
        Dn.zero_grad(set_to_none=True)
        d_loss.backward()
        d_red.finish()
        d_optim.step()
        losses["d"] = d_loss.detach()
        # ---- D regularisation (:479-493)
        if i % args.d_reg_every == 0:
            real = g_ref.detach().clone().requires_grad_(True)
            real_pred = Dn(real)
            r1 = d_r1_loss(real_pred, real)
            Dn.zero_grad(set_to_none=True)
            (10.0 / 2 * r1 * args.d_reg_every + 0 * real_pred[0]).backward()
            d_red.finish()
            d_optim.step()
            losses["r1"] = r1.detach()
        # ---- G step (:495-558)
        for m in (G, E_Tsr, E_W, E_WP):
            requires_grad(m, True)
        requires_grad(Dn, False)
        out = fwd(g_input, r_input)
        g_loss = g_nonsaturating_loss(Dn(out))
        l1 = F.l1_loss(out, g_ref)
        for m in (G, E_Tsr, E_W, E_WP):
            m.zero_grad(set_to_none=True)
        (g_loss + l1).backward()
        g_red.finish()
        g_optim.step()
        losses["g"], losses["l1"] = g_loss.detach(), l1.detach()
        # ---- G regularisation (:561-596)
        if i % args.g_reg_every == 0:
            pb = max(1, B // args.path_batch_shrink)
            out, path_lengths = fwd(g_input[:pb], r_input[:pb], PPL_regularize=True)
            path_mean = mean_path_length + 0.01 * (path_lengths.mean() - mean_path_length)
            path_loss = (path_lengths - path_mean).pow(2).mean()
            mean_path_length = path_mean.detach()
            for m in (G, E_Tsr, E_W, E_WP):
                m.zero_grad(set_to_none=True)
            (2.0 * args.g_reg_every * path_loss + 0 * out[0, 0, 0, 0]).backward()
            g_red.finish()
            g_optim.step()
            losses["path"] = path_loss.detach()

    for i in range(args.warmup):
        iteration(i)
    D_.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        iteration(i)
    e1.record()
    D_.synchronize()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    red = D_.reduce_loss_dict(dict(losses))
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        finite = all(bool(torch.isfinite(v).all()) for v in red.values())
        print(json.dumps({
            "metric": "train_3_encoder iterations/s (D + R1/16 + G + path-length/4)", "value": args.iters / sec, "unit": "it/s",
            "images_per_s": args.iters * B * world / sec, "n_gpus": world, "batch_per_gpu": B, "iters": args.iters,
            "ms_per_iter": sec * 1e3 / args.iters, "size": args.size, "losses_finite": finite,
            "losses": {k: float(v) / (world if world > 1 else 1) for k, v in red.items()},
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "note": "fp32 autograd composition on libfm3d ops + bf16 engine for the frozen-generator forward; LPIPS/face-id/heat-map = 0"}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
