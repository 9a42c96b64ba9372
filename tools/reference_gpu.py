#!/usr/bin/env python
"""GPU reference baseline (SURVEY 8d last row): the UNMODIFIED reference -- its stylegan2.py,
encoders, Util/network_util.py and its own op/*.cu JIT-built for sm_100a -- timed on the same B200, same inputs, same
harness as bench.py, with torch defaults (cuDNN TF32 on, "as shipped") and with TF32 off (strict fp32); plus parity of
this repo's kernels against the reference's CUDA ops and layers.

The reference cannot be imported next to the mirror (same module names), so the work is split over two processes:

  python tools/reference_gpu.py ref   [--batch 32]   # sys.path = baseline/_ref only; writes /tmp/fm3d_ref_gpu.pt
  python tools/reference_gpu.py ours  [--batch 32]   # mirror; reads that file, writes gpurun_out/r02_reference_gpu.json
  python tools/reference_gpu.py both                 # runs the two as subprocesses
  python tools/reference_gpu.py jit                  # (container, no GPU) pre-build the reference's two extensions

Needs the staged reference (tools/stage_reference.sh -> baseline/_ref, git-ignored).  Measurement tool: not on the
product path, not used by tests / bench.py / smoke().
"""
import argparse
import json
import os
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
OUT_DIR = os.path.join(ROOT, "gpurun_out")
PT = os.path.join(os.environ.get("TMPDIR", "/tmp"), "fm3d_ref_gpu.pt")      # hand-over file between the two processes (large)


def build_models(rn, sg, psp, device, seed=0):
    """Same construction order and seeds as bench.build_models (the mirror draws parameters in the reference's order)."""
    import torch
    torch.manual_seed(seed)
    e_tsr = rn.resnet18(tensor_encoding=True)
    e_w = rn.resnet18(tensor_encoding=False)
    e_wp = psp.GradualStyleEncoder(18, 'ir_se', types.SimpleNamespace(input_nc=3, n_styles=14))
    g = sg.Generator(256, 512, 8, channel_multiplier=2)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in g.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    return [m.to(device).eval() for m in (e_tsr, e_w, e_wp, g)]


def synthetic_batch(batch, seed):
    import torch
    gen = torch.Generator().manual_seed(seed)
    p = torch.rand(batch, 3, 256, 256, generator=gen) * 2 - 1
    r = torch.rand(batch, 3, 256, 256, generator=gen) * 2 - 1
    return p, r


def time_forward(fn, warm, iters):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return {"median_ms": ts[len(ts) // 2], "best_ms": ts[0], "iters": iters}


def op_cases():
    """The upfirdn2d configurations the model uses + bias-act, small enough to store."""
    return [("blur_up pad(1,1) k*4", dict(up=1, down=1, pad=(1, 1)), (2, 16, 33, 33), 4.0),
            ("torgb_up up2 pad(2,1) k*4", dict(up=2, down=1, pad=(2, 1)), (2, 3, 32, 32), 4.0),
            ("d_blur pad(2,2)", dict(up=1, down=1, pad=(2, 2)), (2, 16, 32, 32), 1.0),
            ("d_skip_blur pad(1,1)", dict(up=1, down=1, pad=(1, 1)), (2, 16, 32, 32), 1.0),
            ("down2 pad(1,1)", dict(up=1, down=2, pad=(1, 1)), (2, 3, 64, 64), 1.0)]


def run_ops(op, device, dtypes):
    """op-level outputs on seeded inputs (forward + first-order gradient), per dtype."""
    import torch
    res = {}
    k1 = torch.tensor([1., 3., 3., 1.])
    k = torch.outer(k1, k1); k = k / k.sum()
    for dt in dtypes:
        for name, cfg, shape, gain in op_cases():
            g = torch.Generator().manual_seed(len(name))
            x = torch.randn(*shape, generator=g).to(device=device, dtype=dt).requires_grad_(True)
            kern = (k * gain).to(device)          # the reference keeps the FIR buffer fp32? no: it follows the module dtype; fp32 here
            y = op.upfirdn2d(x, kern.to(dt), **cfg)
            gy = torch.randn(y.shape, generator=g).to(device=device, dtype=dt)
            gx, = torch.autograd.grad(y, x, gy)
            res[f"upfirdn2d/{name}/{dt}"] = (y.detach().float().cpu(), gx.detach().float().cpu())
        g = torch.Generator().manual_seed(77)
        x = torch.randn(3, 24, 19, 19, generator=g).to(device=device, dtype=dt).requires_grad_(True)
        b = torch.randn(24, generator=g).to(device=device, dtype=dt).requires_grad_(True)
        y = op.fused_leaky_relu(x, b)
        gy = torch.randn(y.shape, generator=g).to(device=device, dtype=dt)
        gx, gb = torch.autograd.grad(y, [x, b], gy)
        res[f"fused_leaky_relu/{dt}"] = (y.detach().float().cpu(), gx.detach().float().cpu(), gb.detach().float().cpu())
    return res


def layer_dump(g, lat, ext, noise):
    """Per-layer activations of the generator (StyledConv outputs, ToRGB outputs) via forward hooks, B small."""
    import torch
    acts, hooks = {}, []

    def hook(name):
        def f(mod, inp, out):
            o = out[0] if isinstance(out, (tuple, list)) else out
            acts[name] = o.detach().float().cpu()
        return f
    hooks.append(g.conv1.register_forward_hook(hook("conv1")))
    hooks.append(g.to_rgb1.register_forward_hook(hook("to_rgb1")))
    for i, m in enumerate(g.convs):
        hooks.append(m.register_forward_hook(hook(f"convs.{i}")))
    for i, m in enumerate(g.to_rgbs):
        hooks.append(m.register_forward_hook(hook(f"to_rgbs.{i}")))
    with torch.no_grad():
        img = g(None, latent_styles=[lat], input_is_latent=True, noise=noise, use_external_input_tensor=True,
                external_input_tensor=ext)
    for h in hooks:
        h.remove()
    acts["image"] = img.detach().float().cpu()
    return acts


def small_inputs(device):
    import torch
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(2, 14, 512, generator=g).to(device)
    ext = torch.randn(2, 512, 4, 4, generator=g).to(device)
    noise = [torch.randn(2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=g).to(device) for i in range(13)]
    return lat, ext, noise


# ------------------------------------------------------------------------------------------------ reference process
def main_ref(args):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    # build directory inside the staged (git-ignored) tree: a build made in the container (``reference_gpu.py jit``)
    # travels to the GPU box with the snapshot, and ninja finds nothing to do there
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(REF, "_torch_ext"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_env
    ref_env.install_shims()
    sys.path.insert(0, REF)
    import torch
    t0 = time.time()
    import op                                   # JIT-builds op/fused_bias_act*.{cpp,cu}, op/upfirdn2d*.{cpp,cu} for sm_100a
    jit_s = time.time() - t0
    if args.mode == "jit":
        print(f"reference extensions built in {jit_s:.0f} s under {os.environ['TORCH_EXTENSIONS_DIR']}")
        return
    import stylegan2 as sg
    import resnet_encoder as rn
    from psp_encoder_model.encoders import psp_encoders as psp
    from Util import network_util as nu
    assert op.__file__.startswith(REF) and sg.__file__.startswith(REF), (op.__file__, sg.__file__)
    dev = torch.device("cuda:0")
    B = args.batch
    e_tsr, e_w, e_wp, g = build_models(rn, sg, psp, dev)
    G = torch.nn.DataParallel(g, device_ids=[0])            # Forward_Inference_3_Encoder touches g_ema.module (:317-318)
    p, r = [t.to(dev) for t in synthetic_batch(B, 1000)]
    out = {"jit_seconds": jit_s, "batch": B, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}

    def fwd():
        with torch.no_grad():
            return nu.Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, G, tsr_encode='Render Image')

    def g_only(lat, ext):
        with torch.no_grad():
            return G(None, latent_styles=[lat], input_is_latent=True, use_external_input_tensor=True, external_input_tensor=ext)
    gg = torch.Generator().manual_seed(3)
    latB = torch.randn(B, 14, 512, generator=gg).to(dev); extB = torch.randn(B, 512, 4, 4, generator=gg).to(dev)
    timing = {}
    for label, tf32 in (("tf32_on (torch defaults, as shipped)", True), ("tf32_off (strict fp32)", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False if not tf32 else torch.backends.cuda.matmul.allow_tf32
        t = time_forward(fwd, args.warmup, args.iters)
        t["images_per_s"] = B / (t["median_ms"] * 1e-3)
        tg = time_forward(lambda: g_only(latB, extB), args.warmup, args.iters)
        tg["images_per_s"] = B / (tg["median_ms"] * 1e-3)
        timing[label] = {"three_encoder_forward": t, "generator_only": tg}
        print(label, json.dumps(timing[label]), flush=True)
    out["timing"] = timing
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    # ---- parity material (strict fp32): encoder outputs + image for the first 4 pairs, per-layer dump at B=2, op outputs
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    noise4 = [torch.randn(4, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=torch.Generator().manual_seed(40 + i)).to(dev)
              for i in range(13)]
    with torch.no_grad():
        t4, w4, wp4 = e_tsr(r[:4]), e_w(r[:4]), e_wp(p[:4])
        lat4 = torch.stack([w4 * wp4[:, i, :] for i in range(14)]).transpose(0, 1)
        img4 = g(None, latent_styles=[lat4], input_is_latent=True, noise=noise4, use_external_input_tensor=True,
                 external_input_tensor=t4)
    out["enc"] = {"e_tsr": t4.cpu(), "e_w": w4.cpu(), "e_wp": wp4.cpu(), "img": img4.cpu()}
    out["layers"] = layer_dump(g, *small_inputs(dev))
    out["ops"] = run_ops(op, dev, [torch.float32, torch.float16])
    os.makedirs(OUT_DIR, exist_ok=True)
    torch.save(out, PT)
    print("reference process done:", PT, flush=True)


# ------------------------------------------------------------------------------------------------ mirror process
def main_ours(args):
    sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
    import torch
    import op
    import stylegan2 as sg
    import resnet_encoder as rn
    from psp_encoder_model.encoders import psp_encoders as psp
    from Util import network_util as nu
    ref = torch.load(PT, weights_only=False)
    dev = torch.device("cuda:0")
    B = ref["batch"]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    e_tsr, e_w, e_wp, g = build_models(rn, sg, psp, dev)
    p, r = [t.to(dev) for t in synthetic_batch(B, 1000)]
    rel = lambda a, b: float((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-20))
    mabs = lambda a, b: float((a.float().cpu() - b.float().cpu()).abs().max())

    def fwd():
        with torch.no_grad():
            return nu.Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')
    gg = torch.Generator().manual_seed(3)
    latB = torch.randn(B, 14, 512, generator=gg).to(dev); extB = torch.randn(B, 512, 4, 4, generator=gg).to(dev)

    def g_only():
        with torch.no_grad():
            return g(None, latent_styles=[latB], input_is_latent=True, use_external_input_tensor=True, external_input_tensor=extB)
    ours_t = time_forward(fwd, max(args.warmup, 4), args.iters)
    ours_t["images_per_s"] = B / (ours_t["median_ms"] * 1e-3)
    ours_g = time_forward(g_only, max(args.warmup, 4), args.iters)
    ours_g["images_per_s"] = B / (ours_g["median_ms"] * 1e-3)
    # ---- parity: engine (bf16) and fp32 composition vs the reference's CUDA path (strict fp32)
    noise4 = [torch.randn(4, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=torch.Generator().manual_seed(40 + i)).to(dev)
              for i in range(13)]
    parity = {}
    for mode in ("engine_bf16", "composition_native_bf16", "composition_fp32"):
        if mode != "engine_bf16":
            os.environ["FM3D_ENGINE"] = "0"           # module-by-module composition (the autograd-capable path) ...
        if mode == "composition_fp32":
            os.environ["FM3D_NATIVE_GRAD"] = "0"      # ... on ATen convolutions in strict fp32: pins the algebra itself
        try:
            with torch.no_grad():
                t4, w4, wp4 = e_tsr(r[:4]), e_w(r[:4]), e_wp(p[:4])
                lat4 = w4.unsqueeze(1) * wp4
                # the generator is fed the REFERENCE's encoder outputs so its error is its own
                rt, rw, rwp = [ref["enc"][k].to(dev) for k in ("e_tsr", "e_w", "e_wp")]
                img4 = g(None, latent_styles=[rw.unsqueeze(1) * rwp], input_is_latent=True, noise=noise4,
                         use_external_input_tensor=True, external_input_tensor=rt)
        finally:
            os.environ.pop("FM3D_ENGINE", None)
            os.environ.pop("FM3D_NATIVE_GRAD", None)
        parity[mode] = {"e_tsr_rel": rel(t4, ref["enc"]["e_tsr"]), "e_w_rel": rel(w4, ref["enc"]["e_w"]),
                        "e_wp_rel": rel(wp4, ref["enc"]["e_wp"]), "generator_image_rel": rel(img4, ref["enc"]["img"]),
                        "generator_image_max_abs": mabs(img4, ref["enc"]["img"])}
    os.environ["FM3D_ENGINE"] = "0"
    os.environ["FM3D_NATIVE_GRAD"] = "0"
    try:
        mine = layer_dump(g, *small_inputs(dev))
    finally:
        os.environ.pop("FM3D_ENGINE", None)
        os.environ.pop("FM3D_NATIVE_GRAD", None)
    layers = {k: {"max_abs": mabs(mine[k], v), "rel": rel(mine[k], v)} for k, v in ref["layers"].items()}
    with torch.no_grad():
        lat, ext, noise = small_inputs(dev)
        rgbs = g(None, latent_styles=[lat], input_is_latent=True, noise=noise, use_external_input_tensor=True,
                 external_input_tensor=ext, return_rgb_list=True)
    names = ["to_rgb1"] + [f"to_rgbs.{i}" for i in range(6)]
    engine_rgbs = {n: {"max_abs": mabs(a, ref["layers"][n]), "rel": rel(a, ref["layers"][n])} for n, a in zip(names, rgbs)}
    mine_ops = run_ops(op, dev, [torch.float32, torch.float16])
    ops_cmp = {}
    for k, v in ref["ops"].items():
        ops_cmp[k] = {"max_abs": [mabs(a, b) for a, b in zip(mine_ops[k], v)], "rel": [rel(a, b) for a, b in zip(mine_ops[k], v)],
                      "fields": ["out", "grad_input"] + (["grad_bias"] if len(v) == 3 else [])}
    res = {
        "what": "unmodified reference (its own op/*.cu JIT-built for sm_100a, cuDNN convs) vs this repo, same B200, same "
                "weights (seed 0) and inputs; 3-encoder forward 256x256, CUDA events, median",
        "batch": B, "reference": {"timing": ref["timing"], "jit_seconds": ref["jit_seconds"], "peak_mem_gb": ref["peak_mem_gb"],
                                  "torch": ref["torch"], "cudnn": ref["cudnn"]},
        "ours": {"three_encoder_forward": ours_t, "generator_only": ours_g,
                 "note": "single stream, one batch in flight, CUDA-graph replay (bench.py keeps two batches in flight)"},
        "speedup_vs_reference_tf32_on": {
            "three_encoder_forward": ref["timing"]["tf32_on (torch defaults, as shipped)"]["three_encoder_forward"]["median_ms"] / ours_t["median_ms"],
            "generator_only": ref["timing"]["tf32_on (torch defaults, as shipped)"]["generator_only"]["median_ms"] / ours_g["median_ms"]},
        "speedup_vs_reference_tf32_off": {
            "three_encoder_forward": ref["timing"]["tf32_off (strict fp32)"]["three_encoder_forward"]["median_ms"] / ours_t["median_ms"],
            "generator_only": ref["timing"]["tf32_off (strict fp32)"]["generator_only"]["median_ms"] / ours_g["median_ms"]},
        "parity_vs_reference_cuda_fp32": parity,
        "per_layer_fp32_composition_vs_reference": layers,
        "per_resolution_rgb_engine_bf16_vs_reference": engine_rgbs,
        "ops_vs_reference_cuda_ops": ops_cmp,
    }
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "r02_reference_gpu.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: res[k] for k in ("speedup_vs_reference_tf32_on", "speedup_vs_reference_tf32_off", "parity_vs_reference_cuda_fp32")}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["ref", "ours", "both", "jit"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    if not os.path.isfile(os.path.join(REF, "stylegan2.py")):
        raise SystemExit("no staged reference: run tools/stage_reference.sh in the build container first")
    if args.mode == "both":
        for m in ("ref", "ours"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), m, "--batch", str(args.batch),
                                   "--warmup", str(args.warmup), "--iters", str(args.iters)])
        return
    (main_ours if args.mode == "ours" else main_ref)(args)


if __name__ == "__main__":
    main()
