#!/usr/bin/env python
"""Benchmark of the 3-encoder generator forward at 256x256 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One process per GPU (N > 1: launched by torchrun, NCCL only for the barrier / max-over-ranks
timing -- inference shards by batch and needs no data-path collective).
A "step" = one ``Forward_Inference_3_Encoder`` call on a batch of B synthetic photo/render
pairs per GPU (weak scaling).  Prints ONE JSON line on rank 0 (contract in the task brief):

  value     images/s, whole job, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e       same metric through the public API with HOST (pinned) inputs: H2D of photo+render and
            D2H of the fp32 image inside the timed region
  roofline  the dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs of every launch of
            one step / their CUDA-event durations (profiled steps run right after the timed region)
  cpu_baseline  the CPU oracle port (oracle/fm_oracle.py) on a bounded sample, all host cores

``--impl reference`` times that CPU port alone (the reference is Python: it cannot travel to
the GPU box, the port restates its CPU op path and is pinned to it by tests/golden).
"""
import argparse
import json
import os
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "3d-fm-gan_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec 3-encoder G fwd @256²"
UNIT = "images/s"
# algorithmic work per image: full 3-encoder forward
GFLOP_PER_IMAGE = 180.76


def build_models(device, seed=0):
    import resnet_encoder as rn
    import stylegan2
    from psp_encoder_model.encoders import psp_encoders as psp
    torch.manual_seed(seed)
    e_tsr = rn.resnet18(tensor_encoding=True)
    e_w = rn.resnet18(tensor_encoding=False)
    e_wp = psp.GradualStyleEncoder(18, 'ir_se', types.SimpleNamespace(input_nc=3, n_styles=14))
    g = stylegan2.Generator(256, 512, 8, channel_multiplier=2)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # exercise the fused noise / bias terms (they initialise to zero)
        for name, p in g.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    return [m.to(device).eval() for m in (e_tsr, e_w, e_wp, g)]


def synthetic_batch(batch, seed):
    gen = torch.Generator().manual_seed(seed)
    p = torch.rand(batch, 3, 256, 256, generator=gen) * 2 - 1
    r = torch.rand(batch, 3, 256, 256, generator=gen) * 2 - 1
    return p, r


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), read through NVML from a
    background thread (ctypes releases the GIL) a few times per second.  Forking `nvidia-smi` per sample, or
    letting `nvidia-smi -lms` poll at 10 Hz, cost up to 10 % of the device-resident number (driver lock /
    child start-up inside the timed region); the NVML calls here are microseconds each."""

    def __init__(self, index, period=0.1):
        self.index, self.period = index, period
        self.rows = []
        self.stop_flag = None
        self.thread = None
        self.nvml = None
        self.handle = None

    def start(self):
        import threading
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
            return
        self.stop_flag = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((sm, mx, rs))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        n = self.nvml
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        reasons = set()
        for (_, _, rs) in self.rows:
            for k, attr in names.items():
                bit = getattr(n, attr, None)
                if bit is not None and (rs & bit):
                    reasons.add(k)
        sm = sorted(r[0] for r in self.rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in self.rows), "reasons": sorted(reasons),
                "samples": len(self.rows), "source": "NVML"}


def cpu_port_rate(batch, iters, warm):
    """images/s of the CPU oracle port on this host's cores (all threads)."""
    from oracle import fm_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    e_tsr, e_w, e_wp, g = build_models("cpu")
    sds = [{k: v.detach() for k, v in m.state_dict().items()} for m in (e_tsr, e_w, e_wp, g)]
    p, r = synthetic_batch(batch, 123)
    times = []
    with torch.no_grad():
        for i in range(warm + iters):
            t0 = time.perf_counter()
            orc.forward_inference_3_encoder_ref(p, r, *sds, noise=None)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    total = sum(times)
    return batch * len(times) / total, cores, total / len(times)


WORKLOAD = ("3-encoder (E_Tsr+E_W resnet18, E_W_Plus pSp ir_se-18, StyleGAN2 G cm=2) forward 256x256, "
            "batch {B} per GPU, random-init weights, eval mode")

# The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on fd 1 when
# NCCL_DEBUG is set): keep a private handle on the real stdout for the result line and point fd 1 at stderr.
_RESULT_OUT = None


def _protect_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args, rank):
    if rank != 0:
        return
    batch = 2
    rate, cores, sec = cpu_port_rate(batch, max(args.steps, 1), max(min(args.warmup, 1), 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(B=args.batch),
                   "sample": "CPU port of the reference op path (fp32), a bounded sample of the workload: "
                             f"{batch} images per step", "batch_per_step": batch},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {batch} images, fp32, torch CPU threads={cores}"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def main():
    wd = int(os.environ.get("FM3D_BENCH_WATCHDOG", "0"))
    if wd > 0:      # debugging aid: dump every thread's Python stack and exit if the run takes longer than wd seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("FM3D_BENCH_INFLIGHT", "3")),
                    help="batches in flight per GPU: consecutive steps alternate between this many streams (each with "
                         "its own engine plan), so one batch's large kernels fill the SMs another batch's small ones leave idle")
    args = ap.parse_args()
    _protect_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from fm3d import _lib, ops
    from Util.network_util import Forward_Inference_3_Encoder
    _lib.lib()   # fail loudly if the CUDA library is missing

    B = args.batch
    warm = max(args.warmup, 3)
    e_tsr, e_w, e_wp, g = build_models(device, seed=0)

    def step(p, r):
        return Forward_Inference_3_Encoder(p, r, e_tsr, e_w, e_wp, g, tsr_encode='Render Image')

    # several distinct input batches (> L2) so no step re-reads inputs the previous one left in L2
    n_sets = 4
    host = [tuple(t.pin_memory() for t in synthetic_batch(B, 1000 + 17 * rank + i)) for i in range(n_sets)]
    dev_in = [(p.to(device), r.to(device)) for p, r in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    with torch.no_grad():
        # ---------------- device-resident throughput
        # set-up (not a timed or counted step): the engine plans run eagerly twice and are captured into CUDA
        # graphs on the third call, so the W warm-up steps below already replay the steady-state graphs
        NS = max(1, args.inflight)
        main = torch.cuda.current_stream()
        streams = [torch.cuda.Stream(device) for _ in range(NS)]

        def run_steps(n):
            """n steps; step i runs on stream i % NS with engine slot i % NS (its own buffers and CUDA graphs)."""
            out = None
            for s_ in streams:
                s_.wait_stream(main)
            for i in range(n):
                k = i % NS
                with torch.cuda.stream(streams[k]), ops.engine_slot(k):
                    out = step(*dev_in[i % n_sets])
            for s_ in streams:
                main.wait_stream(s_)
            return out

        run_steps(20 * NS)        # also lets clocks / power state settle on a freshly acquired GPU
        torch.cuda.synchronize()
        run_steps(max(warm, NS))
        sampler = ClockSampler(local_rank)
        if os.environ.get("FM3D_BENCH_SAMPLER", "nvml") != "none":
            sampler.start()
        # The timed region is EXACTLY --steps steps between two barriers.  It is measured REPEATS times back to back
        # (a 20-step region lasts ~0.15 s, and run-to-run noise at that length is +-3 %): `value` is the median
        # repeat, the others are reported as `repeats` so the spread is visible.
        REPEATS = 3
        rep_ms, launches = [], 0
        for rep in range(REPEATS):
            barrier()
            l0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run_steps(args.steps)
            e1.record()
            barrier()
            launches = _lib.launch_count() - l0
            rep_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        sampler.stop()
        ms = sorted(rep_ms)[REPEATS // 2]
        value = B * world * args.steps / (ms * 1e-3)

        # ---------------- end to end: pinned host inputs -> device -> image back on the host
        # Every step's photo+render batch is copied from pinned host memory and its fp32 image is
        # copied back, all inside the timed region.  Copies run on a side stream and are double
        # buffered, so the H2D of step i+1 and the D2H of step i-1 overlap the compute of step i.
        h2d, d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
        NB = 2 * NS
        in_bufs = [(torch.empty_like(dev_in[0][0]), torch.empty_like(dev_in[0][1])) for _ in range(NB)]
        out_bufs = [torch.empty(B, 3, 256, 256, device=device) for _ in range(NB)]
        out_hosts = [torch.empty(B, 3, 256, 256).pin_memory() for _ in range(NB)]

        def e2e_loop(n):
            in_ready = [torch.cuda.Event() for _ in range(NB)]
            in_free = [torch.cuda.Event() for _ in range(NB)]
            out_ready = [torch.cuda.Event() for _ in range(NB)]
            out_free = [torch.cuda.Event() for _ in range(NB)]
            for ev in in_free + out_free:
                ev.record(main)
            for s_ in streams + [h2d, d2h]:
                s_.wait_stream(main)

            def fetch(i):
                j = i % NB
                with torch.cuda.stream(h2d):
                    h2d.wait_event(in_free[j])
                    p, r = host[i % n_sets]
                    in_bufs[j][0].copy_(p, non_blocking=True)
                    in_bufs[j][1].copy_(r, non_blocking=True)
                    in_ready[j].record(h2d)
            for i in range(min(NS, n)):
                fetch(i)
            for i in range(n):
                if i + NS < n:
                    fetch(i + NS)
                k, j = i % NS, i % NB
                cs = streams[k]
                with torch.cuda.stream(cs), ops.engine_slot(k):
                    cs.wait_event(in_ready[j])
                    cs.wait_event(out_free[j])
                    img = step(in_bufs[j][0], in_bufs[j][1])
                    out_bufs[j].copy_(img)
                    in_free[j].record(cs)
                    out_ready[j].record(cs)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(out_ready[j])
                    out_hosts[j].copy_(out_bufs[j], non_blocking=True)
                    out_free[j].record(d2h)
            for s_ in streams + [h2d, d2h]:
                main.wait_stream(s_)

        e2e_loop(3 * NS)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        e2e_loop(args.steps)
        e3.record()
        barrier()
        ms_e2e = max_over_ranks(e2.elapsed_time(e3))
        e2e_value = B * world * args.steps / (ms_e2e * 1e-3)

        # ---------------- roofline of the dominant kernel (per-launch CUDA events)
        # The profiled steps run the three encoders back to back on ONE stream (FM3D_STREAMS=0) and without
        # graph replay: with concurrent streams an event pair would also time the wait for SMs another
        # stream's kernel holds (every igemm CTA owns a whole SM), not the kernel itself.
        prof = []
        ops.PROFILE = prof
        prev_env = {k: os.environ.get(k) for k in ("FM3D_STREAMS", "FM3D_PARTITION_SERIAL")}
        os.environ["FM3D_STREAMS"] = "0"
        os.environ["FM3D_PARTITION_SERIAL"] = "1"        # the same launches (SM partitions included) as the timed region
        for i in range(2):
            step(*dev_in[i % n_sets])
        torch.cuda.synchronize()
        # the same step with every launch on all SMs, each timed alone: kernel quality at full width, the figure that is
        # comparable with runs without SM partitions (round 1)
        prof_full = []
        ops.PROFILE = prof_full
        os.environ["FM3D_PARTITION_SERIAL"] = "0"
        for i in range(2):
            step(*dev_in[i % n_sets])
        torch.cuda.synchronize()
        for k, v in prev_env.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
        ops.PROFILE = None
    # A launch confined to n of the device's SMs (fm_conv_desc.max_ctas: the ResNet-18s on 16 SMs each, the W+ encoder
    # on the other 116) is measured against the peak of those n SMs: its duration counts with weight n / SMs -- in the
    # timed region the other SMs run the other networks' kernels at the same time.  The unweighted sum (every launch
    # charged the whole chip for its duration, as if the rest idled) is reported next to it.
    n_sms = torch.cuda.get_device_properties(device).multi_processor_count
    share = lambda cap: (min(cap, n_sms) / n_sms) if cap and cap > 0 else 1.0
    flops = sum(r[2] for r in prof)
    ksec_raw = sum(r[0].elapsed_time(r[1]) for r in prof) * 1e-3
    ksec = sum(r[0].elapsed_time(r[1]) * share(r[4]) for r in prof) * 1e-3
    by_net = {}
    for (a, b, f, tag, cap) in prof:
        e = by_net.setdefault(tag or "other", [0.0, 0.0, 0, 0.0, 0.0])
        dt = a.elapsed_time(b) * 1e-3
        e[0] += f; e[1] += dt * share(cap); e[2] += 1; e[3] += dt; e[4] = share(cap)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_burst = peaks.get("bf16_tflops", 1650.0)
    # dram__bytes_read + dram__bytes_write of the igemm launches of one step / launches, from the committed ncu pass
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "igemm_traffic.json")))
        traffic, traffic_src = tj["bytes_per_launch"], tj["source"]
    except Exception:
        pass
    achieved = flops / ksec / 1e12 if ksec > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "igemm_conv_kernel (tcgen05 implicit GEMM, all launches of one step)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                # the launches are event-timed one by one (each alone on the GPU), so the burst figure is the stricter
                # denominator; the sustained one is what a long step can hold under the power cap
                "peak_burst": peak_burst, "frac_burst": achieved / peak_burst,
                "flops_counted": "algorithmic: 3-channel stems counted with Cin=3 (not their padded K), transposed convs "
                                 "with 9 taps at the input resolution, zero weight blocks not credited",
                # the same figure per network: "generator" = the modulated convolutions (13 3x3 layers incl. the stride-2
                # transposed ones; the 1x1 ToRGBs are fused into their epilogues) -- BASELINE's "modconv TC util"
                "by_network": {k: {"achieved": v[0] / v[1] / 1e12, "frac": v[0] / v[1] / 1e12 / peak_tf,
                                   "frac_burst": v[0] / v[1] / 1e12 / peak_burst, "launches_per_step": v[2] // 2,
                                   "sm_share": v[4], "kernel_ms_per_step": v[3] * 1e3 / 2,
                                   "tflops_on_its_sms": v[0] / v[3] / 1e12,
                                   "algorithmic_gflop_per_step": v[0] / 2 / 1e9}
                               for k, v in sorted(by_net.items()) if v[1] > 0},
                "sm_share_weighted": "a launch confined to n SMs is measured against the peak of n SMs (duration x n/SMs); "
                                     "achieved_unweighted charges every launch the whole chip",
                "full_chip": (lambda f, t: {"achieved": f / t / 1e12, "frac": f / t / 1e12 / peak_tf, "frac_burst": f / t / 1e12 / peak_burst,
                                            "kernel_ms_per_step": t * 1e3 / 2,
                                            "note": "the same launches without SM partitions: every launch on all SMs, timed alone"})(
                    sum(r[2] for r in prof_full), sum(r[0].elapsed_time(r[1]) for r in prof_full) * 1e-3) if prof_full else None,
                "achieved_unweighted": flops / ksec_raw / 1e12 if ksec_raw > 0 else 0.0,
                "frac_unweighted": (flops / ksec_raw / 1e12 / peak_tf) if ksec_raw > 0 else 0.0,
                "launches_per_step": len(prof) // 2, "kernel_ms_per_step": ksec_raw * 1e3 / 2,
                "sm_weighted_ms_per_step": ksec * 1e3 / 2,
                "algorithmic_gflop_per_step": flops / 2 / 1e9,
                "traffic": traffic, "traffic_unit": traffic_src}

    # ---------------- the bandwidth-bound ops of the boundary (BASELINE metric: "upfirdn GB/s"), measured live on rank 0:
    # the largest generator blur through the op API (fp32 NCHW [32,128,257,257] -> 256^2) and fused bias-act on its output
    hbm = None
    if rank == 0:
        import op as op_api
        hbm_peak = peaks.get("hbm_gbs", 6500.0)
        kk = torch.tensor([1., 3., 3., 1.], device=device)
        kk = torch.outer(kk, kk); kk = kk / kk.sum() * 4
        xb = torch.randn(32, 128, 257, 257, device=device)
        bb = torch.zeros(128, device=device)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

        def timed(fn, n=5):
            ts = []
            with torch.no_grad():
                fn()
                for _ in range(n):
                    flush.zero_()                       # > L2: the next iteration finds nothing of its input cached
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); y = fn(); b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e-3)
            return sorted(ts)[len(ts) // 2], y
        t_up, yb = timed(lambda: op_api.upfirdn2d(xb, kk, pad=(1, 1)))
        t_ba, _ = timed(lambda: op_api.fused_leaky_relu(yb, bb))
        up_bytes = (xb.numel() + yb.numel()) * 4
        ba_bytes = 2 * yb.numel() * 4
        hbm = {"peak": hbm_peak, "unit": "GB/s",
               "upfirdn2d_blur_fp32_257": {"achieved": up_bytes / t_up / 1e9, "frac": up_bytes / t_up / 1e9 / hbm_peak,
                                           "ms": t_up * 1e3, "algorithmic_bytes": up_bytes},
               "fused_bias_act_fp32_256": {"achieved": ba_bytes / t_ba / 1e9, "frac": ba_bytes / t_ba / 1e9 / hbm_peak,
                                           "ms": t_ba * 1e3, "algorithmic_bytes": ba_bytes}}
        # the other two HBM-sized generator blurs (the 65^2 / 129^2 planes after the 32->64 and 64->128 up-convs) and the
        # 2-byte walk on the largest one
        for name, shp, dt in (("upfirdn2d_blur_fp32_129", (32, 256, 129, 129), torch.float32),
                              ("upfirdn2d_blur_fp32_65", (32, 512, 65, 65), torch.float32),
                              ("upfirdn2d_blur_bf16_257", (32, 128, 257, 257), torch.bfloat16)):
            del xb, yb
            xb = torch.randn(*shp, device=device).to(dt)
            t_x, yb = timed(lambda: op_api.upfirdn2d(xb, kk, pad=(1, 1)))
            nb = (xb.numel() + yb.numel()) * xb.element_size()
            hbm[name] = {"achieved": nb / t_x / 1e9, "frac": nb / t_x / 1e9 / hbm_peak, "ms": t_x * 1e3, "algorithmic_bytes": nb}
        del xb, yb, flush

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, sec = cpu_port_rate(2, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 timed passes x 2 images (+1 warm-up), fp32 oracle port, torch CPU threads={cores}"}

    img_bytes = B * 3 * 256 * 256 * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(B=B),
                   "global_batch": B * world, "parallelism": f"batch-sharded replicas x{world} (no collective)",
                   "batches_in_flight": f"{NS} per GPU (consecutive steps alternate between {NS} streams, each with its own plan)",
                   "sm_partition": "CTAs per launch (one CTA owns an SM): ResNet-18 {} each / W+ encoder {} / generator {}".format(
                       *[v or "all" for v in ops.sm_partition(device)]),
                   "l2": f"{n_sets} distinct input batches rotate ({n_sets * 2 * img_bytes / 1e6:.0f} MB > L2); "
                         "a step streams > 2 GB of activations"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * img_bytes, "d2h_bytes_per_step": img_bytes,
                "ms_per_step": ms_e2e / args.steps},
        "repeats": {"values": [B * world * args.steps / (m * 1e-3) for m in rep_ms], "unit": UNIT,
                    "spread": (max(rep_ms) - min(rep_ms)) / ms,
                    "note": f"{REPEATS} back-to-back timed regions of exactly {args.steps} steps; value = median"},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "hbm_kernels": hbm,
        "cpu_baseline": cpu,
        "tflops_algorithmic": value * GFLOP_PER_IMAGE / 1e3,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
