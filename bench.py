#!/usr/bin/env python
"""bench.py -- predicted frames/s of the conditional video-diffusion sampling path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our B200 path (default N=1)
    python bench.py --impl reference --steps K --warmup W    reference CPU implementation (oracle port), rank 0 only
    torchrun ... bench.py --gpus N ...                       one rank per GPU, videos sharded, one NCCL gather per step

Workload (config.workload): BASELINE.json configs[1] -- the city_bonn.npy-shaped set of 46 videos (synthetic
stand-in, city_bonn.npy is not shipped), 100-step DDPM (101 UNet evaluations) with the random-init NCSN++ UNet of
configs/mine.yml; one "step" = one sampling cycle = 5 predicted 128x128 frames for every video of the rank's
shard.  Weak scaling: every rank holds its own 46-video set (no collective inside the loop; one NCCL gather of
the predicted frames per step).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# dense GFLOP per sample per UNet evaluation
GFLOP = {"ncsnpp": 345.20, "unet_deep": 566.17, "unet_deeper": 615.09}
MODEL_NAME = {"ncsnpp": "configs/mine.yml NCSN++ (262.1 M parameters)", "unet_deep": "models/unet.py 'deep' (80.4 M parameters)",
              "unet_deeper": "models/unet.py 'deeper' (240.9 M parameters)"}
EVALS = {"ddpm": lambda s: s + 1, "ddim": lambda s: s + 1, "fpndm": lambda s: 12 + (s - 3)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return dict(bf16_sustained=pk.get("bf16_tflops_sustained", 1365.5), bf16_burst=pk.get("bf16_tflops", 1592.9),
                    hbm=pk.get("hbm_gbs", 6551.0), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def synthetic_videos(n, seed=0, size=128, frames=30):
    """city_bonn.npy-shaped uint8 array (n,30,3,128,128): smooth low-frequency fields translating 1-2 px/frame."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = np.empty((n, frames, 3, size, size), dtype=np.uint8)
    for v in range(n):
        base = torch.from_numpy(rng.normal(size=(1, 3, 8, 8)).astype("float32"))
        img = torch.nn.functional.interpolate(base, size=(size + 64, size + 64), mode="bicubic", align_corners=False)[0]
        img = img + 0.05 * torch.from_numpy(rng.normal(size=img.shape).astype("float32"))
        img = (img - img.min()) / (img.max() - img.min() + 1e-6)
        dx, dy = int(rng.integers(1, 3)), int(rng.integers(1, 3))
        for t in range(frames):
            ox, oy = (t * dx) % 64, (t * dy) % 64
            out[v, t] = (img[:, oy:oy + size, ox:ox + size] * 255).round().clamp(0, 255).to(torch.uint8).numpy()
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


_CPU_SD = None


def cpu_reference_fps(n_evals, threads):
    """Reference CPU path (oracle port of models/__init__.py ddpm_sampler + NCSN++ fp32) on the host cores:
    B=1, `n_evals` UNet evaluations + sampler updates timed, extrapolated to the 101 of a DDPM-100 cycle."""
    import common
    from oracle import ncsnpp as O
    from oracle import samplers as S
    torch.set_num_threads(threads)
    cfg = common.full_config()
    global _CPU_SD
    if _CPU_SD is None:
        _CPU_SD = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=0, active=False)
    sd = _CPU_SD
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 15, 128, 128, generator=g)
    cond = torch.rand(1, 6, 128, 128, generator=g, dtype=torch.float64) * 2 - 1
    sched = S.schedule(cfg)
    model = lambda xx, yy: O.ncsnpp_forward(sd, cfg, xx, yy, cond)
    betas, alphas, alphas_prev = sched
    # one untimed evaluation (page-in, thread pool), then n_evals timed DDPM steps
    with torch.no_grad():
        model(x, torch.zeros(1, dtype=torch.long))
        steps, a, ap, b = S._subsample(alphas, alphas_prev, betas, 100)
        t0 = time.perf_counter()
        for i in range(n_evals):
            lab = (steps[i] * torch.ones(1)).long()
            grad = model(x, lab)
            x0 = ((1 / a[i].sqrt()) * (x - (1 - a[i]).sqrt() * grad)).clip_(-1, 1)
            x = (ap[i].sqrt() * b[i] / (1 - a[i])) * x0 + ((1 - b[i]).sqrt() * (1 - ap[i]) / (1 - a[i])) * x
            x = x + ((1 - ap[i]) / (1 - a[i]) * b[i]).sqrt() * torch.randn_like(x)
        dt = time.perf_counter() - t0
    s_per_eval = dt / n_evals
    return 5.0 / (101 * s_per_eval), s_per_eval


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_reference_fps(1, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fps, spe = cpu_reference_fps(args.ref_evals, cores)
        vals.append((fps, spe))
    wall = time.perf_counter() - t0
    fps = sum(v[0] for v in vals) / len(vals)
    spe = sum(v[1] for v in vals) / len(vals)
    sample = (f"B=1, {args.ref_evals} of the 101 UNet evaluations of one DDPM-100 cycle per step (oracle fp32 torch on "
              f"{cores} host threads), extrapolated x101; {spe:.3f} s/evaluation")
    line = {"impl": "reference", "metric": "predicted frames/s (128x128)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"BASELINE.json configs[1]: city_bonn-shaped synthetic set, {args.videos} videos per GPU, "
                        f"{args.sampler.upper()}-{args.subsample} ({EVALS[args.sampler](args.subsample)} UNet evaluations per "
                        f"cycle), {MODEL_NAME[args.model]}, random init, 5 predicted 128x128 frames per "
                        f"video per step",
            "model": args.model, "precision": args.precision,
            "videos_per_gpu": args.videos, "sampler": args.sampler, "subsample": args.subsample,
            "micro_batch": args.micro_batch, "parallelism": f"shard-by-video x{world}, one NCCL gather per step",
            "l2": "working set (>= 2 GB of activations per evaluation) exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--videos", type=int, default=46, help="videos per GPU (city_bonn.npy has 46)")
    ap.add_argument("--micro-batch", type=int, default=46)
    ap.add_argument("--sampler", default="ddpm", choices=["ddpm", "ddim", "fpndm"])
    ap.add_argument("--subsample", type=int, default=100)
    ap.add_argument("--model", default="ncsnpp", choices=["ncsnpp", "unet_deep", "unet_deeper"],
                    help="ncsnpp = BASELINE configs[1-4]; unet_* = configs[4] (models/unet.py variant)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16 (default, the metric's dtype) | fp32 = split-bf16 x3 arithmetic (1e-3 tolerance mode)")
    ap.add_argument("--ref-evals", type=int, default=4)
    ap.add_argument("--cpu-evals", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-launch timing table of one evaluation here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from evcdiff import ops, pipeline
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    import common

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    cfg = common.full_config(device=dev)
    cfg.sampling.subsample = args.subsample
    torch.manual_seed(0)
    if args.model == "ncsnpp":
        net = UNetMore_DDPM(cfg).to(dev).eval()
    else:
        from evcdiff.models.unet import UNet_DDPM
        cfg.mode = "deep" if args.model == "unet_deep" else "deeper"
        net = UNet_DDPM(cfg).to(dev).eval()
    net.precision = args.precision

    videos = synthetic_videos(args.videos, seed=rank)
    data = torch.from_numpy(videos[:, :2].reshape(args.videos, 6, 128, 128)).double() / 255.0  # city_sender.py:487
    if args.model != "ncsnpp":
        data = data.float()  # models/unet.py never casts its input (a float64 cond crashes the reference's first conv)
    host_in = data.pin_memory()
    host_out = torch.empty((args.videos, 5, 3, 128, 128), dtype=torch.float32).pin_memory()
    dev_in = host_in.to(dev)
    torch.manual_seed(1234 + rank)
    torch.cuda.manual_seed_all(1234 + rank)
    sampler = args.sampler.upper()
    kw = dict(config=cfg, sampler=sampler, max_batch=args.micro_batch)
    gather_buf = None
    if world > 1 and rank == 0:
        gather_buf = [torch.empty((args.videos, 5, 3, 128, 128), device=dev) for _ in range(world)]

    def step_device():
        fr = pipeline.generate_frame(net, dev_in, to_host=False, **kw)
        if world > 1:
            dist.gather(fr.contiguous(), gather_buf, dst=0)
        return fr

    def step_e2e():
        fr = pipeline.generate_frame(net, host_in, to_host=False, **kw)  # H2D of the conditioning frames inside
        if world > 1:
            dist.gather(fr.contiguous(), gather_buf, dst=0)
        host_out.copy_(fr, non_blocking=True)  # D2H of the predicted frames
        return fr

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 1)):
        step_device()
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = ops.launch_count()
    ms = timed(step_device, args.steps)
    eng = net.engine(min(args.micro_batch, args.videos), dev)
    loop = eng._loop
    per_run = max(loop.launches_per_run.values()) if loop.launches_per_run else 0
    micro = -(-args.videos // args.micro_batch)
    gpu_launches = (ops.launch_count() - n0) + args.steps * micro * per_run  # graph replays re-run the captured launches
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None

    frames_per_step = args.videos * 5 * world
    value = frames_per_step * args.steps / (ms / 1e3)
    e2e = frames_per_step * args.steps / (ms_e2e / 1e3)
    evals = EVALS[args.sampler](args.subsample)
    tflop_per_frame = GFLOP[args.model] * evals / 5.0 / 1e3

    # ---- roofline of the dominant kernel (evc_gemm_kernel): CUDA events around every launch of one evaluation
    prof = eng.profile(0, reps=3)
    gemm_ms = sum(t for k, m, t in prof if k == "gemm")
    gemm_fl = sum(m["flops"] for k, m, t in prof if k == "gemm")
    tot_ms = sum(t for k, m, t in prof)
    by_kind = {}
    for k, m, t in prof:
        by_kind[k] = by_kind.get(k, 0.0) + t
    achieved_all = gemm_fl / (gemm_ms / 1e3) / 1e12
    # dominant kernel launch: the most expensive GEMM shape of the evaluation (192->192 3x3 at 128x128 for configs/mine.yml)
    shapes = {}
    for k, m, t in prof:
        if k == "gemm":
            key = (m["M"], m["N"], m["K"])
            e = shapes.setdefault(key, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += t
            e[2] += m["flops"]
    dom_key, dom = max(shapes.items(), key=lambda kv: kv[1][1])
    achieved = dom[2] / (dom[1] / 1e3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath):  # DRAM bytes per launch from the committed ncu --set full capture of this shape
        with open(tpath) as f:
            tj = json.load(f).get("shapes", {}).get("x".join(str(v) for v in dom_key))
        if tj is not None:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    roofline = {"bound": "tensor", "kernel": "evc_gemm_kernel",
                "launch": f"M={dom_key[0]} N={dom_key[1]} K={dom_key[2]} ({dom[0]} launches per evaluation, "
                          f"{dom[2] / dom[0] / 1e9:.1f} GFLOP each)",
                "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + ", bf16 dense sustained",
                "all_gemm_launches": {"achieved": achieved_all, "frac": achieved_all / pk["bf16_sustained"],
                                      "share_of_eval": gemm_ms / tot_ms},
                "ms_per_eval_by_kernel": {k: round(v, 3) for k, v in by_kind.items()},
                "step_tensor_frac": value / world * tflop_per_frame / pk["bf16_sustained"]}
    if args.profile_json and rank == 0:
        with open(args.profile_json, "w") as f:
            json.dump([{"kind": k, "ms": t, **({} if m is None else {kk: vv for kk, vv in m.items()})} for k, m, t in prof], f)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fps, spe = cpu_reference_fps(args.cpu_evals, cores)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"B=1, {args.cpu_evals} of the 101 UNet evaluations of one DDPM-100 cycle (oracle fp32 torch port of "
                             f"the reference, {cores} host threads), extrapolated x101; {spe:.3f} s/evaluation"}
        line = {"metric": "predicted frames/s (128x128)", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split hi/lo operands, fp32 accumulate)",
                "data": "synthetic", "config": workload_config(args, world), "clocks": clk,
                "e2e": {"value": e2e, "unit": "frames/s",
                        "h2d_bytes_per_step": host_in.numel() * host_in.element_size(),
                        "d2h_bytes_per_step": host_out.numel() * host_out.element_size()},
                "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu,
                "tflops_per_gpu": value / world * tflop_per_frame}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
