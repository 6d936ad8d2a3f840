#!/usr/bin/env python
"""bench.py -- predicted frames/s of the conditional video-diffusion sampling path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our B200 path (default N=1)
    python bench.py --impl reference --steps K --warmup W    the reference's own CPU path (rank 0 only)
    torchrun ... bench.py --gpus N ...                       one rank per GPU, videos sharded, one NCCL gather per step

Workload (config.workload): BASELINE.json configs[1] -- the city_bonn.npy-shaped set of 46 videos (synthetic
stand-in, city_bonn.npy is not shipped), 100-step DDPM (101 UNet evaluations) with the random-init NCSN++ UNet of
configs/mine.yml; one "step" = one sampling cycle = 5 predicted 128x128 frames for every video.  Strong scaling
(default): the 46 videos are sharded by video over the N ranks (6/6/6/6/6/6/5/5 at N=8), every rank samples its
shard with its slice of the global-batch noise (no collective inside the loop), one NCCL gather of the predicted
frames per step; `--scaling weak` gives every rank its own 46-video set.  BASELINE configs[2]:
`--videos 512 --sampler fpndm --subsample 20`.  One JSON line on stdout (rank 0).
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
REF_ROOTS = [os.environ.get("EVC_REF"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def find_reference():
    """Root of an importable copy of the UNMODIFIED reference (its models/ package), or None."""
    for r in REF_ROOTS:
        if r and os.path.isfile(os.path.join(r, "models", "__init__.py")) and \
                os.path.isfile(os.path.join(r, "models", "better", "ncsnpp_more.py")):
            return r
    return None


_REF_NET = None


def reference_cpu_fps(n_steps, threads):
    """The reference's own CPU path: models.ddpm_sampler + models.better.ncsnpp_more.UNetMore_DDPM, imported unmodified
    from baseline/_ref (or $EVC_REF, /root/reference), fp32, B=1, random init (torch.manual_seed(0)).  One call of the
    reference sampler with subsample_steps=n_steps runs n_steps + 1 UNet evaluations (n_steps updates + the final
    denoise evaluation) through its stock code path; the per-evaluation time is extrapolated to the 101 of DDPM-100."""
    import common
    root = find_reference()
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import models as ref_models  # the reference package (ours is evcdiff.models)
    from models.better.ncsnpp_more import UNetMore_DDPM as RefNet
    torch.set_num_threads(threads)
    global _REF_NET
    cfg = common.full_config()
    if _REF_NET is None:
        torch.manual_seed(0)
        _REF_NET = RefNet(cfg).eval()
    torch.manual_seed(1234)
    x = torch.randn(1, 15, 128, 128)
    cond = torch.from_numpy(synthetic_videos(1)[:, :2].reshape(1, 6, 128, 128)).double() / 255.0 * 2 - 1
    with torch.no_grad():
        t0 = time.perf_counter()
        ref_models.ddpm_sampler(x, _REF_NET, cond=cond, final_only=True, denoise=True, subsample_steps=n_steps,
                                clip_before=True, verbose=False, log=False)
        dt = time.perf_counter() - t0
    s_per_eval = dt / (n_steps + 1)
    return 5.0 / (101 * s_per_eval), s_per_eval


def port_cpu_fps(n_evals, threads):
    """Fallback when no copy of the reference is reachable: the oracle port of the same path (oracle/ncsnpp.py +
    the DDPM update of oracle/samplers.py), B=1, `n_evals` evaluations timed, extrapolated to 101."""
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
    betas, alphas, alphas_prev = S.schedule(cfg)
    model = lambda xx, yy: O.ncsnpp_forward(sd, cfg, xx, yy, cond)
    with torch.no_grad():
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


def cpu_baseline(n_evals, one_thread=True):
    """cpu_baseline object of the JSON line: all host cores (bounded sample) + the 1-thread figure -- city_sender.py
    runs with torch.set_num_threads(1) because importing Inference.py executes it (Inference.py:15)."""
    cores = os.cpu_count() or 1
    ref = find_reference()
    if ref is not None:
        reference_cpu_fps(1, cores)  # untimed: page-in, thread pool
        fps, spe = reference_cpu_fps(max(n_evals - 1, 1), cores)
        kind, what = "reference", f"unmodified reference models.ddpm_sampler + UNetMore_DDPM from {os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref}"
    else:
        port_cpu_fps(1, cores)
        fps, spe = port_cpu_fps(n_evals, cores)
        kind, what = "port", "oracle fp32 torch port of the reference (no copy of the reference reachable)"
    out = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
           "sample": f"B=1, {n_evals} of the 101 UNet evaluations of one DDPM-100 cycle ({what}, fp32, {cores} host threads), "
                     f"extrapolated x101; {spe:.3f} s/evaluation"}
    if one_thread:
        f1, s1 = (reference_cpu_fps(1, 1) if ref is not None else port_cpu_fps(1, 1))
        out["value_1thread"] = f1
        out["sample_1thread"] = (f"same code with torch.set_num_threads(1) (what city_sender.py runs with, Inference.py:15): "
                                 f"2 evaluations, {s1:.2f} s/evaluation" if ref is not None else f"1 evaluation, {s1:.2f} s")
        torch.set_num_threads(cores)
    return out


def gpu_eager_context(dev, batch=8, evals=3):
    """CONTEXT number, not a baseline arm and not the product: the same UNet evaluation as plain
    PyTorch eager on this GPU (cuDNN / cuBLAS through oracle/ncsnpp.py, the fp32 restatement of the reference modules),
    in fp32 with TF32 off and under bf16 autocast, `batch` videos, `evals` timed evaluations each, extrapolated to the
    101 of DDPM-100.  Shows what "the reference's modules moved to the B200 unchanged" would give."""
    import common
    from oracle import ncsnpp as O
    cfg = common.full_config(device=dev)
    sd = {k: v.to(dev) for k, v in common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=0, active=False).items()}
    g = torch.Generator(device=dev).manual_seed(1234)
    x = torch.randn(batch, 15, 128, 128, device=dev, generator=g)
    cond = torch.rand(batch, 6, 128, 128, device=dev, generator=g) * 2 - 1
    lab = torch.full((batch,), 500, device=dev, dtype=torch.long)
    out = {}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            with torch.no_grad(), ctx:
                O.ncsnpp_forward(sd, cfg, x, lab, cond)  # untimed: cuDNN algorithm selection
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(evals):
                    O.ncsnpp_forward(sd, cfg, x, lab, cond)
                e1.record()
                torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / evals
            out[name] = {"frames_per_s": batch * 5.0 / (101 * ms / 1e3), "ms_per_evaluation": ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    out["what"] = (f"plain PyTorch eager (cuDNN/cuBLAS) evaluation of the same NCSN++ on this GPU, {batch} videos, {evals} "
                   "evaluations timed, extrapolated x101; context only (BASELINE.md section 4)")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ref = find_reference()
    fn = (lambda n: reference_cpu_fps(max(n - 1, 1), cores)) if ref is not None else (lambda n: port_cpu_fps(n, cores))
    vals = []
    for _ in range(args.warmup):
        fn(2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(fn(args.ref_evals))
    wall = time.perf_counter() - t0
    fps = sum(v[0] for v in vals) / len(vals)
    spe = sum(v[1] for v in vals) / len(vals)
    kind = "reference" if ref is not None else "port"
    what = ("the unmodified reference (models.ddpm_sampler + UNetMore_DDPM, stock code path, imported from "
            f"{os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref})") if ref is not None else "oracle fp32 torch port"
    sample = (f"B=1 (the reference samples one video per call, city_sender.py:526), {args.ref_evals} of the 101 UNet "
              f"evaluations of one DDPM-100 cycle per step, {what} on {cores} host threads, extrapolated x101; "
              f"{spe:.3f} s/evaluation")
    line = {"impl": "reference", "metric": "predicted frames/s (128x128)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    per = "in total, sharded by video over the GPUs" if args.scaling == "strong" else "per GPU"
    # which BASELINE.json config the arguments select (0-based indices as in BASELINE.json `configs`)
    if args.model != "ncsnpp":
        which = "configs[4] (models/unet.py variant)" + ("" if args.videos == 256 else f" at {args.videos} videos")
    elif args.sampler == "ddpm" and args.subsample == 100 and args.videos == 46:
        which = "configs[1]"
    elif args.sampler == "fpndm" and args.subsample == 20 and args.videos == 512:
        which = "configs[2]"
    elif args.sampler == "ddim":
        which = "configs[3] (DDIM step sweep)"
    elif args.sampler == "ddpm" and args.subsample == 100 and args.videos == 1:
        which = "configs[0] (one video per call)"
    else:
        which = "custom (not a BASELINE.json config)"
    return {"workload": f"BASELINE.json {which}: city_bonn-shaped synthetic set, {args.videos} videos {per}, "
                        f"{args.sampler.upper()}-{args.subsample} ({EVALS[args.sampler](args.subsample)} UNet evaluations per "
                        f"cycle), {MODEL_NAME[args.model]}, random init, 5 predicted 128x128 frames per "
                        f"video per step",
            "model": args.model, "precision": args.precision, "scaling": args.scaling,
            "videos": args.videos, "sampler": args.sampler, "subsample": args.subsample,
            "micro_batch": args.micro_batch, "gather": args.gather,
            "parallelism": f"shard-by-video x{world}, global-seed noise sliced per rank, one NCCL gather per step",
            "l2": "working set (>= 2 GB of activations per evaluation) exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): --videos is the GLOBAL set, sharded by video over the ranks (BASELINE configs[1]: "
                         "46 videos over 8 GPUs = 6/6/6/6/6/6/5/5); weak: --videos per rank")
    ap.add_argument("--videos", type=int, default=46, help="videos (city_bonn.npy has 46)")
    ap.add_argument("--micro-batch", type=int, default=64, help="largest batch sampled at once on one GPU")
    ap.add_argument("--sampler", default="ddpm", choices=["ddpm", "ddim", "fpndm"])
    ap.add_argument("--subsample", type=int, default=100)
    ap.add_argument("--model", default="ncsnpp", choices=["ncsnpp", "unet_deep", "unet_deeper"],
                    help="ncsnpp = BASELINE configs[0-3]; unet_* = configs[4] (models/unet.py variant)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16 (default, the metric's dtype) | fp32 = split-bf16 x3 arithmetic (1e-3 tolerance mode)")
    ap.add_argument("--gather", default="fp32", choices=["fp32", "uint8"], help="format of the end-of-step frame gather")
    ap.add_argument("--ref-evals", type=int, default=4)
    ap.add_argument("--cpu-evals", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-launch roofline pass")
    ap.add_argument("--gpu-eager-context", action="store_true",
                    help="also time plain PyTorch eager (fp32 and bf16 autocast) on this GPU: context number, N=1 only")
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

    # strong: every rank holds the same global set and samples its contiguous shard; weak: a rank's own set (seed = rank)
    strong = args.scaling == "strong"
    V = args.videos
    videos = synthetic_videos(V, seed=0 if strong else rank)
    data = torch.from_numpy(videos[:, :2].reshape(V, 6, 128, 128)).double() / 255.0  # city_sender.py:487
    if args.model != "ncsnpp":
        data = data.float()  # models/unet.py never casts its input (a float64 cond crashes the reference's first conv)
    lo, hi = pipeline.shard_range(V, rank, world) if strong else (0, V)
    n_local = hi - lo
    host_in = data.pin_memory()
    host_out = torch.empty((V, 5, 3, 128, 128), dtype=torch.uint8 if args.gather == "uint8" else torch.float32).pin_memory()
    dev_in = host_in.to(dev)
    sampler = args.sampler.upper()
    kw = dict(config=cfg, sampler=sampler, max_batch=args.micro_batch, gather=args.gather)
    weak_buf = None
    if not strong and world > 1 and rank == 0:
        weak_buf = [torch.empty((V, 5, 3, 128, 128), device=dev) for _ in range(world)]

    def step(inp):
        """One sampling cycle of the whole job: 5 predicted frames for every video.  The Gaussian draws are those a
        single GPU would make for the global batch from seed 1234 (x_T, then one draw per non-final DDPM step), sliced
        per rank; predicted frames are gathered on rank 0."""
        if strong:
            return pipeline.generate_frame_sharded(net, inp, rank, world, seed=1234, **kw)
        fr = pipeline.generate_frame_sharded(net, inp, 0, 1, seed=1234 + rank, **kw)
        if world > 1:
            dist.gather(fr.contiguous(), weak_buf, dst=0)
        return fr

    def step_device():
        return step(dev_in)

    def step_e2e():
        fr = step(host_in)  # H2D of this rank's conditioning frames (pinned host memory) inside
        if rank == 0:
            host_out[: fr.shape[0]].copy_(fr, non_blocking=True)  # D2H of the gathered predicted frames
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
    mb = min(args.micro_batch, max(n_local, 1))
    eng = net.engine(mb, dev)
    loop = eng._loop
    per_run = max(loop.launches_per_run.values()) if loop.launches_per_run else 0
    micro = -(-n_local // args.micro_batch)
    gpu_launches = (ops.launch_count() - n0) + args.steps * micro * per_run  # graph replays re-run the captured launches
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None

    frames_per_step = V * 5 * (1 if strong else world)
    value = frames_per_step * args.steps / (ms / 1e3)
    e2e = frames_per_step * args.steps / (ms_e2e / 1e3)
    evals = EVALS[args.sampler](args.subsample)
    tflop_per_frame = GFLOP[args.model] * evals / 5.0 / 1e3

    roofline = None
    if not args.no_profile:
        # ---- roofline of the dominant kernel (evc_gemm_kernel): CUDA events around every launch of one evaluation
        prof = eng.profile(0, reps=3)
        gemm_ms = sum(t for k, m, t in prof if k == "gemm")
        gemm_fl = sum(m["flops"] for k, m, t in prof if k == "gemm")
        tot_ms = sum(t for k, m, t in prof)
        by_kind = {}
        for k, m, t in prof:
            by_kind[k] = by_kind.get(k, 0.0) + t
        achieved_all = gemm_fl / (gemm_ms / 1e3) / 1e12
        # dominant kernel launch: the most expensive GEMM shape of the evaluation (384->192 3x3 at 128x128 for configs/mine.yml)
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
        traffic, traffic_src = None, None
        import glob
        for tpath in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")), reverse=True):
            # DRAM bytes per launch from the newest committed `ncu --set full` capture of this launch shape (static: ncu
            # cannot run inside the timed benchmark)
            with open(tpath) as f:
                tj = json.load(f).get("shapes", {}).get("x".join(str(v) for v in dom_key))
            if tj is not None:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_src = f"static ncu --set full capture, {os.path.relpath(tpath, ROOT)}"
                break
        roofline = {"bound": "tensor", "kernel": "evc_gemm_kernel",
                    "launch": f"M={dom_key[0]} N={dom_key[1]} K={dom_key[2]} ({dom[0]} launches per evaluation, "
                              f"{dom[2] / dom[0] / 1e9:.1f} GFLOP each)",
                    # per-launch CUDA events of an eager pass = kernels timed alone => the burst peak; the whole step
                    # (`step_tensor_frac`, graph replay under the power cap) is held against the sustained one
                    "achieved": achieved, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16_burst"], "frac_of_sustained": achieved / pk["bf16_sustained"],
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": pk["source"] + ", bf16 dense burst (launches timed one by one); step_tensor_frac "
                                   f"against the sustained figure {pk['bf16_sustained']}",
                    "all_gemm_launches": {"achieved": achieved_all, "frac": achieved_all / pk["bf16_burst"],
                                          "frac_of_sustained": achieved_all / pk["bf16_sustained"],
                                          "share_of_eval": gemm_ms / tot_ms},
                    "ms_per_eval_by_kernel": {k: round(v, 3) for k, v in by_kind.items()},
                    "batch_profiled": mb,
                    "step_tensor_frac": value / world * tflop_per_frame / pk["bf16_sustained"]}
        if args.profile_json and rank == 0:
            with open(args.profile_json, "w") as f:
                json.dump([{"kind": k, "ms": t, **({} if m is None else {kk: vv for kk, vv in m.items()})} for k, m, t in prof], f)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(args.cpu_evals)
        line = {"metric": "predicted frames/s (128x128)", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split hi/lo operands, fp32 accumulate)",
                "data": "synthetic", "config": workload_config(args, world), "clocks": clk,
                "e2e": {"value": e2e, "unit": "frames/s",
                        "h2d_bytes_per_step": host_in.numel() * host_in.element_size(),
                        "d2h_bytes_per_step": host_out.numel() * host_out.element_size()},
                "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu,
                **({"gpu_eager_context": gpu_eager_context(dev)} if (args.gpu_eager_context and world == 1) else {}),
                "videos_per_gpu": [pipeline.shard_range(V, r, world)[1] - pipeline.shard_range(V, r, world)[0]
                                   for r in range(world)] if strong else [V] * world,
                "tflops_per_gpu": value / world * tflop_per_frame}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
