"""Probe: does running a small local batch as L concurrent lanes (L independent captured graphs on L streams, V/L videos
each) hide the per-launch latency floor?  DDPM-100, V videos on one GPU; prints frames/s for L = 1, 2, 3.
Run with EVC_GEMM_FUSE_GN=0 EVC_GEMM_SPLIT_K=0 for L > 1: ticket / slice spin-waits assume a grid owns the whole GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200"),
                os.path.join(ROOT, "tests")]
import torch
import common
from evcdiff import pipeline
from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM

V = int(os.environ.get("V", "6"))
lanes = [int(x) for x in os.environ.get("LANES", "1,2,3").split(",")]
dev = torch.device("cuda", 0)
cfg = common.full_config(device=dev)
cfg.sampling.subsample = 100
cond = torch.rand(V, 6, 128, 128, device=dev)
for L in lanes:
    nets = []
    for _ in range(L):
        torch.manual_seed(0)
        nets.append(UNetMore_DDPM(cfg).to(dev).eval())
    streams = [torch.cuda.Stream(device=dev) for _ in range(L)]
    x_T, noise = pipeline.global_draws(V, cfg, "DDPM", 100, 1234, dev)
    bounds = [pipeline.shard_range(V, i, L) for i in range(L)]

    def step():
        outs = []
        cur = torch.cuda.current_stream(dev)
        for i, (lo, hi) in enumerate(bounds):
            streams[i].wait_stream(cur)
            with torch.cuda.stream(streams[i]):
                outs.append(pipeline.generate_frame(nets[i], cond[lo:hi], config=cfg, sampler="DDPM", init_samples=x_T[lo:hi],
                                                    noise=[n[lo:hi] for n in noise], to_host=False))
        for s in streams:
            cur.wait_stream(s)
        return torch.cat(outs, 0)

    ref = step(); torch.cuda.synchronize()
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2
    e0.record()
    for _ in range(n):
        out = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"V={V} lanes={L} ({[h - l for l, h in bounds]}): {ms:8.1f} ms/step  {V * 5 / ms * 1e3:7.2f} frames/s  "
          f"finite={bool(torch.isfinite(out).all())} mean={float(out.mean()):.5f}", flush=True)
    del nets
    torch.cuda.empty_cache()
