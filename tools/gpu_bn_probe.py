"""In-graph device time of conv3x3 launches whose M tiles quantise badly on 148 SMs (6 videos at 64x64: 192 tiles = 2
rounds for 1.3 rounds of work), for the N tiles the plan could use; and the sampler-update kernels' bandwidth."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import torch
from evcdiff import ops, _lib
from evcdiff._lib import StepCoef

DEV = "cuda"
lib = _lib.load()
lib.evc_set_pdl(1)


def time_graph(fn, n=50, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


def gemm(B, H, C, N, taps, bn):
    a = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    K = taps * C
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(B, H, H, N, device=DEV, dtype=torch.bfloat16)
    st = torch.zeros(B, N, 2, device=DEV, dtype=torch.int64)
    ws = torch.empty(ops.SPLIT_K_WS_BYTES, dtype=torch.uint8, device=DEV)
    plan = ops.GemmPlan([(a, taps)], w, out, 0, out_ld=N, bias=torch.zeros(N, device=DEV), stats=st, split_k="auto" if bn is None else 1,
                        sk_ws=ws, bn=bn)
    return plan


for B, H, C, N in [(6, 64, 192, 192), (6, 64, 384, 192), (6, 64, 384, 384), (6, 128, 192, 192), (5, 128, 192, 192), (6, 32, 384, 384)]:
    for bn in (None, 192, 96, 64):
        if bn is not None and N % bn:
            continue
        p = gemm(B, H, C, N, 9, bn)
        us = time_graph(p.launch)
        print(f"B={B} conv3x3 {C}->{N} @{H}^2 bn={bn} cg={p.cta_group} split={p.split_k}: {us:7.2f} us  "
              f"{p.flops / us / 1e6:7.0f} TFLOP/s", flush=True)

# sampler update (DDPM step): reads x, eps, noise (fp32 NCHW planes), writes x (fp32) + the bf16 NHWC UNet input rows
for B in (46, 6):
    x = torch.randn(B, 15, 128, 128, device=DEV)
    eps = torch.randn_like(x); nz = torch.randn_like(x)
    xin = torch.zeros(B, 128, 128, 64, device=DEV, dtype=torch.bfloat16)
    c = StepCoef(0, 1, 1.01, 0.1, 0.5, 0.5, 0.0, 0.1)
    fn = lambda: ops.sampler_update(x, eps, nz, x, xin, c)
    us = time_graph(fn)
    nbytes = 4 * x.numel() * 4 + B * 128 * 128 * 15 * 2
    print(f"B={B} sampler_update (DDPM): {us:7.2f} us  {nbytes / us / 1e3:7.0f} GB/s of {nbytes / 1e6:.1f} MB", flush=True)
