"""Per-launch cost of small launches inside a CUDA graph (the strong-scaling limiter, profiles/r02_notes.md): graphs of
100 identical launches of (a) a 1x1 projection at the 8x8 level, (b) a conv3x3 at the 8x8 level, (c) gn_apply at the 8x8
level, 6 videos, with and without programmatic dependent launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import torch
from evcdiff import ops, _lib

DEV = "cuda"
B = int(os.environ.get("B", "6"))
lib = _lib.load()


def gemm(H, C, N, taps, resid=False, stats=False, split="auto"):
    a = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    K = taps * C
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(B, H, H, N, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(B, H, H, N, device=DEV).to(torch.bfloat16) if resid else None
    st = torch.zeros(B, N, 2, device=DEV, dtype=torch.int64) if stats else None
    ws = torch.empty(ops.SPLIT_K_WS_BYTES, dtype=torch.uint8, device=DEV)
    plan = ops.GemmPlan([(a, taps)], w, out, 0, out_ld=N, bias=torch.zeros(N, device=DEV), resid=r, resid_ld=N if resid else 0,
                        stats=st, split_k=split, sk_ws=ws)
    return plan.launch, f"cg={plan.cta_group} split={plan.split_k}"


def gnapply(H, C):
    x = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    y = torch.empty_like(x)
    st = torch.zeros(B, C, 2, device=DEV, dtype=torch.int64)
    ops.gn_stats(x, B, H * H, C, st)
    ss = torch.randn(2 * C, device=DEV) * 0.1
    return (lambda: ops.gn_apply(x, C, None, 0, B, H * H, st, None, 32, 1e-5, ss, True, True, y)), ""


def time_graph(fn, n=100, reps=5):
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


cases = [("1x1 768->768 @8^2 +resid+stats", lambda: gemm(8, 768, 768, 1, True, True)),
         ("1x1 768->1536 @8^2", lambda: gemm(8, 768, 1536, 1)),
         ("conv3x3 768->768 @8^2 +stats", lambda: gemm(8, 768, 768, 9, False, True)),
         ("conv3x3 768->768 @8^2 +stats no split", lambda: gemm(8, 768, 768, 9, False, True, split=1)),
         ("conv3x3 576->576 @16^2 +stats", lambda: gemm(16, 576, 576, 9, False, True)),
         ("conv3x3 384->384 @32^2 +stats", lambda: gemm(32, 384, 384, 9, False, True)),
         ("conv3x3 192->192 @64^2 +stats", lambda: gemm(64, 192, 192, 9, False, True)),
         ("gn_apply 768 ch @8^2", lambda: gnapply(8, 768)),
         ("gn_apply 384 ch @32^2", lambda: gnapply(32, 384))]
for pdl in (0, 1):
    lib.evc_set_pdl(pdl)
    for name, mk in cases:
        fn, info = mk()
        print(f"B={B} pdl={pdl} {name:42s} {info:16s} {time_graph(fn):7.2f} us per launch in a 100-launch graph", flush=True)
