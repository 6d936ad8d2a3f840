"""Practical ceiling of a 1 read : 1 write streaming kernel at the sizes gn_apply runs at (context for its roofline
fraction): torch copy_ / mul of bf16 tensors of 72 MB ... 1.16 GB, plus gn_apply itself on the same sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import torch
from evcdiff import ops

DEV = "cuda"


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for B, H, C in ((46, 128, 192), (46, 128, 384), (46, 64, 192), (46, 64, 384), (46, 32, 384), (6, 128, 192)):
    x = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    y = torch.empty_like(x)
    mb = x.numel() * 2 / 1e6
    ms_copy = t(lambda: y.copy_(x))
    ms_mul = t(lambda: torch.mul(x, 1.5, out=y))
    st = torch.zeros(B, C, 2, device=DEV, dtype=torch.int64)
    ops.gn_stats(x, B, H * H, C, st)
    ss = torch.randn(2 * C, device=DEV) * 0.1
    ms_gn = t(lambda: ops.gn_apply(x, C, None, 0, B, H * H, st, None, 32, 1e-5, ss, True, True, y))
    half = C // 2
    xa, xb = x[..., :half].contiguous(), x[..., half:].contiguous()
    sta, stb = st[:, :half].contiguous(), st[:, half:].contiguous()
    ms_gn2 = t(lambda: ops.gn_apply(xa, half, xb, half, B, H * H, sta, stb, 32, 1e-5, ss, True, True, y))
    print(f"B={B} {H}x{H}x{C} {mb:7.1f} MB in + out: copy_ {2 * mb / ms_copy / 1e3:5.2f} TB/s  mul {2 * mb / ms_mul / 1e3:5.2f} TB/s  "
          f"gn_apply {2 * mb / ms_gn / 1e3:5.2f} TB/s ({ms_gn * 1e3:6.1f} us)  gn_apply on a concat {2 * mb / ms_gn2 / 1e3:5.2f} TB/s", flush=True)
