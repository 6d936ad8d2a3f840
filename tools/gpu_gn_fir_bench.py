"""Device time of the fused res-block prologue gn_fir (FIR(SiLU(GN(x))) + FIR(x)) at the shapes of the benchmark batch.
EVC_GN_FIR_DOWN=0 selects the older down-sampling walk (two activations per input element) for an A/B in two processes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import torch
from evcdiff import ops

DEV = "cuda"
B = int(os.environ.get("B", "46"))
cases = [(128, 192, 0, False), (64, 192, 0, False), (32, 384, 0, False), (16, 576, 0, False),
         (64, 192, 192, True), (32, 384, 192, True), (16, 576, 384, True), (8, 768, 576, True)]
for H, C0, C1, up in cases:
    C = C0 + C1
    x0 = torch.randn(B, H, H, C0, device=DEV).to(torch.bfloat16)
    x1 = torch.randn(B, H, H, C1, device=DEV).to(torch.bfloat16) if C1 else None
    st0 = torch.zeros(B, C0, 2, device=DEV, dtype=torch.int64); ops.gn_stats(x0, B, H * H, C0, st0)
    st1 = None
    if C1:
        st1 = torch.zeros(B, C1, 2, device=DEV, dtype=torch.int64); ops.gn_stats(x1, B, H * H, C1, st1)
    ss = torch.randn(2 * C, device=DEV) * 0.1
    H2 = 2 * H if up else H // 2
    ya = torch.empty(B, H2, H2, C, device=DEV, dtype=torch.bfloat16)
    r0 = torch.empty(B, H2, H2, C0, device=DEV, dtype=torch.bfloat16)
    r1 = torch.empty(B, H2, H2, C1, device=DEV, dtype=torch.bfloat16) if C1 else None
    fn = lambda: ops.gn_fir(x0, C0, x1, C1, B, H, H, st0, st1, min(C // 4, 32), 1e-5, ss, True, up, ya, r0, r1)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = (x0.numel() + (x1.numel() if C1 else 0) + 2 * ya.numel()) * 2
    print(f"B={B} {'up  ' if up else 'down'} {H:3d}^2 C={C0}+{C1}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s "
          f"(EVC_GN_FIR_DOWN={os.environ.get('EVC_GN_FIR_DOWN', '1')})", flush=True)
