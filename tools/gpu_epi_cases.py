"""Three epilogue-bound launches of evc_gemm_kernel at B=46 (for an `ncu --set full --import-source on` capture):
NIN 384->768 @32^2 (rows), NIN 384->384 +residual +statistics @32^2, conv3x3 192->192 @128^2 with the fused
GroupNorm apply."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import torch
from evcdiff import ops

DEV, B = "cuda", 46


def mk(H, C, N, taps, resid=False, stats=False, gn=False):
    a = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    K = taps * C
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(B, H, H, N, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(B, H, H, N, device=DEV).to(torch.bfloat16) if resid else None
    st = torch.zeros(B, N, 2, device=DEV, dtype=torch.int64) if (stats or gn) else None
    gnd = dict(ss=torch.randn(2 * N, device=DEV) * 0.1, ticket=torch.zeros(B, dtype=torch.int32, device=DEV), eps=1e-5,
               groups=32, adagn=True) if gn else None
    plan = ops.GemmPlan([(a, taps)], w, out, 0, out_ld=N, bias=torch.zeros(N, device=DEV), resid=r, resid_ld=N if resid else 0,
                        stats=st, gn=gnd)

    def launch():
        if gn:
            gnd["ticket"].zero_()
            st.zero_()
        plan.launch()
    return launch


cases = [mk(32, 384, 768, 1), mk(32, 384, 384, 1, resid=True, stats=True), mk(128, 192, 192, 9, gn=True)]
for rep in range(3):
    for c in cases:
        c()
torch.cuda.synchronize()
print("ok")
