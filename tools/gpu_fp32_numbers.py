"""Accuracy of the two arithmetic modes against the fp32 oracle (run on a GPU box; prints a small table).
Test infrastructure: imports oracle/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import numpy as np
import torch
import common
from oracle import ncsnpp as O
from oracle import samplers as S
from evcdiff import models as M
from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
cfg = common.gpu64_config(device=DEV)
net = UNetMore_DDPM(cfg)
sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=4, active=True)
net.load_state_dict(sd, strict=False)
net = net.to(DEV).eval()
sd = {k: v.to(DEV) for k, v in sd.items()}
g = torch.Generator(device=DEV).manual_seed(15)
x_T = torch.randn(2, 15, 32, 32, device=DEV, generator=g)
cond = torch.rand(2, 6, 32, 32, device=DEV, generator=g, dtype=torch.float64) * 2 - 1
sched = S.schedule(cfg, DEV)
model = lambda x, y: O.ncsnpp_forward(sd, cfg, x, y, cond)


def psnr(a, b):
    a = ((a + 1) / 2).clamp(0, 1).double(); b = ((b + 1) / 2).clamp(0, 1).double()
    return float(10 * torch.log10(1.0 / ((a - b) ** 2).mean()))


gt_frames = torch.rand(2, 15, 32, 32, device=DEV, generator=g) * 2 - 1
for prec in ("bf16", "fp32"):
    for lab in (0, 500, 990):
        y = torch.full((2,), lab, dtype=torch.long, device=DEV)
        net.precision = prec
        e = net(x_T, y, cond=cond)
        print(f"eps {prec} label {lab}: rel-L2 {common.rel_l2(e, model(x_T, y)):.3e}", flush=True)
    for kind, steps in (("ddpm", 100), ("ddim", 10), ("ddim", 25), ("ddim", 50), ("ddim", 100)):
        gt = torch.Generator().manual_seed(78)
        tape = [torch.randn(x_T.shape, generator=gt).to(DEV) for _ in range(steps)]
        trace = []
        if kind == "ddpm":
            imgs = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=False, subsample_steps=steps, noise=tape, precision=prec)
            S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i], subsample_steps=steps, trace=trace)
        else:
            imgs = M.ddim_sampler(x_T.clone(), net, cond=cond, final_only=False, subsample_steps=steps, precision=prec)
            S.ddim_sampler(x_T.clone(), model, sched, subsample_steps=steps, trace=trace)
        n = min(len(imgs), len(trace))
        errs = [common.rel_l2(imgs[i].to(DEV), trace[i][0]) for i in range(n)]
        dp = abs(psnr(imgs[-1].to(DEV), gt_frames) - psnr(trace[n - 1][0], gt_frames))
        print(f"{kind}-{steps} {prec}: max per-step x_t rel-L2 {max(errs):.3e} (final {errs[-1]:.3e}), |dPSNR| {dp:.4f} dB", flush=True)
