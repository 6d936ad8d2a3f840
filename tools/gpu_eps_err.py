"""Prints single-evaluation eps rel-L2 against the golden vectors for the small configs (test infrastructure)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import numpy as np
import torch
import common
from oracle import ncsnpp as O
from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
DEV = "cuda"
small = dict(np.load(os.path.join(ROOT, "tests", "golden", "ncsnpp_small.npz")))
for tag, cfgf, seed in (("tiny_act", common.tiny_config, 1), ("gpu64", common.gpu64_config, 4)):
    cfg = cfgf(device=DEV)
    net = UNetMore_DDPM(cfg)
    sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=seed, active=True)
    net.load_state_dict(sd, strict=False)
    net = net.to(DEV).eval()
    x = torch.from_numpy(small[f"{tag}_x"]).to(DEV)
    cond = torch.from_numpy(small[f"{tag}_cond"]).to(DEV)
    for lab in (0, 990):
        eps = net(x, torch.full((2,), lab, dtype=torch.long, device=DEV), cond=cond)
        print(tag, lab, f"{common.rel_l2(eps, torch.from_numpy(small[f'{tag}_eps_{lab}']).to(DEV)):.4e}", flush=True)

    if tag == "tiny_act" and os.environ.get("REPORT"):
        taps = {}
        lab = torch.full((2,), 0, dtype=torch.long, device=DEV)
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        O.ncsnpp_forward(sdd, cfg, x, lab, cond, taps=taps)
        net(x, lab, cond=cond)
        eng = net.engine(2, DEV)
        print([(n, a.t.shape[1], a.t.shape[3], f"{common.rel_l2(a.t.float().permute(0, 3, 1, 2), taps[n]):.2e}")
               for n, a in eng.taps.items() if n in taps and a.t.shape[0] == 2], flush=True)
