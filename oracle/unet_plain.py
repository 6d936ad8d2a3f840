"""ORACLE (test infrastructure, never imported by the product path).

fp32 torch restatement of the reference's plain DDPM UNet (models/unet.py), the BASELINE config-5 variant.  Pinned
against the live reference by tests/golden/make_golden.py (tests/golden/unet_plain.npz).

  unet_spec          models/unet.py:175-252  (UNet.__init__: downblocks / middleblocks / upblocks)
  unet_forward       models/unet.py:255-298
  res_block          models/unet.py:66-97    (GN(32,eps 1e-6)+Swish, conv, + Linear(temb), GN+Swish, conv, nin(x)+h)
  attn_block         models/unet.py:100-120  (single head, scale 1/sqrt(C), x + OUT(h))
  timestep_embedding models/unet.py:146-165  ; temb MLP ends WITH Swish (:247-252)
"""
import math

import torch
import torch.nn.functional as F


def unet_spec(cfg):
    """{'down': [...], 'mid': [...], 'up': [...]} module descriptors in ModuleList order."""
    ch = cfg.model.ngf
    d = cfg.data
    mode = getattr(cfg, "mode", "deep")
    mult = {"deepest": (1, 2, 2, 2, 4, 4), "deeper": (1, 2, 2, 4, 4), "deep": (1, 2, 2, 2)}[mode]
    ch_mult = [ch * n for n in mult]
    n_in = d.channels * (d.num_frames + d.num_frames_cond + getattr(d, "num_frames_future", 0))
    down = [dict(kind="conv3", cin=n_in, cout=ch, stride=1)]
    prev_ch = ch_mult[0]
    ch_size = [ch]
    for i, ich in enumerate(ch_mult):
        for firstarg in [prev_ch, ich]:
            down.append(dict(kind="res", cin=firstarg, cout=ich))
            ch_size.append(ich)
            if i == 1:
                down.append(dict(kind="attn", ch=ich))
        if i != len(ch_mult) - 1:
            down.append(dict(kind="conv3", cin=ich, cout=ich, stride=2))
            ch_size.append(ich)
        prev_ch = ich
    mid = [dict(kind="res", cin=ch_mult[-1], cout=ch_mult[-1]), dict(kind="attn", ch=ch_mult[-1]),
           dict(kind="res", cin=ch_mult[-1], cout=ch_mult[-1])]
    up = []
    prev_ich = ch_mult[-1]
    for i, ich in reversed(list(enumerate(ch_mult))):
        for _ in range(3):
            up.append(dict(kind="res", cin=prev_ich + ch_size.pop(), cout=ich))
            if i == 1:
                up.append(dict(kind="attn", ch=ich))
            prev_ich = ich
        if i != 0:
            up.append(dict(kind="upsample", ch=ich))
    n_out = d.channels * d.num_frames
    return dict(down=down, mid=mid, up=up, ch=ch, n_out=n_out)


def unet_param_shapes(cfg, prefix="unet."):
    spec = unet_spec(cfg)
    ch = spec["ch"]
    shapes = {}

    def add(p, s):
        if s["kind"] == "conv3":
            shapes[p + ".weight"] = (s["cout"], s["cin"], 3, 3)
            shapes[p + ".bias"] = (s["cout"],)
        elif s["kind"] == "res":
            shapes[p + ".normalize0.weight"] = (s["cin"],)
            shapes[p + ".normalize0.bias"] = (s["cin"],)
            shapes[p + ".conv0.weight"] = (s["cout"], s["cin"], 3, 3)
            shapes[p + ".conv0.bias"] = (s["cout"],)
            shapes[p + ".dense.weight"] = (s["cout"], 4 * ch)
            shapes[p + ".dense.bias"] = (s["cout"],)
            shapes[p + ".normalize1.weight"] = (s["cout"],)
            shapes[p + ".normalize1.bias"] = (s["cout"],)
            shapes[p + ".conv1.weight"] = (s["cout"], s["cout"], 3, 3)
            shapes[p + ".conv1.bias"] = (s["cout"],)
            if s["cin"] != s["cout"]:
                shapes[p + ".nin.weights"] = (s["cout"], s["cin"])
                shapes[p + ".nin.bias"] = (s["cout"],)
        elif s["kind"] == "attn":
            for n in ("Q", "K", "V", "OUT"):
                shapes[p + f".{n}.weights"] = (s["ch"], s["ch"])
                shapes[p + f".{n}.bias"] = (s["ch"],)
            shapes[p + ".normalize.weight"] = (s["ch"],)
            shapes[p + ".normalize.bias"] = (s["ch"],)
        elif s["kind"] == "upsample":
            shapes[p + ".conv.weight"] = (s["ch"], s["ch"], 3, 3)
            shapes[p + ".conv.bias"] = (s["ch"],)

    for name in ("down", "mid", "up"):
        for j, s in enumerate(spec[name]):
            add(f"{prefix}{ {'down': 'downblocks', 'mid': 'middleblocks', 'up': 'upblocks'}[name] }.{j}", s)
    shapes[prefix + "normalize.weight"] = (ch,)
    shapes[prefix + "normalize.bias"] = (ch,)
    shapes[prefix + "out.weight"] = (spec["n_out"], ch, 3, 3)
    shapes[prefix + "out.bias"] = (spec["n_out"],)
    shapes[prefix + "temb_dense.0.weight"] = (4 * ch, ch)
    shapes[prefix + "temb_dense.0.bias"] = (4 * ch,)
    shapes[prefix + "temb_dense.2.weight"] = (4 * ch, 4 * ch)
    shapes[prefix + "temb_dense.2.bias"] = (4 * ch,)
    return shapes


def swish(x):
    return x * torch.sigmoid(x)


def timestep_embedding(t, dim):
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float, device=t.device) * -e)
    e = t.float().unsqueeze(1) * e.unsqueeze(0)
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


def res_block(sd, p, s, x, temb):
    h = swish(F.group_norm(x, 32, sd[p + ".normalize0.weight"], sd[p + ".normalize0.bias"], eps=1e-6))
    h = F.conv2d(h, sd[p + ".conv0.weight"], sd[p + ".conv0.bias"], padding=1)
    h = h + F.linear(temb, sd[p + ".dense.weight"], sd[p + ".dense.bias"]).unsqueeze(-1).unsqueeze(-1)
    h = swish(F.group_norm(h, 32, sd[p + ".normalize1.weight"], sd[p + ".normalize1.bias"], eps=1e-6))
    h = F.conv2d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    if s["cin"] != s["cout"]:
        x = torch.einsum("oc,bchw->bohw", sd[p + ".nin.weights"], x) + sd[p + ".nin.bias"][None, :, None, None]
    return x + h


def attn_block(sd, p, s, x):
    B, C, H, W = x.shape
    h = F.group_norm(x, 32, sd[p + ".normalize.weight"], sd[p + ".normalize.bias"], eps=1e-6)
    nin = lambda n, t: torch.einsum("oc,bchw->bohw", sd[p + f".{n}.weights"], t) + sd[p + f".{n}.bias"][None, :, None, None]
    q, k, v = nin("Q", h).flatten(2), nin("K", h).flatten(2), nin("V", h).flatten(2)
    w = torch.einsum("bcq,bck->bqk", q, k) * (1 / math.sqrt(C))
    w = F.softmax(w, dim=-1)
    o = torch.einsum("bqk,bck->bcq", w, v).reshape(B, C, H, W)
    return x + nin("OUT", o)


def unet_forward(sd, cfg, x, y, cond=None, prefix="unet.", taps=None):
    spec = unet_spec(cfg)
    ch = spec["ch"]
    temb = timestep_embedding(y, ch)
    temb = swish(F.linear(temb, sd[prefix + "temb_dense.0.weight"], sd[prefix + "temb_dense.0.bias"]))
    temb = swish(F.linear(temb, sd[prefix + "temb_dense.2.weight"], sd[prefix + "temb_dense.2.bias"]))
    if cond is not None:
        x = torch.cat([x, cond.to(x.dtype)], dim=1)
    x = x.float()
    if not cfg.data.logit_transform and not cfg.data.rescaled:
        x = 2 * x - 1.0

    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    hs = []
    for j, s in enumerate(spec["down"]):
        p = f"{prefix}downblocks.{j}"
        if s["kind"] == "res":
            x = res_block(sd, p, s, x, temb)
        elif s["kind"] == "attn":
            x = attn_block(sd, p, s, x)
            hs.pop()
        else:
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=s["stride"], padding=1)
        rec(f"d{j}", x)
        hs.append(x)
    for j, s in enumerate(spec["mid"]):
        p = f"{prefix}middleblocks.{j}"
        x = res_block(sd, p, s, x, temb) if s["kind"] == "res" else attn_block(sd, p, s, x)
        rec(f"m{j}", x)
    for j, s in enumerate(spec["up"]):
        p = f"{prefix}upblocks.{j}"
        if s["kind"] == "res":
            x = res_block(sd, p, s, torch.cat((x, hs.pop()), dim=1), temb)
        elif s["kind"] == "attn":
            x = attn_block(sd, p, s, x)
        else:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        rec(f"u{j}", x)
    x = swish(F.group_norm(x, 32, sd[prefix + "normalize.weight"], sd[prefix + "normalize.bias"], eps=1e-6))
    return F.conv2d(x, sd[prefix + "out.weight"], sd[prefix + "out.bias"], padding=1)
