"""ORACLE (test infrastructure, never imported by the product path).

CPU/GPU-agnostic fp32 restatement, in plain torch.nn.functional calls, of the reference NCSN++ conditional
UNet forward.  Pinned against the live reference by tests/golden/make_golden.py (golden vectors in tests/golden/,
checked by tests/test_oracle.py); the reference itself ships no tests / golden vectors.

Each function cites the reference lines it restates (paths relative to /root/reference):
  ncsnpp_spec      models/better/ncsnpp_more.py:36-249   (module list construction)
  ncsnpp_forward   models/better/ncsnpp_more.py:251-392
  res_block        models/better/layerspp.py:553-624     (ResnetBlockBigGANppGN)
  act_norm         models/better/layerspp.py:486-549     (get_act_norm, AdaGN)
  attn_block       models/better/layerspp.py:207-249     (AttnBlockpp)
  nin              models/better/layers.py:535-544
  timestep_embedding  models/better/layers.py:504-518
  fir_up2 / fir_down2  models/better/up_or_down_sampling.py:196-258 + op/upfirdn2d.py:163-204
"""
import math

import torch
import torch.nn.functional as F


def gn_groups(ch):
    # layerspp.py:475-478 / :212-214
    g = min(ch // 4, 32)
    while ch % g != 0:
        g -= 1
    return g


def ncsnpp_spec(cfg):
    """List of module descriptors in all_modules order (ncsnpp_more.py:88-249)."""
    m = cfg.model
    d = cfg.data
    nf = m.ngf
    ch_mult = list(m.ch_mult)
    nres = m.num_res_blocks
    attn_res = list(m.attn_resolutions)
    nlev = len(ch_mult)
    all_res = [d.image_size // (2 ** i) for i in range(nlev)]
    n_frames = d.num_frames + d.num_frames_cond + getattr(d, "num_frames_future", 0)
    spec = [dict(kind="linear", cin=nf, cout=4 * nf), dict(kind="linear", cin=4 * nf, cout=4 * nf),
            dict(kind="conv3", cin=d.channels * n_frames, cout=nf)]
    hs_c = [nf]
    in_ch = nf
    for lvl in range(nlev):
        for _ in range(nres):
            out_ch = nf * ch_mult[lvl]
            spec.append(dict(kind="res", cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            if all_res[lvl] in attn_res:
                spec.append(dict(kind="attn", ch=in_ch))
            hs_c.append(in_ch)
        if lvl != nlev - 1:
            spec.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=True))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    spec.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=False))
    spec.append(dict(kind="attn", ch=in_ch))
    spec.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=False))
    for lvl in reversed(range(nlev)):
        for _ in range(nres + 1):
            out_ch = nf * ch_mult[lvl]
            spec.append(dict(kind="res", cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        if all_res[lvl] in attn_res:
            spec.append(dict(kind="attn", ch=in_ch))
        if lvl != 0:
            spec.append(dict(kind="res", cin=in_ch, cout=in_ch, up=True, down=False))
    assert not hs_c
    spec.append(dict(kind="actnorm_final", ch=in_ch))
    spec.append(dict(kind="conv3", cin=in_ch, cout=d.channels * d.num_frames))
    return spec


def timestep_embedding(t, dim, max_positions=10000):
    half = dim // 2
    e = math.log(max_positions) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -e)
    e = t.float()[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, (0, 1))
    return e


def _fir_kernel(device, dtype):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], device=device, dtype=dtype)
    k = torch.outer(k, k)
    return k / k.sum()


def fir_up2(x):
    """upsample_2d(x, [1,3,3,1], factor=2): zero-insert x2, pad (2,1), 4x4 FIR scaled by 4."""
    B, C, H, W = x.shape
    k = _fir_kernel(x.device, x.dtype) * 4.0
    z = x.new_zeros(B, C, H, 2, W, 2)
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(B * C, 1, 2 * H, 2 * W)
    z = F.pad(z, (2, 1, 2, 1))
    y = F.conv2d(z, torch.flip(k, [0, 1]).view(1, 1, 4, 4))
    return y.reshape(B, C, 2 * H, 2 * W)


def fir_down2(x):
    """downsample_2d(x, [1,3,3,1], factor=2): pad (1,1), 4x4 FIR, keep every 2nd sample."""
    B, C, H, W = x.shape
    k = _fir_kernel(x.device, x.dtype)
    z = F.pad(x.reshape(B * C, 1, H, W), (1, 1, 1, 1))
    y = F.conv2d(z, torch.flip(k, [0, 1]).view(1, 1, 4, 4))
    return y[:, :, ::2, ::2].reshape(B, C, H // 2, W // 2)


def act_norm(sd, pre, x, temb):
    C = x.shape[1]
    if temb is not None:
        emb = F.linear(F.silu(temb), sd[pre + ".Dense_0.weight"], sd[pre + ".Dense_0.bias"])[:, :, None, None]
        scale, shift = torch.chunk(emb, 2, dim=1)
        h = F.group_norm(x, gn_groups(C), None, None, eps=1e-5)
        h = h * (1 + scale) + shift
    else:
        h = F.group_norm(x, gn_groups(C), sd[pre + ".Norm_0.weight"], sd[pre + ".Norm_0.bias"], eps=1e-5)
    return F.silu(h)


def res_block(sd, pre, s, x, temb):
    h = act_norm(sd, pre + ".actnorm0", x, temb)
    if s["up"]:
        h, x = fir_up2(h), fir_up2(x)
    elif s["down"]:
        h, x = fir_down2(h), fir_down2(x)
    h = F.conv2d(h, sd[pre + ".Conv_0.weight"], sd[pre + ".Conv_0.bias"], padding=1)
    h = act_norm(sd, pre + ".actnorm1", h, temb)
    h = F.conv2d(h, sd[pre + ".Conv_1.weight"], sd[pre + ".Conv_1.bias"], padding=1)
    if s["cin"] != s["cout"] or s["up"] or s["down"]:
        x = F.conv2d(x, sd[pre + ".Conv_2.weight"], sd[pre + ".Conv_2.bias"])
    return (x + h) / math.sqrt(2.0)


def nin(sd, pre, x):
    y = torch.einsum("bchw,cd->bdhw", x, sd[pre + ".W"])
    return y + sd[pre + ".b"][None, :, None, None]


def attn_block(sd, pre, s, x, head_ch):
    B, C, H, W = x.shape
    heads = 1 if (head_ch == -1 or C < head_ch) else C // head_ch
    h = F.group_norm(x, gn_groups(C), sd[pre + ".GroupNorm_0.weight"], sd[pre + ".GroupNorm_0.bias"], eps=1e-6)
    q, k, v = nin(sd, pre + ".NIN_0", h), nin(sd, pre + ".NIN_1", h), nin(sd, pre + ".NIN_2", h)
    ch = C // heads
    q = q.reshape(B * heads, ch, H * W)
    k = k.reshape(B * heads, ch, H * W)
    v = v.reshape(B * heads, ch, H * W)
    w = torch.einsum("bcq,bck->bqk", q, k) * (int(ch) ** (-0.5))
    w = F.softmax(w, dim=-1)
    o = torch.einsum("bqk,bck->bcq", w, v).reshape(B, C, H, W)
    o = nin(sd, pre + ".NIN_3", o)
    return (x + o) / math.sqrt(2.0)


def ncsnpp_forward(sd, cfg, x, labels, cond=None, prefix="unet.all_modules.", taps=None):
    """eps = NCSNpp(x, labels, cond).  sd: reference state dict (fp32 tensors on x.device).

    taps: optional dict filled with named intermediate activations (for per-layer parity tests)."""
    spec = ncsnpp_spec(cfg)
    m = cfg.model
    nres = m.num_res_blocks
    nlev = len(m.ch_mult)
    attn_res = list(m.attn_resolutions)
    head_ch = getattr(m, "n_head_channels", -1)
    if cond is not None:
        x = torch.cat([x, cond.to(x.dtype) if cond.dtype != x.dtype else cond], dim=1)
    x = x.contiguous().float()
    P = lambda i: prefix + str(i)
    temb = timestep_embedding(labels, m.ngf)
    temb = F.linear(temb, sd[P(0) + ".weight"], sd[P(0) + ".bias"])
    temb = F.linear(F.silu(temb), sd[P(1) + ".weight"], sd[P(1) + ".bias"])
    i = 2

    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    hs = [rec("m2", F.conv2d(x, sd[P(i) + ".weight"], sd[P(i) + ".bias"], padding=1))]
    i += 1
    for lvl in range(nlev):
        for _ in range(nres):
            h = rec(f"m{i}", res_block(sd, P(i), spec[i], hs[-1], temb))
            i += 1
            if h.shape[-1] in attn_res:
                h = rec(f"m{i}", attn_block(sd, P(i), spec[i], h, head_ch))
                i += 1
            hs.append(h)
        if lvl != nlev - 1:
            hs.append(rec(f"m{i}", res_block(sd, P(i), spec[i], hs[-1], temb)))
            i += 1
    h = hs[-1]
    h = rec(f"m{i}", res_block(sd, P(i), spec[i], h, temb)); i += 1
    h = rec(f"m{i}", attn_block(sd, P(i), spec[i], h, head_ch)); i += 1
    h = rec(f"m{i}", res_block(sd, P(i), spec[i], h, temb)); i += 1
    for lvl in reversed(range(nlev)):
        for _ in range(nres + 1):
            h = rec(f"m{i}", res_block(sd, P(i), spec[i], torch.cat([h, hs.pop()], dim=1), temb))
            i += 1
        if h.shape[-1] in attn_res:
            h = rec(f"m{i}", attn_block(sd, P(i), spec[i], h, head_ch))
            i += 1
        if lvl != 0:
            h = rec(f"m{i}", res_block(sd, P(i), spec[i], h, temb))
            i += 1
    assert not hs
    h = rec(f"m{i}", act_norm(sd, P(i), h, None)); i += 1
    h = F.conv2d(h, sd[P(i) + ".weight"], sd[P(i) + ".bias"], padding=1); i += 1
    assert i == len(spec)
    return h


def ncsnpp_param_shapes(cfg, prefix="unet.all_modules."):
    """Ordered {state-dict key: shape} of every learnable tensor of UNetMore_DDPM(cfg) (446 keys minus the 4
    schedule buffers for configs/mine.yml); mirrors the constructors at ncsnpp_more.py:88-249,
    layerspp.py:207-224, 486-520, 553-594 and layers.py:535-540."""
    shapes = {}
    temb_dim = 4 * cfg.model.ngf
    for i, s in enumerate(ncsnpp_spec(cfg)):
        p = prefix + str(i)
        if s["kind"] == "linear":
            shapes[p + ".weight"] = (s["cout"], s["cin"])
            shapes[p + ".bias"] = (s["cout"],)
        elif s["kind"] == "conv3":
            shapes[p + ".weight"] = (s["cout"], s["cin"], 3, 3)
            shapes[p + ".bias"] = (s["cout"],)
        elif s["kind"] == "res":
            shapes[p + ".actnorm0.Dense_0.weight"] = (2 * s["cin"], temb_dim)
            shapes[p + ".actnorm0.Dense_0.bias"] = (2 * s["cin"],)
            shapes[p + ".Conv_0.weight"] = (s["cout"], s["cin"], 3, 3)
            shapes[p + ".Conv_0.bias"] = (s["cout"],)
            shapes[p + ".actnorm1.Dense_0.weight"] = (2 * s["cout"], temb_dim)
            shapes[p + ".actnorm1.Dense_0.bias"] = (2 * s["cout"],)
            shapes[p + ".Conv_1.weight"] = (s["cout"], s["cout"], 3, 3)
            shapes[p + ".Conv_1.bias"] = (s["cout"],)
            if s["cin"] != s["cout"] or s["up"] or s["down"]:
                shapes[p + ".Conv_2.weight"] = (s["cout"], s["cin"], 1, 1)
                shapes[p + ".Conv_2.bias"] = (s["cout"],)
        elif s["kind"] == "attn":
            c = s["ch"]
            shapes[p + ".GroupNorm_0.weight"] = (c,)
            shapes[p + ".GroupNorm_0.bias"] = (c,)
            for j in range(4):
                shapes[p + f".NIN_{j}.W"] = (c, c)
                shapes[p + f".NIN_{j}.b"] = (c,)
        elif s["kind"] == "actnorm_final":
            shapes[p + ".Norm_0.weight"] = (s["ch"],)
            shapes[p + ".Norm_0.bias"] = (s["ch"],)
    return shapes
