"""Launch plan of the plain DDPM UNet (reference models/unet.py:175-298) on the same kernels as the NCSN++ engine.

Differences from NCSN++ that shape the plan:
  * additive time conditioning: `h += dense(temb)` after conv0 (unet.py:87-88) -> a per-label bias table
    (conv0.bias + dense(temb_label)) selected at launch (`bias_override`), no extra pass;
  * down-sampling = 3x3 convolution with stride 2 (unet.py:219) -> TMA traversal stride 2 on the A operand;
  * up-sampling = nearest x2 + 3x3 convolution (unet.py:123-131) -> evc_nearest_up2 then the implicit GEMM;
  * residual `nin(x) + h` without 1/sqrt(2) (unet.py:97); single-head attention with scale 1/sqrt(C) (unet.py:114);
  * GroupNorm(32, eps=1e-6, affine) + Swish everywhere (unet.py:44-46).
"""
import math

import torch

from . import ops
from ._lib import EVC_OUT_BF16_ROWS, EVC_OUT_F32_T, EvcError
from .engine import CIN_PAD, Act, EngineBase, cond_offset, pack_conv3, pack_conv_in


class PlainUNetEngine(EngineBase):
    def __init__(self, net, B, device, precision="bf16"):
        cfg = net.config
        H = cfg.data.image_size
        super().__init__(device, B, H, precision)
        self.net, self.cfg = net, cfg
        self.fixed_groups = 32
        d = cfg.data
        self.ch = cfg.model.ngf
        self.nf = self.ch
        self.c_x = d.channels * d.num_frames
        self.c_cond = d.channels * (d.num_frames_cond + getattr(d, "num_frames_future", 0))
        if self.ch % 32 != 0:
            raise EvcError("models/unet.py uses GroupNorm(32): model.ngf must be a multiple of 32")
        self.c_cond_off = cond_offset(self.c_x)
        if self.c_cond_off + self.c_cond > CIN_PAD:
            raise EvcError("more than 64 input channels is not supported")
        self.sd = {k: v.detach() for k, v in net.state_dict().items()}
        self._build()
        self.finalize()

    def f32(self, key):
        t = self.sd[key].to(self.device, torch.float32).contiguous()
        self._keep.append(t)
        return t

    def _gn_ss(self, prefix):
        ss = torch.cat([self.sd[prefix + ".weight"].float(), self.sd[prefix + ".bias"].float()]).to(self.device)
        ss = ss.contiguous()
        self._keep.append(ss)
        return ss

    def _build(self):
        from .models.unet import unet_spec
        sd, dev, B, H = self.sd, self.device, self.B, self.H
        spec = unet_spec(self.cfg)
        self.spec = spec
        half = self.ch // 2
        e = math.log(10000) / (half - 1)
        self.freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -e).to(dev)
        self.t_w0, self.t_b0 = self.f32("temb_dense.0.weight"), self.f32("temb_dense.0.bias")
        self.t_w2, self.t_b2 = self.f32("temb_dense.2.weight"), self.f32("temb_dense.2.bias")
        # per-label conv0 bias table: conv0.bias + dense(temb)
        dw, db, self.bias_off = [], [], {}
        off = 0
        for grp, name in (("down", "downblocks"), ("mid", "middleblocks"), ("up", "upblocks")):
            for j, s in enumerate(spec[grp]):
                if s["kind"] == "res":
                    p = f"{name}.{j}"
                    dw.append(sd[p + ".dense.weight"].float())
                    db.append(sd[p + ".dense.bias"].float() + sd[p + ".conv0.bias"].float())
                    self.bias_off[p] = (off, s["cout"])
                    off += s["cout"]
        self.dense_w = torch.cat(dw, 0).to(dev).contiguous()
        self.dense_b = torch.cat(db, 0).to(dev).contiguous()
        self.bias_total = off
        self.bias_table = None

        self.xin = torch.zeros((B, H, H, CIN_PAD), dtype=torch.bfloat16, device=dev)
        self.xin_lo = torch.zeros_like(self.xin) if self.split else None
        self.eps = torch.zeros((B, spec["n_out"], H, H), dtype=torch.float32, device=dev)
        x = Act(self.xin, self.xin_lo)
        hs = []
        for j, s in enumerate(spec["down"]):
            p = f"downblocks.{j}"
            if s["kind"] == "res":
                x = self.res_block(p, s, x, None)
            elif s["kind"] == "attn":
                x = self.attn_block(p, s, x)
                hs.pop()
            else:
                st = s["stride"]
                out = self.new_act(x.H // st, x.W // st, s["cout"], scratch=False)
                w = pack_conv_in(sd[p + ".weight"].to(dev), self.c_x, CIN_PAD) if j == 0 else pack_conv3(sd[p + ".weight"].to(dev))
                self.gemm([(x, 9)], w, out.t, EVC_OUT_BF16_ROWS, s["cout"], bias=self.f32(p + ".bias"), stats_of=out,
                          stride=st)
                x = out
            self.taps[f"d{j}"] = x
            hs.append(x)
        for j, s in enumerate(spec["mid"]):
            p = f"middleblocks.{j}"
            x = self.res_block(p, s, x, None) if s["kind"] == "res" else self.attn_block(p, s, x)
            self.taps[f"m{j}"] = x
        for j, s in enumerate(spec["up"]):
            p = f"upblocks.{j}"
            if s["kind"] == "res":
                x = self.res_block(p, s, x, hs.pop())
            elif s["kind"] == "attn":
                x = self.attn_block(p, s, x)
            else:
                up = self.new_act(x.H * 2, x.W * 2, x.C)
                self._op(lambda li, a=x, o=up: ops.nearest_up2(a.t, o.t, a.B, a.H, a.W, a.C), "fir",
                         dict(bytes=x.t.numel() * 2 * 5))
                if self.split:  # nearest-neighbour is a pure copy: replicate the residual plane too
                    self._op(lambda li, a=x, o=up: ops.nearest_up2(a.lo, o.lo, a.B, a.H, a.W, a.C), "fir",
                             dict(bytes=x.t.numel() * 2 * 5))
                out = self.new_act(up.H, up.W, s["ch"], scratch=False)
                self.gemm([(up, 9)], pack_conv3(sd[p + ".conv.weight"].to(dev)), out.t, EVC_OUT_BF16_ROWS, s["ch"],
                          bias=self.f32(p + ".conv.bias"), stats_of=out)
                self.release(up)
                x = out
            self.taps[f"u{j}"] = x
        assert not hs
        ss = self._gn_ss("normalize")
        hn = self.new_act(x.H, x.W, x.C)
        self.gn_apply(x, None, lambda li: ss, 1e-6, False, True, hn)
        self.final_conv([(hn, 9)], pack_conv3(sd["out.weight"].to(dev)), self.f32("out.bias"), spec["n_out"], H)

    def res_block(self, p, s, xa, xb):
        """ResnetBlock (unet.py:66-97) on the virtual concat [xa | xb]."""
        sd, dev = self.sd, self.device
        cin, cout = s["cin"], s["cout"]
        assert cin == xa.C + (xb.C if xb is not None else 0)
        ss0 = self._gn_ss(p + ".normalize0")
        h = self.new_act(xa.H, xa.W, cin)
        self.gn_apply(xa, xb, lambda li: ss0, 1e-6, False, True, h)
        off, n = self.bias_off[p]
        ss1 = self._gn_ss(p + ".normalize1")
        a1 = self.new_act(xa.H, xa.W, cout)
        bias_fn = lambda li, off=off, n=n: self.bias_table[li, off:off + n]
        if self.gn_fusable(xa.B, xa.H, xa.W, [(cin, 9)], cout):
            # conv0 (+ per-label temb bias) -> normalize1 -> Swish in one launch
            c0 = None
            self.gemm([(h, 9)], pack_conv3(sd[p + ".conv0.weight"].to(dev)), a1.t, EVC_OUT_BF16_ROWS, cout,
                      bias=self.dense_b[off:off + n], bias_fn=bias_fn,
                      gn=dict(ss_fn=lambda li: ss1, eps=1e-6, adagn=False, groups=self.fixed_groups))
            self.release(h)
        else:
            c0 = self.new_act(xa.H, xa.W, cout)
            self.gemm([(h, 9)], pack_conv3(sd[p + ".conv0.weight"].to(dev)), c0.t, EVC_OUT_BF16_ROWS, cout,
                      bias=self.dense_b[off:off + n], bias_fn=bias_fn, stats_of=c0)
            self.release(h)
            self.gn_apply(c0, None, lambda li: ss1, 1e-6, False, True, a1)
        out = self.new_act(xa.H, xa.W, cout, scratch=False)
        w1 = pack_conv3(sd[p + ".conv1.weight"].to(dev))
        b1 = sd[p + ".conv1.bias"].float()
        xs = [xa] + ([xb] if xb is not None else [])
        if (p + ".nin.weights") in sd:
            w = torch.cat([w1, sd[p + ".nin.weights"].to(dev).float()], dim=1).contiguous()
            bias = (b1 + sd[p + ".nin.bias"].float()).to(dev).contiguous()
            self._keep.append(bias)
            self.gemm([(a1, 9)] + [(x, 1) for x in xs], w, out.t, EVC_OUT_BF16_ROWS, cout, bias=bias, stats_of=out)
        else:
            assert len(xs) == 1 and xs[0].C == cout
            bias = b1.to(dev).contiguous()
            self._keep.append(bias)
            self.gemm([(a1, 9)], w1, out.t, EVC_OUT_BF16_ROWS, cout, bias=bias, resid=xs[0], stats_of=out)
        self.release(c0, a1)
        return out

    def attn_block(self, p, s, x):
        """AttnBlock (unet.py:100-120): single head, scale 1/sqrt(C), x + OUT(h)."""
        sd, dev = self.sd, self.device
        ws = [sd[p + f".{n}.weights"].to(dev).float().contiguous() for n in ("Q", "K", "V", "OUT")]
        bs = [sd[p + f".{n}.bias"].float().to(dev).contiguous() for n in ("Q", "K", "V", "OUT")]
        return self.attn_core(x, self._gn_ss(p + ".normalize"), 1e-6, ws, bs, 1, 1.0)

    def set_labels(self, labels):
        """temb = Swish(Linear(Swish(Linear(emb(y))))) (unet.py:247-259) and every block's dense(temb) + conv0.bias,
        once per distinct label."""
        lab = torch.tensor([float(v) for v in labels], dtype=torch.float32, device=self.device)
        L = lab.numel()
        emb = torch.empty((L, self.ch), dtype=torch.float32, device=self.device)
        ops.timestep_embedding(lab, self.freqs, self.ch, emb)
        t1 = torch.empty((L, 4 * self.ch), dtype=torch.float32, device=self.device)
        ops.linear_f32(emb, self.t_w0, self.t_b0, t1, act_out=True)
        t2 = torch.empty_like(t1)
        ops.linear_f32(t1, self.t_w2, self.t_b2, t2, act_out=True)
        if self.bias_table is None or self.bias_table.shape[0] < L:
            self.bias_table = torch.empty((L, self.bias_total), dtype=torch.float32, device=self.device)
        ops.linear_f32(t2, self.dense_w, self.dense_b, self.bias_table[:L])
        self.labels = [float(v) for v in labels]
        return L

    @property
    def ss_table(self):  # the sampler loop keys its graphs on the label-table pointer
        return self.bias_table

    def load_input(self, x, cond):
        ops.pack_nchw(x.contiguous(), self.xin, 0, dst_lo=self.xin_lo)
        if cond is not None:
            ops.pack_nchw(cond.contiguous(), self.xin, self.c_cond_off, dst_lo=self.xin_lo)

    def refresh_x(self, x):
        if self.split:
            ops.pack_nchw(x, self.xin, 0, dst_lo=self.xin_lo)
