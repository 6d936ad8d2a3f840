"""Execution plans for one UNet evaluation on the CUDA kernels of libevcdiff.so.

An engine is built once per (model, batch size): it repacks the reference-layout state dict into bf16 K-major
GEMM operands, allocates every activation buffer (NHWC bf16) up front, encodes the TMA descriptors and records
the launch sequence.  `forward(label_index)` then only enqueues kernels -- no allocation, no host sync -- so a
whole sampling loop can be captured in a single CUDA graph.

Layout decisions (DESIGN.md has the full table):
  activations   (B, H, W, C) bf16, C contiguous; UNet input padded to 64 channels [x_t(15) | cond(6) | 0]
  weights       conv3x3 (Cout,Cin,3,3) -> (Cout, 9*Cin) tap-major/channel-minor bf16; 1x1 / NIN -> (Cout, Cin)
  residual 1x1  folded into the block's second conv as extra K segments (bias = b1 + b2)
  GroupNorm     per-(sample, channel) [sum, sumsq] fp32, memoised per tensor (a skip tensor is reduced once)
  AdaGN         scale/shift rows of a per-label table evaluated before the loop (labels are batch-uniform)
"""
import math
import os

import torch

from . import ops
from ._lib import EVC_OUT_BF16_ROWS, EVC_OUT_BF16_T, EVC_OUT_F32_ROWS, EVC_OUT_F32_T, EvcError

CIN_PAD = 64
RSQRT2 = 1.0 / math.sqrt(2.0)


def gn_groups(ch):
    g = min(ch // 4, 32)
    while ch % g != 0:
        g -= 1
    return g


class Act:
    """NHWC bf16 activation with lazily computed GroupNorm statistics."""
    __slots__ = ("t", "lo", "B", "H", "W", "C", "stats")

    def __init__(self, t, lo=None):
        self.t = t
        self.lo = lo  # split-precision residual plane (value = t + lo), None in bf16 mode
        self.B, self.H, self.W, self.C = t.shape
        self.stats = None


class _StatSlot:
    """A slice of the statistics arena that belongs to no activation (fused GroupNorm apply: raw conv statistics and
    the per-sample tile tickets)."""
    __slots__ = ("B", "C", "stats")

    def __init__(self, B, C):
        self.B, self.C, self.stats = B, C, None


class Pool:
    """Static scratch reuse: buffers are handed out at plan-build time; stream order makes reuse safe."""

    def __init__(self, device):
        self.device = device
        self.free = {}
        self.bytes = 0

    def get(self, shape, dtype=torch.bfloat16):
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        t = torch.zeros(shape, dtype=dtype, device=self.device)
        self.bytes += t.numel() * t.element_size()
        return t

    def put(self, t):
        self.free.setdefault((tuple(t.shape), t.dtype), []).append(t)


def pack_conv3(w, cin_pad=None):
    """(Cout, Cin, 3, 3) fp32 -> (Cout, 9*Cin') fp32, K = tap-major (ky,kx), channel-minor (EngineBase.gemm rounds it to
    bf16, or to a (hi, lo) bf16 pair in split-precision mode)."""
    co, ci, kh, kw = w.shape
    w = w.float().permute(0, 2, 3, 1)  # co, ky, kx, ci
    if cin_pad is not None and cin_pad != ci:
        w = torch.nn.functional.pad(w, (0, cin_pad - ci))
    return w.reshape(co, -1).contiguous()


def pack_conv_in(w, c_x, cin_pad):
    """First conv of the UNet: (Cout, c_x + c_cond, 3, 3) -> packed K with the input-row layout of the sampler-update
    kernels: [x_t (c_x) | zero pad to a multiple of 8 | cond (c_cond) | zero pad to cin_pad] (evc_sampler_update writes
    x_t as whole 16-byte chunks, so the conditioning frames of the reference's torch.cat start at roundup8(c_x))."""
    co, ci = w.shape[:2]
    off = cond_offset(c_x)
    w2 = torch.zeros((co, cin_pad, 3, 3), dtype=torch.float32, device=w.device)
    w2[:, :c_x] = w[:, :c_x].float()
    w2[:, off:off + ci - c_x] = w[:, c_x:].float()
    return pack_conv3(w2)


def cond_offset(c_x):
    return (c_x + 7) // 8 * 8


def split_bf16(w):
    """fp32 -> (hi, lo) bf16 planes with hi + lo = w to 2^-17."""
    hi = w.to(torch.bfloat16)
    lo = (w.float() - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


class EngineBase:
    def __init__(self, device, B, H, precision="bf16"):
        if torch.device(device).type != "cuda":
            raise EvcError("evcdiff engines run on CUDA devices only (no CPU fallback)")
        if precision not in ("bf16", "fp32"):
            raise EvcError("precision must be 'bf16' (default) or 'fp32' (split-bf16 x3, fp32-tolerance mode)")
        self.precision = precision
        self.split = precision == "fp32"
        self.device = torch.device(device)
        self.B, self.H = B, H
        self.pool = Pool(self.device)
        self.ops = []  # list of callables(label_idx)
        self.op_info = []  # (kind, meta) per op, for profiling
        self.stats_slices = []  # (offset, numel) in the stats arena
        self.stats_total = 0
        self.flops = 0.0
        self.n_launch = 0
        self.taps = {}
        self._keep = []
        self._stat_acts = []
        self._deferred = []
        self.fixed_groups = None  # models/unet.py: always 32 groups; NCSN++: min(C//4, 32)
        self.fused_attention = os.environ.get("EVC_FUSED_ATTENTION", "1") != "0"
        self.fused_qkv = os.environ.get("EVC_FUSED_QKV", "1") != "0"  # one q|k|v projection, V rows as an MN-major operand
        self.ws_bytes = 256
        # split-K partial tiles (fp32) of the launches with few M tiles; one buffer, the launches are stream-ordered
        self.sk_ws = None if self.split else torch.empty(ops.SPLIT_K_WS_BYTES, dtype=torch.uint8, device=self.device)

    # ----------------------------------------------------------------- recording helpers
    def _op(self, fn, kind="other", meta=None):
        self.ops.append(fn)
        self.op_info.append((kind, meta))
        self.n_launch += 1

    def new_act(self, H, W, C, scratch=True):
        def one():
            return self.pool.get((self.B, H, W, C)) if scratch else torch.zeros((self.B, H, W, C), dtype=torch.bfloat16,
                                                                                device=self.device)
        return Act(one(), one() if self.split else None)

    def release(self, *acts):
        for a in acts:
            if a is not None:
                self.pool.put(a.t)
                if a.lo is not None:
                    self.pool.put(a.lo)

    def alloc_stats(self, a):
        n = a.B * a.C * 2
        a.stats = ("slice", self.stats_total, n)
        self.stats_total += n
        self._stat_acts.append(a)

    @staticmethod
    def can_fuse_stats(H, W):
        tw = min(W, 128)
        th = max(1, min(H, 128 // tw))
        return (tw * th) % 32 == 0

    def gn_fusable(self, B, H, W, cin_segs, cout):
        """Can conv -> GroupNorm -> SiLU run as one GEMM launch?  (whole 128-row tiles inside one sample, N tile of
        whole 32-column chunks; not in split-precision mode)"""
        if self.split or os.environ.get("EVC_GEMM_FUSE_GN", "1") == "0":
            return False
        tw = min(W, 128)
        th = max(1, min(H, 128 // tw))
        if tw * th != 128 or W % tw or H % th:
            return False
        kblocks = sum(taps * (-(-c // 64)) for c, taps in cin_segs)
        mt = ops.m_tiles(B, H, W, False)
        if ops.pick_tile(cout, mt, kblocks)[1] > 1:
            return False  # few tiles (small batch / low resolution): split-K + a separate gn_apply launch is faster
        bn = ops.pick_bn(cout, mt, kblocks)
        if bn % 32 != 0 or bn > 192:  # two bn x 256 B tile slots + >= 3 pipeline stages must fit in 227 KB
            return False
        # mirror of evc_gemm_plan_create: a CTA must never own a third tile of a sample whose statistics it waits for
        return ops.gn_fuse_fits((W // tw) * (H // th), mt, cout // bn)

    def ensure_stats(self, a):
        if a.stats is not None:
            return
        self.alloc_stats(a)
        act = a

        self.ws_bytes = max(self.ws_bytes, ops.gn_stats_workspace_bytes(act.B, act.H * act.W, act.C))

        def run(_):
            ops.gn_stats(act.t, act.B, act.H * act.W, act.C, act.stats, workspace=self.workspace, x_lo=act.lo)
        self._op(run, "gn_stats", dict(bytes=act.t.numel() * 2))

    def _stats_view(self, a):
        return a.stats

    def gn_apply(self, xa, xb, ss_fn, eps, adagn, silu, out):
        """out = act(GN([xa|xb]) * gamma' + beta'); ss_fn(label_idx) -> fp32 tensor slice of 2*C floats."""
        self.ensure_stats(xa)
        if xb is not None:
            self.ensure_stats(xb)
        C = xa.C + (xb.C if xb is not None else 0)
        groups = self.fixed_groups or gn_groups(C)

        def run(li):
            ops.gn_apply(xa.t, xa.C, xb.t if xb is not None else None, xb.C if xb is not None else 0, xa.B,
                         xa.H * xa.W, self._stats_view(xa), self._stats_view(xb) if xb is not None else None,
                         groups, eps, ss_fn(li), adagn, silu, out.t, x0_lo=xa.lo,
                         x1_lo=xb.lo if xb is not None else None, y_lo=out.lo)
        self._op(run, "gn_apply", dict(bytes=out.t.numel() * 4))

    def gemm(self, segs, w, out_t, out_mode, out_ld, out_bs=0, bias=None, resid=None, alpha=1.0, bias_fn=None,
             stats_of=None, stride=1, w_lo=None, out_lo=None, segs_lo=None, gn=None):
        """One implicit-GEMM launch.  segs: [(Act | tensor (B,H,W,C), taps)].  w: fp32 (N,K) weights (rounded here to bf16,
        or split into a (hi, lo) pair in split-precision mode) or a bf16 per-sample operand (B,N,K) (+ w_lo).
        stats_of: the Act being written; when given (and the tile geometry allows it) its GroupNorm statistics are
        accumulated by the GEMM epilogue and no separate gn_stats launch is recorded for it."""
        if w.dtype != torch.bfloat16:  # model weights
            if self.split:
                w, w_lo = split_bf16(w.to(self.device))
            else:
                w = w.to(self.device).to(torch.bfloat16).contiguous()
        seg_t = [(a.t if isinstance(a, Act) else a, taps) for a, taps in segs]
        if self.split:
            if segs_lo is None:
                segs_lo = [a.lo for a, _ in segs]
            assert all(x is not None for x in segs_lo) and w_lo is not None, "split precision: missing residual planes"
            if isinstance(resid, Act):
                resid_lo = resid.lo
            else:
                resid_lo = None
            if out_lo is None and stats_of is not None:
                out_lo = stats_of.lo
        else:
            segs_lo = w_lo = out_lo = resid_lo = None
        stats_t = None
        if gn is not None:
            # fused GroupNorm apply (see evc_gemm_desc.gn_ss): `out_t` receives act(GN(conv)); the raw statistics and
            # the tile tickets live in anonymous arena slots.  gn = dict(ss_fn, eps, adagn, groups)
            assert stats_of is None and not self.split
            raw = _StatSlot(out_t.shape[0], w.shape[-2])
            tick = _StatSlot(out_t.shape[0], 1)
            self.alloc_stats(raw)
            self.alloc_stats(tick)
            stats_of = raw
            stats_t = ("deferred", raw)
            gn = dict(gn, ticket=tick)
        elif stats_of is not None and out_mode == EVC_OUT_BF16_ROWS and self.can_fuse_stats(stats_of.H, stats_of.W):
            self.alloc_stats(stats_of)
            stats_t = ("deferred", stats_of)
        plan_args = dict(out_bs=out_bs, bias=bias, resid=resid.t if isinstance(resid, Act) else resid,
                         resid_ld=(resid.C if isinstance(resid, Act) else 0), alpha=alpha, stride=stride,
                         segs_lo=segs_lo, w_lo=w_lo, out_lo=out_lo, resid_lo=resid_lo,
                         split_k="auto", sk_ws=self.sk_ws)
        a0 = seg_t[0][0]
        shp = tuple(a0.shape)
        M = shp[0] * (shp[1] // stride) * (shp[2] // stride)
        flops = 2.0 * M * w.shape[-2] * w.shape[-1]
        self.flops += flops
        meta = dict(flops=flops, M=M, N=w.shape[-2], K=w.shape[-1], taps=[t for _, t in segs], hw=shp[2] // stride)
        if stats_t is not None:
            # the arena does not exist yet: create the plan in finalize()
            self._deferred.append((len(self.ops), seg_t, w, out_t, out_mode, out_ld, plan_args, stats_of, bias_fn, gn))
            self._op(None, "gemm", meta)
            return None
        plan = ops.GemmPlan(seg_t, w, out_t, out_mode, out_ld, **plan_args)
        if bias_fn is None:
            self._op(lambda li: plan.launch(), "gemm", meta)
        else:
            self._op(lambda li: plan.launch(bias_fn(li)), "gemm", meta)
        return plan

    def final_conv(self, segs, w, bias, c_out, H):
        """Last conv: NHWC bf16 -> eps (B, c_out, H, H) fp32 NCHW, written by the epilogue directly (transposed fp32
        store).  forward(eps_out=...) redirects it, e.g. into a slot of the F-PNDM eps history ring."""
        self._eps_out = None
        plan = self.gemm(segs, w, self.eps, EVC_OUT_F32_T, H * H, out_bs=c_out * H * H, bias=bias)
        self.ops[-1] = lambda li, plan=plan: plan.launch(out=self._eps_out)

    def gn_fir(self, xa, xb, ss_fn, eps, adagn, up):
        """Fused prologue of an up / down res block: returns (FIR(SiLU(GN([xa|xb]))), [FIR(xa), FIR(xb)])."""
        self.ensure_stats(xa)
        if xb is not None:
            self.ensure_stats(xb)
        C = xa.C + (xb.C if xb is not None else 0)
        groups = self.fixed_groups or gn_groups(C)
        H2, W2 = (xa.H * 2, xa.W * 2) if up else (xa.H // 2, xa.W // 2)
        h = self.new_act(H2, W2, C)
        ra = self.new_act(H2, W2, xa.C)
        rb = self.new_act(H2, W2, xb.C) if xb is not None else None

        def run(li):
            ops.gn_fir(xa.t, xa.C, xb.t if xb is not None else None, xb.C if xb is not None else 0, xa.B, xa.H, xa.W,
                       self._stats_view(xa), self._stats_view(xb) if xb is not None else None, groups, eps, ss_fn(li),
                       adagn, up, h.t, ra.t, rb.t if rb is not None else None)
        nin = xa.t.numel() + (xb.t.numel() if xb is not None else 0)
        self._op(run, "gn_fir", dict(bytes=(nin + 2 * h.t.numel()) * 2))
        return h, [ra] + ([rb] if rb is not None else [])

    def fir(self, a, up):
        H2, W2 = (a.H * 2, a.W * 2) if up else (a.H // 2, a.W // 2)
        out = self.new_act(H2, W2, a.C)
        self._op(lambda li: ops.fir_resample(a.t, out.t, a.B, a.H, a.W, a.C, up, x_lo=a.lo, y_lo=out.lo), "fir",
                 dict(bytes=(a.t.numel() + out.t.numel()) * 2))
        return out

    def attn_core(self, x, ss, gn_eps, ws, bs, heads, out_alpha):
        """out = out_alpha * (x + OUT(softmax(q k^T / sqrt(d)) v)) with q,k,v = 1x1 projections of GN(x).
        ws/bs: [Wq, Wk, Wv, Wo] as (out, in) fp32 and fp32 biases.  The softmax(QK^T)V core is the fused tcgen05
        attention kernel (evc_attn_*) behind ONE q|k|v projection when N % 64 == 0 and the head dim is a multiple of 64
        (ops.attn_supported); otherwise (and always in split-precision mode) batched GEMMs with a per-sample B operand
        (K, then V^T from a transposed-store epilogue) + row softmax."""
        dev = self.device
        C, N, B = x.C, x.H * x.W, self.B
        d = C // heads
        if d % 8 != 0:
            raise EvcError("attention head dim must be a multiple of 8")
        self._keep += [ss] + list(bs)
        wq, wk, wv, wo = ws
        bq, bk, bv, bo = bs
        hn = self.new_act(x.H, x.W, C)
        self.gn_apply(x, None, lambda li: ss, gn_eps, False, False, hn)
        sp = self.split
        fused = self.fused_attention and not sp and ops.attn_supported(N, C, heads)
        if fused:
            # one q|k|v projection; the attention kernel reads V rows as an MN-major operand (no transposed V^T store)
            vT = None
            if self.fused_qkv:
                qkv = self.pool.get((B, x.H, x.W, 3 * C))
                self.gemm([(hn, 1)], torch.cat([wq, wk, wv], 0).contiguous(), qkv, EVC_OUT_BF16_ROWS, 3 * C,
                          bias=torch.cat([bq, bk, bv]).contiguous())
            else:  # A/B (EVC_FUSED_QKV=0): round-2a layout, q|k rows + V^T from a transposed-store projection
                qkv = self.pool.get((B, x.H, x.W, 2 * C))
                vT = self.pool.get((B, C, N))
                self.gemm([(hn, 1)], torch.cat([wq, wk], 0).contiguous(), qkv, EVC_OUT_BF16_ROWS, 2 * C,
                          bias=torch.cat([bq, bk]).contiguous())
                self.gemm([(hn, 1)], wv, vT, EVC_OUT_BF16_T, N, out_bs=C * N, bias=bv)
            o = self.new_act(x.H, x.W, C)
            qkv3 = qkv.view(B, N, qkv.shape[-1])
            plan = ops.AttnPlan(qkv3, vT, o.t.view(B, N, C), heads, float(int(d) ** (-0.5)),
                                v=qkv3[:, :, 2 * C:] if vT is None else None)
            self.flops += plan.flops
            self._op(lambda li, plan=plan: plan.launch(), "attn", dict(flops=plan.flops, N=N, d=d, heads=heads))
            out = self.new_act(x.H, x.W, C, scratch=False)
            self.gemm([(o, 1)], wo, out.t, EVC_OUT_BF16_ROWS, C, bias=bo, resid=x, alpha=out_alpha, stats_of=out)
            self.pool.put(qkv)
            if vT is not None:
                self.pool.put(vT)
            self.release(hn, o)
            return out
        qk = self.pool.get((B, x.H, x.W, 2 * C))
        qk_lo = self.pool.get((B, x.H, x.W, 2 * C)) if sp else None
        self.gemm([(hn, 1)], torch.cat([wq, wk], 0).contiguous(), qk, EVC_OUT_BF16_ROWS, 2 * C,
                  bias=torch.cat([bq, bk]).contiguous(), out_lo=qk_lo)
        # key axis padded to a multiple of 8 (16-byte TMA strides); only toy shapes (N < 8) ever pad.  Pad columns:
        # S = -inf (never written by the GEMM, so softmax gives P = 0) and V^T = 0.
        Np = max(8, (N + 7) // 8 * 8)
        S = Pm = Pm_lo = vT_lo = None
        if Np == N:
            vT = self.pool.get((B, C, N))
            S = self.pool.get((B, N, N), torch.float32)
            Pm = self.pool.get((B, N, N))
            if sp:
                vT_lo = self.pool.get((B, C, N))
                Pm_lo = self.pool.get((B, N, N))
        else:
            vT = torch.zeros((B, C, Np), dtype=torch.bfloat16, device=dev)
            S = torch.full((B, N, Np), float("-inf"), dtype=torch.float32, device=dev)
            Pm = torch.zeros((B, N, Np), dtype=torch.bfloat16, device=dev)
            if sp:
                vT_lo = torch.zeros_like(vT)
                Pm_lo = torch.zeros_like(Pm)
        self.gemm([(hn, 1)], wv, vT, EVC_OUT_BF16_T, Np, out_bs=C * Np, bias=bv, out_lo=vT_lo)
        o = self.new_act(x.H, x.W, C)
        qk3 = qk.view(B, N, 2 * C)
        o3 = o.t.view(B, N, C)
        qk3_lo = qk_lo.view(B, N, 2 * C) if sp else None
        o3_lo = o.lo.view(B, N, C) if sp else None
        for hd in range(heads):
            q = qk3[:, :, hd * d:(hd + 1) * d].unsqueeze(1)  # (B,1,N,d)
            k = qk3[:, :, C + hd * d:C + (hd + 1) * d]  # (B,N,d)  per-sample B operand
            q_lo = [qk3_lo[:, :, hd * d:(hd + 1) * d].unsqueeze(1)] if sp else None
            k_lo = qk3_lo[:, :, C + hd * d:C + (hd + 1) * d] if sp else None
            self.gemm([(q, 1)], k, S, EVC_OUT_F32_ROWS, Np, alpha=float(int(d) ** (-0.5)), w_lo=k_lo, segs_lo=q_lo)
            self._op(lambda li, S=S, Pm=Pm, Pl=Pm_lo: ops.softmax_rows(S, Pm, B * N, Np, P_lo=Pl), "softmax",
                     dict(bytes=B * N * Np * 6))
            self.gemm([(Pm.view(B, 1, N, Np), 1)], vT[:, hd * d:(hd + 1) * d, :], o3[:, :, hd * d:],
                      EVC_OUT_BF16_ROWS, C, w_lo=vT_lo[:, hd * d:(hd + 1) * d, :] if sp else None,
                      segs_lo=[Pm_lo.view(B, 1, N, Np)] if sp else None, out_lo=o3_lo[:, :, hd * d:] if sp else None)
        out = self.new_act(x.H, x.W, C, scratch=False)
        self.gemm([(o, 1)], wo, out.t, EVC_OUT_BF16_ROWS, C, bias=bo, resid=x, alpha=out_alpha, stats_of=out)
        self.pool.put(qk)
        if sp:
            self.pool.put(qk_lo)
        if Np == N:
            for t in (vT, S, Pm, vT_lo, Pm_lo):
                if t is not None:
                    self.pool.put(t)
        else:
            self._keep += [t for t in (vT, S, Pm, vT_lo, Pm_lo) if t is not None]
        self.release(hn, o)
        return out

    # ----------------------------------------------------------------- execution
    def finalize(self):
        self.stats_arena = torch.zeros(max(self.stats_total, 2), dtype=torch.int64, device=self.device)
        self.workspace = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.device)
        for a in self._stat_acts:
            _, off, n = a.stats
            a.stats = self.stats_arena[off:off + n]
        for idx, segs, w, out_t, out_mode, out_ld, plan_args, act, bias_fn, gn in self._deferred:
            if gn is not None:
                N = w.shape[-2]
                dummy = torch.zeros(2 * N, dtype=torch.float32, device=self.device)  # always overridden at launch
                ticket = gn["ticket"].stats.view(torch.int32)
                self._keep += [dummy, ticket]
                plan = ops.GemmPlan(segs, w, out_t, out_mode, out_ld, stats=act.stats,
                                    gn=dict(ss=dummy, ticket=ticket, eps=gn["eps"], groups=gn["groups"], adagn=gn["adagn"]),
                                    **plan_args)
                if bias_fn is None:
                    self.ops[idx] = (lambda li, plan=plan, fn=gn["ss_fn"]: plan.launch(gn_ss=fn(li)))
                else:
                    self.ops[idx] = (lambda li, plan=plan, fn=gn["ss_fn"], bf=bias_fn: plan.launch(bf(li), gn_ss=fn(li)))
                continue
            plan = ops.GemmPlan(segs, w, out_t, out_mode, out_ld, stats=act.stats, **plan_args)
            if bias_fn is None:
                self.ops[idx] = (lambda li, plan=plan: plan.launch())
            else:
                self.ops[idx] = (lambda li, plan=plan, bf=bias_fn: plan.launch(bf(li)))
        self._deferred = []

    def forward(self, label_idx=0, eps_out=None):
        """Enqueue one UNet evaluation: reads self.xin, writes self.eps (B, C_out, H, W) fp32 -- or `eps_out`, a
        contiguous fp32 tensor of the same shape (not available in split-precision mode)."""
        if eps_out is not None:
            if self.split or eps_out.shape != self.eps.shape or eps_out.dtype != torch.float32 or not eps_out.is_contiguous():
                raise EvcError("eps_out must be a contiguous fp32 tensor shaped like eps (bf16 mode only)")
        self._eps_out = eps_out
        ops.fill_zero(self.stats_arena)  # fused statistics accumulate with integer atomics
        for fn in self.ops:
            fn(label_idx)
        self._eps_out = None
        return self.eps if eps_out is None else eps_out

    def profile(self, label_idx=0, reps=3):
        """Per-launch device time of one UNet evaluation (CUDA events around every launch, eager mode).
        Returns [(kind, meta, best_ms)] in launch order; used by bench.py for the roofline of the GEMM kernel."""
        n = len(self.ops)
        best = [float("inf")] * n
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
            ev[0].record()
            for i, fn in enumerate(self.ops):
                fn(label_idx)
                ev[i + 1].record()
            torch.cuda.synchronize(self.device)
            for i in range(n):
                best[i] = min(best[i], ev[i].elapsed_time(ev[i + 1]))
        return [(k, m, best[i]) for i, (k, m) in enumerate(self.op_info)]


class NCSNppEngine(EngineBase):
    """Launch plan of NCSNpp.forward (models/better/ncsnpp_more.py:251-392) for batch size B."""

    def __init__(self, net, B, device, precision="bf16"):
        cfg = net.config
        H = cfg.data.image_size
        super().__init__(device, B, H, precision)
        self.net = net
        self.cfg = cfg
        m, d = cfg.model, cfg.data
        self.nf = m.ngf
        self.c_x = d.channels * d.num_frames
        self.c_cond = d.channels * (d.num_frames_cond + getattr(d, "num_frames_future", 0))
        self.head_ch = getattr(m, "n_head_channels", -1)
        if self.nf % 8 != 0:
            raise EvcError("evcdiff CUDA path needs model.ngf % 8 == 0 (16-byte channel vectors; % 64 for full speed)")
        self.c_cond_off = cond_offset(self.c_x)
        if self.c_cond_off + self.c_cond > CIN_PAD:
            raise EvcError("more than 64 input channels is not supported")
        sd = {k: v.detach() for k, v in net.state_dict().items()}
        self.sd = sd
        self.P = lambda i: f"all_modules.{i}"
        self._build(sd)
        self.finalize()

    # ------------------------------------------------------------------ weights
    def f32(self, key):
        t = self.sd[key].to(self.device, torch.float32).contiguous()
        self._keep.append(t)
        return t

    def _build(self, sd):
        from .models.better.ncsnpp_more import ncsnpp_spec
        dev = self.device
        B, H = self.B, self.H
        spec = ncsnpp_spec(self.cfg)
        self.spec = spec
        P = self.P
        # --- time-embedding MLP + all AdaGN projections (evaluated per label set, see set_labels)
        self.temb_w0, self.temb_b0 = self.f32(P(0) + ".weight"), self.f32(P(0) + ".bias")
        self.temb_w1, self.temb_b1 = self.f32(P(1) + ".weight"), self.f32(P(1) + ".bias")
        half = self.nf // 2
        e = math.log(10000) / (half - 1)
        self.freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -e).to(dev)
        dw, db, self.ss_off = [], [], {}
        off = 0
        for i, s in enumerate(spec):
            if s["kind"] == "res":
                for j in (0, 1):
                    k = f"{P(i)}.actnorm{j}.Dense_0"
                    dw.append(sd[k + ".weight"].float())
                    db.append(sd[k + ".bias"].float())
                    self.ss_off[(i, j)] = (off, dw[-1].shape[0])
                    off += dw[-1].shape[0]
        self.dense_w = torch.cat(dw, 0).to(dev).contiguous()
        self.dense_b = torch.cat(db, 0).to(dev).contiguous()
        self.ss_total = off
        self.ss_table = None  # (L, ss_total) fp32, filled by set_labels

        # --- buffers
        self.xin = torch.zeros((B, H, H, CIN_PAD), dtype=torch.bfloat16, device=dev)
        self.xin_lo = torch.zeros_like(self.xin) if self.split else None
        self.eps = torch.zeros((B, self.c_x, H, H), dtype=torch.float32, device=dev)
        xin = Act(self.xin, self.xin_lo)

        m = self.cfg.model
        nres, nlev = m.num_res_blocks, len(m.ch_mult)
        attn_res = list(m.attn_resolutions)
        i = 2
        h0 = self.new_act(H, H, spec[i]["cout"], scratch=False)
        self.gemm([(xin, 9)], pack_conv_in(sd[P(i) + ".weight"].to(dev), self.c_x, CIN_PAD), h0.t, EVC_OUT_BF16_ROWS, h0.C,
                  bias=self.f32(P(i) + ".bias"), stats_of=h0)
        self.taps["m2"] = h0
        i += 1
        hs = [h0]
        for lvl in range(nlev):
            for _ in range(nres):
                h = self.res_block(i, spec[i], hs[-1], None)
                i += 1
                if h.W in attn_res:
                    h = self.attn_block(i, spec[i], h)
                    i += 1
                hs.append(h)
            if lvl != nlev - 1:
                hs.append(self.res_block(i, spec[i], hs[-1], None))
                i += 1
        h = hs[-1]
        h = self.res_block(i, spec[i], h, None); i += 1
        h = self.attn_block(i, spec[i], h); i += 1
        h = self.res_block(i, spec[i], h, None); i += 1
        for lvl in reversed(range(nlev)):
            for _ in range(nres + 1):
                h = self.res_block(i, spec[i], h, hs.pop())
                i += 1
            if h.W in attn_res:
                h = self.attn_block(i, spec[i], h)
                i += 1
            if lvl != 0:
                h = self.res_block(i, spec[i], h, None)
                i += 1
        assert not hs
        # final GroupNorm(affine)+SiLU and conv3x3 -> NCHW fp32 eps
        ss = torch.cat([sd[P(i) + ".Norm_0.weight"].float(), sd[P(i) + ".Norm_0.bias"].float()]).to(dev).contiguous()
        self._keep.append(ss)
        hn = self.new_act(h.H, h.W, h.C)
        self.gn_apply(h, None, lambda li: ss, 1e-5, False, True, hn)
        self.taps[f"m{i}"] = hn
        i += 1
        self.final_conv([(hn, 9)], pack_conv3(sd[P(i) + ".weight"].to(dev)), self.f32(P(i) + ".bias"), self.c_x, H)
        i += 1
        assert i == len(spec)

    # ------------------------------------------------------------------ blocks
    def _ss(self, i, j):
        off, n = self.ss_off[(i, j)]
        return lambda li: self.ss_table[li, off:off + n]

    def res_block(self, i, s, xa, xb):
        """ResnetBlockBigGANppGN (layerspp.py:595-624) on the virtual concat [xa | xb]."""
        sd, P, dev = self.sd, self.P, self.device
        cin, cout = s["cin"], s["cout"]
        assert cin == xa.C + (xb.C if xb is not None else 0)
        xs = [xa] + ([xb] if xb is not None else [])
        tmp = []
        fused = (s["up"] or s["down"]) and not self.split and os.environ.get("EVC_GN_FIR", "1") != "0"
        if fused:
            h, xs = self.gn_fir(xa, xb, self._ss(i, 0), 1e-5, True, s["up"])
            tmp += xs
        else:
            h = self.new_act(xa.H, xa.W, cin)
            self.gn_apply(xa, xb, self._ss(i, 0), 1e-5, True, True, h)
        if (s["up"] or s["down"]) and not fused:
            h2 = self.fir(h, s["up"])
            self.release(h)
            h = h2
            xs = [self.fir(x, s["up"]) for x in xs]
            tmp += xs
        a1 = self.new_act(h.H, h.W, cout)
        if self.gn_fusable(h.B, h.H, h.W, [(h.C, 9)], cout):
            # Conv_0 -> GroupNorm_1 -> SiLU in one launch: the raw convolution output never reaches memory
            c0 = None
            self.gemm([(h, 9)], pack_conv3(sd[P(i) + ".Conv_0.weight"].to(dev)), a1.t, EVC_OUT_BF16_ROWS, cout,
                      bias=self.f32(P(i) + ".Conv_0.bias"),
                      gn=dict(ss_fn=self._ss(i, 1), eps=1e-5, adagn=True, groups=self.fixed_groups or gn_groups(cout)))
            self.release(h)
        else:
            c0 = self.new_act(h.H, h.W, cout)
            self.gemm([(h, 9)], pack_conv3(sd[P(i) + ".Conv_0.weight"].to(dev)), c0.t, EVC_OUT_BF16_ROWS, cout,
                      bias=self.f32(P(i) + ".Conv_0.bias"), stats_of=c0)
            self.release(h)
            self.gn_apply(c0, None, self._ss(i, 1), 1e-5, True, True, a1)
        out = self.new_act(a1.H, a1.W, cout, scratch=False)
        w1 = pack_conv3(sd[P(i) + ".Conv_1.weight"].to(dev))
        b1 = sd[P(i) + ".Conv_1.bias"].float()
        if (P(i) + ".Conv_2.weight") in sd:
            w2 = sd[P(i) + ".Conv_2.weight"].to(dev).float().reshape(cout, cin)
            w = torch.cat([w1, w2], dim=1).contiguous()
            bias = (b1 + sd[P(i) + ".Conv_2.bias"].float()).to(dev).contiguous()
            self._keep.append(bias)
            self.gemm([(a1, 9)] + [(x, 1) for x in xs], w, out.t, EVC_OUT_BF16_ROWS, cout, bias=bias, alpha=RSQRT2,
                      stats_of=out)
        else:
            assert len(xs) == 1 and xs[0].C == cout
            bias = b1.to(dev).contiguous()
            self._keep.append(bias)
            self.gemm([(a1, 9)], w1, out.t, EVC_OUT_BF16_ROWS, cout, bias=bias, resid=xs[0], alpha=RSQRT2,
                      stats_of=out)
        self.release(c0, a1, *tmp)
        self.taps[f"m{i}"] = out
        return out

    def attn_block(self, i, s, x):
        """AttnBlockpp (layerspp.py:230-249): GN(affine, eps 1e-6) -> q,k,v NIN -> softmax(q k^T / sqrt(d)) v -> NIN_3."""
        sd, P, dev = self.sd, self.P, self.device
        C = x.C
        heads = 1 if (self.head_ch == -1 or C < self.head_ch) else C // self.head_ch
        ss = torch.cat([sd[P(i) + ".GroupNorm_0.weight"].float(), sd[P(i) + ".GroupNorm_0.bias"].float()]).to(dev)
        # NIN: y = x @ W + b with W (in, out) -> GEMM weight rows = outputs
        ws = [sd[P(i) + f".NIN_{j}.W"].to(dev).float().t().contiguous() for j in range(4)]
        bs = [sd[P(i) + f".NIN_{j}.b"].float().to(dev).contiguous() for j in range(4)]
        out = self.attn_core(x, ss.contiguous(), 1e-6, ws, bs, heads, RSQRT2)
        self.taps[f"m{i}"] = out
        return out

    # ------------------------------------------------------------------ per-label tables
    def set_labels(self, labels):
        """labels: iterable of floats (the distinct `y` values the sampler will use, in order).  Evaluates
        get_timestep_embedding + the temb MLP + every Dense_0(SiLU(temb)) once per label (layers.py:504-518,
        ncsnpp_more.py:277-281, layerspp.py:520-522)."""
        lab = torch.tensor([float(v) for v in labels], dtype=torch.float32, device=self.device)
        L = lab.numel()
        emb = torch.empty((L, self.nf), dtype=torch.float32, device=self.device)
        ops.timestep_embedding(lab, self.freqs, self.nf, emb)
        t1 = torch.empty((L, 4 * self.nf), dtype=torch.float32, device=self.device)
        ops.linear_f32(emb, self.temb_w0, self.temb_b0, t1)
        t2 = torch.empty_like(t1)
        ops.linear_f32(t1, self.temb_w1, self.temb_b1, t2, act_in=True)
        if self.ss_table is None or self.ss_table.shape[0] < L:
            self.ss_table = torch.empty((L, self.ss_total), dtype=torch.float32, device=self.device)
        ops.linear_f32(t2, self.dense_w, self.dense_b, self.ss_table[:L], act_in=True)
        self.labels = [float(v) for v in labels]
        self.temb = t2
        return L

    def load_input(self, x, cond):
        """x (B,15,H,W) fp32, cond (B,6,H,W) fp32/fp64 or None -> NHWC bf16 UNet input (torch.cat of
        ncsnpp_more.py:256-257 + the fp32 cast of :293, fused with the layout change)."""
        ops.pack_nchw(x.contiguous(), self.xin, 0, dst_lo=self.xin_lo)
        if cond is not None:
            ops.pack_nchw(cond.contiguous(), self.xin, self.c_cond_off, dst_lo=self.xin_lo)

    def refresh_x(self, x):
        """Split-precision mode: the sampler-update kernels only write the bf16 hi plane of x_t; rewrite both planes."""
        if self.split:
            ops.pack_nchw(x, self.xin, 0, dst_lo=self.xin_lo)
