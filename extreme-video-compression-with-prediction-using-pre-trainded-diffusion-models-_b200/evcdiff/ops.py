"""Tensor-level wrappers over the C ABI.  torch is used for device memory and the current stream only."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (EVC_OUT_BF16_ROWS, EVC_OUT_BF16_T, EVC_OUT_F32_ROWS, EVC_OUT_F32_T, AttnDesc, GemmDesc, PndmCoef,
                   StepCoef, check, load, stream_ptr)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.EvcError("evcdiff kernels need CUDA tensors (no CPU fallback)")


class GemmPlan:
    """One implicit-GEMM launch: out = alpha * (sum_seg conv_taps(A_seg) @ W^T + bias + resid).

    segs: list of (tensor viewed as (B,H,W,C) bf16 with unit channel stride, taps in {1,9})
    w:    bf16 (N,K) shared weights or (B,N,K) per-sample operand, K contiguous
    out:  bf16/fp32 tensor; out_mode selects row-major pixel rows or channel-major (transposed) layout
    """

    def __init__(self, segs, w, out, out_mode, out_ld, out_bs=0, bias=None, resid=None, resid_ld=0, alpha=1.0,
                 bn=None, max_ctas=0, stats=None, stride=1, cta_group=None, segs_lo=None, w_lo=None, out_lo=None,
                 resid_lo=None, gn=None, split_k=1, sk_ws=None):
        lib = load()
        _require_cuda(w, out, bias, resid, *[s[0] for s in segs])
        d = GemmDesc()
        d.n_seg = len(segs)
        B, H, W_, _ = segs[0][0].shape
        H, W_ = H // stride, W_ // stride
        d.stride = stride
        ktot = 0
        for i, (a, taps) in enumerate(segs):
            assert a.dtype == torch.bfloat16 and a.dim() == 4 and a.stride(3) == 1, "A segment must be (B,H,W,C) bf16"
            d.a[i].ptr = a.data_ptr()
            d.a[i].B, d.a[i].H, d.a[i].W, d.a[i].C = a.shape
            d.a[i].stride_b, d.a[i].stride_h, d.a[i].stride_w = a.stride(0), a.stride(1), a.stride(2)
            d.taps[i] = taps
            ktot += taps * a.shape[3]
        assert w.dtype == torch.bfloat16 and w.stride(-1) == 1
        if w.dim() == 2:
            d.w_rows, d.w_k, d.w_batches = w.shape[0], w.shape[1], 1
            d.w_row_stride, d.w_batch_stride = w.stride(0), 0
        else:
            d.w_batches, d.w_rows, d.w_k = w.shape
            d.w_batch_stride, d.w_row_stride = w.stride(0), w.stride(1)
        assert d.w_k == ktot, f"weight K {d.w_k} != sum(taps*C) {ktot}"
        d.w = w.data_ptr()
        d.B, d.H, d.W = B, H, W_
        n = d.w_rows
        kblocks = sum(taps * (-(-a.shape[3] // 64)) for a, taps in segs)
        can_split = (split_k == "auto" and w.dim() == 2 and gn is None and w_lo is None and sk_ws is not None)
        if bn is None:
            bn, auto_s = pick_tile(n, m_tiles(B, H, W_, False), kblocks, allow_split=can_split) if w.dim() == 2 else \
                (pick_bn(n, m_tiles(B, H, W_, True), kblocks), 1)
        else:
            auto_s = 1
        if split_k == "auto":
            split_k = auto_s if can_split else 1
        d.bn = bn
        d.out = out.data_ptr()
        d.out_mode = out_mode
        d.out_ld = out_ld
        d.out_bs = out_bs
        d.bias = bias.data_ptr() if bias is not None else None
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() >= n
        d.resid = resid.data_ptr() if resid is not None else None
        d.resid_ld = resid_ld
        d.alpha = alpha
        d.max_ctas = max_ctas
        if cta_group is None:
            cta_group = int(os.environ.get("EVC_GEMM_CTA_GROUP", "0"))
            if cta_group == 2 and (w.dim() == 3 or bn % 16 != 0):
                cta_group = 1  # the env override only applies where pairing is possible
        d.cta_group = cta_group
        if stats is not None:
            assert stats.dtype == torch.int64 and stats.is_cuda
            d.stats = stats.data_ptr()
        if w_lo is not None:  # split-precision operands: same shapes / strides as the hi planes
            assert segs_lo is not None and len(segs_lo) == len(segs) and w_lo.shape == w.shape and w_lo.stride() == w.stride()
            _require_cuda(w_lo, out_lo, resid_lo, *segs_lo)
            for i, a in enumerate(segs_lo):
                hi = segs[i][0]
                assert a.dtype == torch.bfloat16 and a.shape == hi.shape and a.stride() == hi.stride()
                d.a_lo[i].ptr = a.data_ptr()
                d.a_lo[i].B, d.a_lo[i].H, d.a_lo[i].W, d.a_lo[i].C = a.shape
                d.a_lo[i].stride_b, d.a_lo[i].stride_h, d.a_lo[i].stride_w = a.stride(0), a.stride(1), a.stride(2)
            d.w_lo = w_lo.data_ptr()
            d.out_lo = out_lo.data_ptr() if out_lo is not None else None
            d.resid_lo = resid_lo.data_ptr() if resid_lo is not None else None
        if gn is not None:
            # fused GroupNorm apply: dict(ss=fp32 [gamma'|beta'] (2N), ticket=int32 (B), eps, groups, adagn)
            _require_cuda(gn["ss"], gn["ticket"])
            assert gn["ss"].dtype == torch.float32 and gn["ticket"].dtype == torch.int32 and gn["ticket"].numel() >= B
            d.gn_ss = gn["ss"].data_ptr()
            d.gn_ticket = gn["ticket"].data_ptr()
            d.gn_eps = float(gn["eps"])
            d.gn_groups = int(gn["groups"])
            d.gn_adagn = int(bool(gn["adagn"]))
        self.split_k = 1
        sk_ticket = None
        if split_k and split_k > 1:
            # split-K: K slices of a tile on different CTAs, slices added in a fixed order by the last arriver
            mt2 = (m_tiles(B, H, W_, w.dim() == 3) + 1) // 2 * 2
            tiles_n = -(-n // bn)
            need = split_k * mt2 * 128 * tiles_n * bn * 4
            assert sk_ws is not None and sk_ws.numel() * sk_ws.element_size() >= need, "split-K workspace too small"
            sk_ticket = torch.zeros(mt2 * tiles_n, dtype=torch.int32, device=out.device)
            d.split_k = int(split_k)
            d.sk_ws = sk_ws.data_ptr()
            d.sk_ws_bytes = sk_ws.numel() * sk_ws.element_size()
            d.sk_ticket = sk_ticket.data_ptr()
            self.split_k = int(split_k)
        self._keep = (segs, w, out, bias, resid, stats, segs_lo, w_lo, out_lo, resid_lo, gn, sk_ws, sk_ticket)
        self._lib = lib
        h = C.c_void_p()
        check(lib.evc_gemm_plan_create(C.byref(d), C.byref(h)), "evc_gemm_plan_create")
        self._h = h
        self.flops = lib.evc_gemm_plan_flops(h)
        self.cta_group = lib.evc_gemm_plan_cta_group(h)

    def launch(self, bias_override=None, gn_ss=None, out=None):
        """out: redirect a per-thread-store output (fp32 / transposed modes) to another tensor of the same layout."""
        if gn_ss is None and out is None:
            check(self._lib.evc_gemm_plan_launch(self._h, _ptr(bias_override), stream_ptr()), "evc_gemm_plan_launch")
        else:
            check(self._lib.evc_gemm_plan_launch_ex(self._h, _ptr(bias_override), _ptr(gn_ss), _ptr(out), stream_ptr()),
                  "evc_gemm_plan_launch_ex")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.evc_gemm_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def attn_supported(N, C, heads):
    d = C // heads
    if N % 64 or d % 64:
        return False
    if d <= 384:
        return d <= 256 or d % 128 == 0
    # larger heads: output columns split over CTAs in slices of <= 256 columns (multiples of 64)
    return any(d % s == 0 and (d // s) % 64 == 0 and d // s <= 256 for s in range(2, 17))


class AttnPlan:
    """Fused attention launch: out (B,N,C) = softmax(scale q k^T) v per head, from qk (B,N,>=2C) and either vT (B,C,N)
    or, with vT=None, v (B,N,C) rows (a strided view, e.g. the last C columns of a fused q|k|v projection)."""

    def __init__(self, qk, vT, out, heads, scale, v=None):
        lib = load()
        B, N, C2 = qk.shape
        d = AttnDesc()
        if vT is not None:
            _require_cuda(qk, vT, out)
            Cc = C2 // 2
            assert vT.shape == (B, Cc, N) and vT.stride(2) == 1
            d.vT, d.vT_ld = vT.data_ptr(), vT.stride(1)
        else:
            _require_cuda(qk, v, out)
            Cc = v.shape[2]
            assert v.shape == (B, N, Cc) and v.stride(2) == 1 and v.stride(0) == N * v.stride(1) and C2 >= 2 * Cc
            d.v, d.v_ld = v.data_ptr(), v.stride(1)
        assert out.shape[:2] == (B, N) and qk.stride(2) == 1 and qk.stride(0) == N * qk.stride(1)
        d.qk, d.qk_ld = qk.data_ptr(), qk.stride(1)
        d.out, d.out_ld = out.data_ptr(), out.stride(1)
        d.B, d.N, d.C, d.heads, d.scale = B, N, Cc, heads, scale
        self._keep = (qk, vT, v, out)
        self._lib = lib
        h = C.c_void_p()
        check(lib.evc_attn_plan_create(C.byref(d), C.byref(h)), "evc_attn_plan_create")
        self._h = h
        self.flops = lib.evc_attn_plan_flops(h)

    def launch(self):
        check(self._lib.evc_attn_plan_launch(self._h, stream_ptr()), "evc_attn_plan_launch")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.evc_attn_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def m_tiles(B, H, W, batched):
    """Number of 128-pixel M tiles the kernel will use (mirrors evc_gemm_plan_create)."""
    tw = min(W, 128)
    th = max(1, min(H, 128 // tw))
    tb = 1 if batched else max(1, 128 // (tw * th))
    return (W // tw) * (H // th) * (-(-B // tb))


def _pairs(mt, bn):
    """CTA pairs?  (mirrors evc_gemm_plan_create)"""
    return mt >= 8 and (mt % 2 == 0 or mt >= 23) and bn % 16 == 0


def kblock_cycles(bn, cg):
    """SM cycles per 64-wide K block of one 128-row tile: four tcgen05.mma (N/2 cycles each, ~48 at least:
    profiles/r02_umma_microbench.jsonl) or the delivery of 128 A rows + this CTA's share of the B rows by the TMA unit,
    whichever is longer.  Measured (profiles/r02_bn_probe.txt, 6 videos): 389 / ~285 / ~280 cycles for N tiles of 192 / 96 /
    64 in CTA pairs -- a narrower N tile re-reads the A tile, so it is never much cheaper per tile."""
    row = 1.74 if cg == 2 else TMA_CYCLES_PER_ROW
    return max(4 * max(bn // 2, 48), row * (128 + bn / cg))


def unsplit_cost(n, mt, kblocks, bn, sms):
    tiles = mt * (n // bn)
    waves = -(-tiles // sms)
    epi = 1500 + 12 * bn  # drains behind the next tile's main loop (two accumulator stages) unless the K loop is shorter
    per_tile = max(kblocks * kblock_cycles(bn, 2 if _pairs(mt, bn) else 1), epi)
    return waves * per_tile + epi


def pick_bn(n, mt=None, kblocks=None, sms=148):
    """N tile of the 128-row UMMA.  Large problems: the widest tile that divides N (fewest A re-reads, best
    MMA efficiency).  Problems with few M tiles (8x8 / 16x16 levels, small batches): the tile that minimises
    waves x per-tile time, so that all SMs get work."""
    cands = [bn for bn in (256, 192, 128, 96, 64, 48, 32, 16) if n % bn == 0]
    if not cands:
        return min(256, ((n + 15) // 16) * 16)
    if mt is None or kblocks is None:
        return cands[0]
    # more tiles than SMs even with the widest N tile: several rounds of the persistent grid, where a narrower tile pays
    # the A re-read in every round (unsplit_cost); otherwise the fill-the-SMs model below.  EVC_PICK_MODEL=old: A/B
    old = os.environ.get("EVC_PICK_MODEL", "new") == "old" or mt * (n // cands[0]) <= sms
    best, best_cost = None, None
    for bn in cands:
        if old:
            tiles = mt * (n // bn)
            waves = -(-tiles // sms)
            cost = waves * (kblocks * 4 * max(bn // 2, 32) + 1500 + 12 * bn)
        else:
            cost = unsplit_cost(n, mt, kblocks, bn, sms)
        if best_cost is None or cost < best_cost * 0.97:  # prefer wider tiles unless clearly slower
            best, best_cost = bn, cost
    return best


SPLIT_K_WS_BYTES = 48 << 20  # fp32 partial tiles of one split-K launch (pick_tile keeps its choices below this)
TMA_CYCLES_PER_ROW = 2.4     # measured: one SM's TMA unit delivers a 128-byte box row every ~2.4 cycles (profiles/r02_notes.md)


def pick_tile(n, mt, kblocks, sms=None, allow_split=True):
    """(N tile, K slices) of a launch.  Large problems: the widest N tile (fewest operand re-reads per FLOP), one slice.
    Launches with few 128-row M tiles (small batches, the 8x8 / 16x16 levels) are bound by how fast ONE SM streams its
    operands -- 128 A rows + bn B rows per 64-wide K block through the TMA unit -- so instead of shrinking the N tile
    until every SM has a tile (which re-reads A once per N tile), the K loop of a wide tile is cut into slices that run
    on otherwise idle SMs (split-K, slices summed in a fixed order)."""
    sms = sms or num_sms()
    if not allow_split or os.environ.get("EVC_GEMM_SPLIT_K", "1") == "0":
        return pick_bn(n, mt, kblocks, sms), 1
    cands = [bn for bn in (256, 192, 128, 96, 64, 48, 32, 16) if n % bn == 0]
    if not cands:
        return pick_bn(n, mt, kblocks, sms), 1
    best1, bests = None, None  # best unsplit / best split candidate: (cost, bn, S)
    for bn in cands:
        tiles = mt * (n // bn)
        cg = 2 if (mt >= 8 and (mt % 2 == 0 or mt >= 23) and bn % 16 == 0) else 1  # mirrors evc_gemm_plan_create
        per_kb = max(2 * bn, TMA_CYCLES_PER_ROW * (128 + bn / cg))
        for S in (1, 2, 3, 4, 6, 8, 12, 16):
            if S > 1 and (bn % 32 != 0 or kblocks // S < 4 or tiles * S > sms or
                          S * ((mt + 1) // 2 * 2) * 128 * n * 4 > SPLIT_K_WS_BYTES):
                continue
            waves = -(-(tiles * S) // sms)
            pk = per_kb if S == 1 else max(2 * bn, TMA_CYCLES_PER_ROW * (128 + bn))  # split launches run unpaired
            cost = waves * (-(-kblocks // S) * pk + 1500 + 12 * bn)
            if S == 1:
                if best1 is None or cost < best1[0] * 0.97:  # prefer wider tiles unless clearly slower
                    best1 = (cost, bn, 1)
            else:
                # partial tiles through L2 + one ticket round trip
                # cooperative finish (tiles * S <= SMs, so every slice is resident): each slice's CTA reduces and stores
                # the chunks it owns
                cost += 9000 + 150 * S * max(1, (bn // 32) // S)
                if bests is None or cost < bests[0] * 0.97:
                    bests = (cost, bn, S)
    if bests is not None and bests[0] < 0.8 * best1[0]:
        return bests[1], bests[2]
    return pick_bn(n, mt, kblocks, sms), 1


def gn_fuse_fits(sample_m_tiles, m_tiles_total, tiles_n, sms=None):
    """Host mirror of the fused-GroupNorm-apply guard in evc_gemm_plan_create (gemm_tc.cu): the tiles of one sample must
    fit in two rounds of the persistent grid, otherwise a CTA would wait for a tile it owns itself."""
    sms = sms or num_sms()
    for cg in (1, 2):  # whichever grouping the plan picks (conservative: both must fit)
        cap = max(1, sms // cg)
        units_total = ((m_tiles_total + cg - 1) // cg) * tiles_n
        num_units = min(units_total, cap)
        sample_units = ((sample_m_tiles + cg - 1) // cg) * tiles_n + (tiles_n if cg > 1 else 0)
        if sample_units > 2 * num_units:
            return False
    return True


_sms = {}


def num_sms():
    dev = torch.cuda.current_device()
    if dev not in _sms:
        _sms[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _sms[dev]


def gn_stats_workspace_bytes(B, HW, Cc):
    n = C.c_int64(0)
    check(load().evc_gn_stats_workspace(B, HW, Cc, C.byref(n)), "evc_gn_stats_workspace")
    return int(n.value)


_ws_cache = {}


def _default_workspace(device, nbytes):
    key = str(device)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def gn_stats(x, B, HW, C, stats, ldx=None, workspace=None, x_lo=None):
    """stats (B,C,2) int64 (2^20 fixed point) = per-channel [sum, sumsq] of x (B*HW rows of C bf16).  workspace: zero-initialised
    uint8 scratch (shared between calls on one stream); a per-device default is used when omitted."""
    _require_cuda(x, stats, x_lo)
    if workspace is None:
        workspace = _default_workspace(x.device, gn_stats_workspace_bytes(B, HW, C))
    if x_lo is None:
        check(load().evc_gn_stats(_ptr(x), ldx or C, B, HW, C, _ptr(stats), C, 0, _ptr(workspace), workspace.numel(),
                                  stream_ptr()), "evc_gn_stats")
    else:
        check(load().evc_gn_stats_split(_ptr(x), _ptr(x_lo), ldx or C, B, HW, C, _ptr(stats), C, 0, _ptr(workspace),
                                        workspace.numel(), stream_ptr()), "evc_gn_stats_split")


def gn_apply(x0, C0, x1, C1, B, HW, stats0, stats1, groups, eps, ss, adagn, silu, y, x0_lo=None, x1_lo=None, y_lo=None):
    _require_cuda(x0, x1, stats0, stats1, ss, y, x0_lo, x1_lo, y_lo)
    if y_lo is None:
        check(load().evc_gn_apply(_ptr(x0), C0, _ptr(x1), C1, B, HW, _ptr(stats0), _ptr(stats1), groups, eps, _ptr(ss),
                                  int(adagn), int(silu), _ptr(y), stream_ptr()), "evc_gn_apply")
    else:
        check(load().evc_gn_apply_split(_ptr(x0), _ptr(x0_lo), C0, _ptr(x1), _ptr(x1_lo), C1, B, HW, _ptr(stats0),
                                        _ptr(stats1), groups, eps, _ptr(ss), int(adagn), int(silu), _ptr(y), _ptr(y_lo),
                                        stream_ptr()), "evc_gn_apply_split")


def fir_resample(x, y, B, H, W, C, up, x_lo=None, y_lo=None):
    _require_cuda(x, y, x_lo, y_lo)
    if y_lo is None:
        check(load().evc_fir_resample(_ptr(x), _ptr(y), B, H, W, C, int(up), stream_ptr()), "evc_fir_resample")
    else:
        check(load().evc_fir_resample_split(_ptr(x), _ptr(x_lo), _ptr(y), _ptr(y_lo), B, H, W, C, int(up), stream_ptr()),
              "evc_fir_resample_split")


def gn_fir(x0, C0, x1, C1, B, H, W, stats0, stats1, groups, eps, ss, adagn, up, y_act, y_raw0, y_raw1):
    """y_act = FIR(SiLU(GN([x0|x1]))), y_raw{0,1} = FIR(x{0,1}) from one read of the inputs (up / down res blocks)."""
    _require_cuda(x0, x1, stats0, stats1, ss, y_act, y_raw0, y_raw1)
    check(load().evc_gn_fir(_ptr(x0), C0, _ptr(x1), C1, B, H, W, _ptr(stats0), _ptr(stats1), groups, eps, _ptr(ss),
                            int(adagn), int(up), _ptr(y_act), _ptr(y_raw0), _ptr(y_raw1), stream_ptr()), "evc_gn_fir")


def nearest_up2(x, y, B, H, W, C):
    _require_cuda(x, y)
    check(load().evc_nearest_up2(_ptr(x), _ptr(y), B, H, W, C, stream_ptr()), "evc_nearest_up2")


def softmax_rows(S, P, rows, cols, P_lo=None):
    _require_cuda(S, P, P_lo)
    if P_lo is None:
        check(load().evc_softmax_rows(_ptr(S), _ptr(P), rows, cols, stream_ptr()), "evc_softmax_rows")
    else:
        check(load().evc_softmax_rows_split(_ptr(S), _ptr(P), _ptr(P_lo), rows, cols, stream_ptr()), "evc_softmax_rows_split")


def timestep_embedding(labels, freqs, dim, out):
    _require_cuda(labels, freqs, out)
    check(load().evc_timestep_embedding(_ptr(labels), _ptr(freqs), labels.numel(), dim, _ptr(out), stream_ptr()),
          "evc_timestep_embedding")


def linear_f32(x, W, b, y, act_in=False, act_out=False):
    _require_cuda(x, W, b, y)
    L, K = x.shape
    N = W.shape[0]
    assert W.shape[1] == K and x.is_contiguous() and W.is_contiguous() and y.is_contiguous()
    check(load().evc_linear_f32(_ptr(x), _ptr(W), _ptr(b), _ptr(y), L, K, N, int(act_in), int(act_out), stream_ptr()),
          "evc_linear_f32")


def pack_nchw(src, dst, c_off=0, scale=1.0, shift=0.0, dst_lo=None):
    """src (B,C,H,W) fp32/fp64 contiguous -> channels [c_off, c_off+C) of dst (B,H,W,Cpad) bf16 (+ residual plane)."""
    _require_cuda(src, dst, dst_lo)
    assert src.is_contiguous() and dst.is_contiguous() and src.dtype in (torch.float32, torch.float64)
    B, Cc, H, W = src.shape
    if dst_lo is None:
        check(load().evc_pack_nchw(_ptr(src), int(src.dtype == torch.float64), B, Cc, H * W, scale, shift, _ptr(dst),
                                   dst.shape[-1], c_off, stream_ptr()), "evc_pack_nchw")
    else:
        check(load().evc_pack_nchw_split(_ptr(src), int(src.dtype == torch.float64), B, Cc, H * W, scale, shift, _ptr(dst),
                                         _ptr(dst_lo), dst.shape[-1], c_off, stream_ptr()), "evc_pack_nchw_split")


def fill_zero(t):
    _require_cuda(t)
    check(load().evc_fill_zero(_ptr(t), t.numel() * t.element_size(), stream_ptr()), "evc_fill_zero")


def sampler_update(x, eps, noise, x_out, xin, coef: StepCoef):
    _require_cuda(x, eps, noise, x_out, xin)
    B, Cc, H, W = x.shape
    cpad = xin.shape[-1] if xin is not None else Cc
    check(load().evc_sampler_update(_ptr(x), _ptr(eps), _ptr(noise), _ptr(x_out), _ptr(xin), B, Cc, H * W, cpad,
                                    C.byref(coef), stream_ptr()), "evc_sampler_update")


def pndm_update(x, eps_list, x_out, et_out, xin, coef: PndmCoef):
    _require_cuda(x, x_out, et_out, xin, *eps_list)
    B, Cc, H, W = x.shape
    cpad = xin.shape[-1] if xin is not None else Cc
    arr = (C.c_void_p * 4)(*[e.data_ptr() for e in eps_list] + [None] * (4 - len(eps_list)))
    check(load().evc_pndm_update(_ptr(x), arr, _ptr(x_out), _ptr(et_out), _ptr(xin), B, Cc, H * W, cpad,
                                 C.byref(coef), stream_ptr()), "evc_pndm_update")


def inverse_transform(x, frames):
    _require_cuda(x, frames)
    check(load().evc_inverse_transform(_ptr(x), _ptr(frames), x.numel(), stream_ptr()), "evc_inverse_transform")


def frames_to_uint8(frames, out):
    _require_cuda(frames, out)
    assert frames.dtype == torch.float32 and out.dtype == torch.uint8 and frames.numel() == out.numel()
    check(load().evc_frames_to_uint8(_ptr(frames), _ptr(out), frames.numel(), stream_ptr()), "evc_frames_to_uint8")


def gemm_fault_count():
    return int(load().evc_gemm_fault_count())


def frame_psnr(a, b, maxvalue=1.0):
    """a, b: (..., C, H, W) fp32 with identical shapes, leading dims = frames.  Returns float64 PSNR per frame."""
    _require_cuda(a, b)
    assert a.shape == b.shape and a.dtype == torch.float32 and b.dtype == torch.float32
    a, b = a.contiguous(), b.contiguous()
    frame_elems = a.shape[-1] * a.shape[-2] * a.shape[-3]
    n = a.numel() // frame_elems
    out = torch.empty(a.shape[:-3], dtype=torch.float64, device=a.device)
    check(load().evc_frame_psnr(_ptr(a), _ptr(b), n, frame_elems, float(maxvalue), _ptr(out), stream_ptr()),
          "evc_frame_psnr")
    return out


def accept_prefix(score, threshold, higher_is_better=True):
    """score (V, F) float64 -> int32 (V,): frames accepted before the first one failing the threshold."""
    _require_cuda(score)
    V, F = score.shape
    counts = torch.empty((V,), dtype=torch.int32, device=score.device)
    check(load().evc_accept_prefix(_ptr(score.contiguous()), V, F, float(threshold), int(higher_is_better), _ptr(counts),
                                   stream_ptr()), "evc_accept_prefix")
    return counts


def launch_count():
    return int(load().evc_launch_count())
