"""Where do the GEMM kernel's warps wait?  Needs a library built with -DEVC_GEMM_PROF (tools/build_prof.sh).
Prints, per conv shape of the 128x128 / 64x64 levels at B=46, the share of the MMA-issuing thread's time spent
waiting for operands (TMA -> `full`) and for a free accumulator (epilogue -> `tempty`)."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200")]
import torch
from evcdiff import ops, _lib

lib = _lib.load()
HAVE_PROF = hasattr(lib, "evc_gemm_prof_read")
if HAVE_PROF:
    lib.evc_gemm_prof_read.argtypes = [C.POINTER(C.c_uint64)]
DEV = "cuda"
B = int(os.environ.get("B", "46"))


def run(name, H, Cins, N, taps=9, resid=False, stats=False, cg=None, reps=5, transposed=False, gn=False, skip=0, bn=None,
        stages=0):
    os.environ["EVC_EXP_SKIP"] = str(skip)  # read by evc_gemm_plan_create of the probe build
    os.environ["EVC_EXP_STAGES"] = str(stages)
    if skip:
        name += f" skip={skip}"
    if stages:
        name += f" stages={stages}"
    segs = [(torch.randn(B, H, H, c, device=DEV).to(torch.bfloat16), taps) for c in Cins]
    K = sum(taps * c for c in Cins)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(B, H, H, N, device=DEV, dtype=torch.bfloat16)
    if transposed:  # V^T of the attention blocks: (B, C, H*W), epilogue writes column-major
        out = torch.empty(B, N, H * H, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(B, H, H, N, device=DEV).to(torch.bfloat16) if resid else None
    st = torch.zeros(B, N, 2, device=DEV, dtype=torch.int64) if stats else None
    gnd = None
    if gn:  # Conv_0 -> GroupNorm -> AdaGN -> SiLU in one launch (tickets and statistics zeroed before every launch)
        st = torch.zeros(B, N, 2, device=DEV, dtype=torch.int64)
        gnd = dict(ss=torch.randn(2 * N, device=DEV) * 0.1, ticket=torch.zeros(B, dtype=torch.int32, device=DEV), eps=1e-5,
                   groups=32, adagn=True)
    if gn:
        plan = ops.GemmPlan(segs, w, out, 0, out_ld=N, bias=torch.zeros(N, device=DEV), stats=st, cta_group=cg, gn=gnd, bn=bn)
    elif transposed:
        plan = ops.GemmPlan(segs, w, out, _lib.EVC_OUT_BF16_T, out_ld=H * H, out_bs=N * H * H, bias=torch.zeros(N, device=DEV))
    else:
        plan = ops.GemmPlan(segs, w, out, 0, out_ld=N, bias=torch.zeros(N, device=DEV), resid=r, resid_ld=N if resid else 0,
                            alpha=1.0, stats=st, cta_group=cg, bn=bn)

    def launch():
        if gn:
            gnd["ticket"].zero_()
            st.zero_()
        plan.launch()
    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 16)()
    if HAVE_PROF:
        lib.evc_gemm_prof_read(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = plan.flops / ms / 1e9
    if not HAVE_PROF:
        print(f"{name:34s} cg={plan.cta_group} {ms:7.3f} ms {tf:7.1f} TF", flush=True)
        return
    lib.evc_gemm_prof_read(buf)
    v = [float(x) for x in buf]
    ncta = max(v[7], 1.0)
    print(f"{name:34s} cg={plan.cta_group} {ms:7.3f} ms {tf:7.1f} TF | {v[0] / ncta:9.0f} cyc/CTA = {v[0] / ncta / ms / 1e3:5.0f} MHz | MMA thread: operands-wait {100 * v[1] / v[0]:5.1f}% "
          f"accumulator-wait {100 * v[2] / v[0]:5.1f}% issue {100 * (v[0] - v[1] - v[2]) / v[0]:5.1f}% | producer waits for a free "
          f"stage {100 * v[4] / v[3]:5.1f}% | epilogue: waits for the accumulator {100 * v[6] / v[5]:5.1f}%, prefetch+bar.sync "
          f"{100 * v[8] / v[5]:5.1f}%, tcgen05.ld {100 * v[9] / v[5]:5.1f}%, math+stores+stats {100 * v[10] / v[5]:5.1f}% "
          f"(of which: wait for a free staging buffer {100 * v[11] / v[5]:5.1f}%, smem writes + fence + TMA issue "
          f"{100 * v[12] / v[5]:5.1f}%, statistics {100 * v[13] / v[5]:5.1f}%) | epilogue thread {v[5] / ncta:9.0f} cyc", flush=True)


if os.environ.get("SKIP"):
    # operand-delivery experiments of the probe build: which operand's loads bound the K loop?
    for sk in (0, 1, 4, 2, 6):
        run("128^2 192->192", 128, [192], 192, skip=sk)
    for sk in (0, 1, 4, 2, 6):
        run("128^2 384->192", 128, [192, 192], 192, skip=sk)
    for sk in (0, 1, 4, 2, 6):
        run("64^2 384->384", 64, [384], 384, skip=sk)
    for sk in (0, 1, 4, 2, 6):
        run("32^2 384->384", 32, [384], 384, skip=sk)
    for sk in (0, 4, 2, 6):
        run("8^2 768->768", 8, [768], 768, skip=sk)
    for sk in (0, 4, 2, 6):
        run("32^2 NIN 384->384 +resid+stats", 32, [384], 384, taps=1, resid=True, stats=True, skip=sk)
    sys.exit(0)
if os.environ.get("SKIP2"):
    # 8: no MMAs (pure operand delivery), 16: epilogue only hands the accumulator back, stages: pipeline depth
    for kw in (dict(), dict(skip=8), dict(skip=16), dict(skip=6 + 16), dict(skip=8 + 16), dict(stages=3), dict(stages=4),
               dict(skip=6, stages=3), dict(cg=1), dict(cg=1, skip=8 + 16), dict(cg=1, skip=6 + 16), dict(cg=1, skip=16)):
        run("128^2 192->192", 128, [192], 192, **kw)
    for kw in (dict(), dict(skip=8 + 16), dict(skip=6 + 16), dict(skip=16)):
        run("32^2 384->384", 32, [384], 384, **kw)
        run("8^2 768->768", 8, [768], 768, **kw)
        run("128^2 192->64", 128, [192], 64, **kw)
        run("128^2 192->256", 128, [192], 256, **kw)
    sys.exit(0)
if os.environ.get("CLK"):
    for reps in (1, 5, 40):
        for kw in (dict(), dict(skip=6 + 16), dict(skip=8 + 16)):
            run("128^2 192->192", 128, [192], 192, reps=reps, **kw)
            run("128^2 192->256", 128, [192], 256, reps=reps, **kw)
    sys.exit(0)
if os.environ.get("EPI"):
    run("128^2 192->192 +stats", 128, [192], 192, stats=True)
    run("128^2 192->192 +resid +stats", 128, [192], 192, stats=True, resid=True)
    for H, Ch in ((32, 384), (16, 576)):
        run(f"{H}^2 NIN {Ch}->{Ch} rows", H, [Ch], Ch, taps=1)
        run(f"{H}^2 NIN {Ch}->{Ch} +stats", H, [Ch], Ch, taps=1, stats=True)
        run(f"{H}^2 NIN {Ch}->{Ch} +resid+stats", H, [Ch], Ch, taps=1, resid=True, stats=True)
        run(f"{H}^2 NIN {Ch}->{2 * Ch} rows", H, [Ch], 2 * Ch, taps=1)
        run(f"{H}^2 NIN {Ch}->{2 * Ch} rows no loads", H, [Ch], 2 * Ch, taps=1, skip=6)
    sys.exit(0)
if os.environ.get("GNFUSE"):
    run("128^2 192->192 +stats", 128, [192], 192, stats=True)
    run("128^2 192->192 fused GN", 128, [192], 192, gn=True)
    run("128^2 192->192 fused GN skip", 128, [192], 192, gn=True, skip=6)
    run("128^2 384->192 fused GN", 128, [192, 192], 192, gn=True)
    run("64^2 192->192 +stats", 64, [192], 192, stats=True)
    run("64^2 192->192 fused GN", 64, [192], 192, gn=True)
    run("64^2 384->384 fused GN", 64, [384], 384, gn=True)
    sys.exit(0)
if os.environ.get("SMALLN"):
    for n in (16, 32, 48):
        run(f"128^2 192->{n} 3x3 cg1", 128, [192], n, cg=1)
        run(f"128^2 192->{n} 3x3 cg2", 128, [192], n, cg=2)
    sys.exit(0)
if os.environ.get("NIN"):
    for H, Ch in ((32, 384), (16, 576), (8, 768)):
        run(f"{H}^2 NIN {Ch}->{Ch} rows", H, [Ch], Ch, taps=1)
        run(f"{H}^2 NIN {Ch}->{Ch} +resid", H, [Ch], Ch, taps=1, resid=True)
        run(f"{H}^2 NIN {Ch}->{Ch} transposed", H, [Ch], Ch, taps=1, transposed=True)
        run(f"{H}^2 NIN {Ch}->{2 * Ch} rows", H, [Ch], 2 * Ch, taps=1)
    sys.exit(0)
if os.environ.get("NSWEEP"):
    # does the tensor pipe run faster with a wider N tile?  (shared-memory bandwidth model, profiles/r01_notes.md)
    for n in (64, 96, 128, 192, 256):
        run(f"128^2 192->{n} 3x3", 128, [192], n)
        run(f"128^2 192->{n} 3x3 cg1", 128, [192], n, cg=1)
    for n in (192, 256):
        run(f"128^2 1x1 K=1728 ->{n}", 128, [1728], n, taps=1)
    sys.exit(0)
run("128^2 192->192", 128, [192], 192)
run("128^2 192->192 +stats", 128, [192], 192, stats=True)
run("128^2 192->192 +resid", 128, [192], 192, resid=True)
run("128^2 192->192 +stats cg1", 128, [192], 192, stats=True, cg=1)
run("128^2 384->192 +stats", 128, [192, 192], 192, stats=True)
run("128^2 192->192 + 1x1 skip", 128, [192], 192, resid=True)
run("64^2 192->192 +stats", 64, [192], 192, stats=True)
run("64^2 384->384 +stats", 64, [384], 384, stats=True)
run("32^2 384->384 +stats", 32, [384], 384, stats=True)
run("32^2 NIN 384->384 1x1", 32, [384], 384, taps=1)
run("16^2 576->576 +stats", 16, [576], 576, stats=True)
run("8^2 768->768 +stats", 8, [768], 768, stats=True)
