"""Measurement only (never on the product path): cuBLAS bf16 GEMM throughput at the shapes of the dominant implicit-GEMM
launches, as a same-box yardstick for evc_gemm_kernel.  A is the materialised im2col matrix the
product never builds; cuBLAS reads it from HBM once, so this is an upper bound for a library GEMM at this N."""
import json
import sys

import torch

dev = torch.device("cuda", 0)
out = []
for name, M, N, K in [("conv3x3 384->192 @128^2 B=46", 753664, 192, 3456), ("conv3x3 192->192 @128^2 B=46", 753664, 192, 1728),
                      ("conv3x3 384->384 @64^2 B=46", 188416, 384, 3456), ("conv3x3 768->768 @8^2 B=46", 2944, 768, 6912),
                      ("conv3x3 768->768 @8^2 B=6", 384, 768, 6912), ("square 8192", 8192, 8192, 8192)]:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        c = a @ w.t()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            c = a @ w.t()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 4)
    out.append(dict(shape=name, M=M, N=N, K=K, ms=best, tflops=2.0 * M * N * K / best / 1e9))
    print(json.dumps(out[-1]), flush=True)
    del a, w, c
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
