#!/bin/bash
# correctness of the TMA-store epilogue, then same-box sweep of the knob (0 = per-thread stores, 1 / 2 = staging buffers)
PYTHONUNBUFFERED=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_ncsnpp_gpu.py tests/test_unet_plain_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8
bash tools/gpu_sweep_env.sh EVC_GEMM_TMA_STORE 0 1 2
