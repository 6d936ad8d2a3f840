#!/bin/bash
# correctness of the TMA epilogue paths, then same-box sweep of one knob
# usage: gpu_tma_epi.sh VAR v1 v2 ...
PYTHONUNBUFFERED=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_ncsnpp_gpu.py tests/test_unet_plain_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8
bash tools/gpu_sweep_env.sh "$@"
