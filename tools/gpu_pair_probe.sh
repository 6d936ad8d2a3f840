#!/bin/bash
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 150 python -m pytest -v tests/test_gemm_pair_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pair.log 2>&1
echo "pair rc=$?"
grep -E "PASSED|FAILED|^E  |passed|failed|Terminated" gpurun_out/pair.log | cut -c1-400 | head -40
