#!/bin/bash
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 150 python -m pytest -v tests/test_attn_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/attn.log 2>&1
echo "attn rc=$?"
grep -E "PASSED|FAILED|^E  |passed|failed|Terminated" gpurun_out/attn.log | cut -c1-300 | head -40
