#!/bin/bash
# Runs the GPU test-suite file by file (own process + timeout each) and the smoke entry point.
mkdir -p gpurun_out
: > gpurun_out/tests.log
FILES=${@:-tests/test_gemm_gpu.py tests/test_elementwise_gpu.py tests/test_ncsnpp_gpu.py}
for f in $FILES; do
  echo "=== $f" >> gpurun_out/tests.log
  PYTHONUNBUFFERED=1 timeout 420 python -m pytest -v "$f" -q -m gpu --tb=short -p no:cacheprovider >> gpurun_out/tests.log 2>&1
done
echo "=== smoke" >> gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/tests.log 2>&1
grep -E "^===|passed|failed|^FAILED|^E  |smoke\]|Terminated" gpurun_out/tests.log | cut -c1-1200 | tail -80
