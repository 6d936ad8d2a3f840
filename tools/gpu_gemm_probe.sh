#!/bin/bash
# First-contact probe for the tcgen05 GEMM kernel on a B200 box: every case in its own process under `timeout`
# so a hang or a fault in one case cannot take the others (or the box) with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/probe_gpu.txt 2>&1
ids=$(python -m pytest tests/test_gemm_gpu.py --collect-only -q 2>/dev/null | grep "::" )
echo "$ids" > gpurun_out/probe_ids.txt
for id in $ids; do
  echo "=== $id" >> gpurun_out/probe.log
  timeout 120 python -m pytest "$id" -x -q -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/probe.log
  echo "exit=$?" >> gpurun_out/probe.log
done
grep -E "^===|passed|failed|rel-L2|Error|error|exit=" gpurun_out/probe.log | tail -80
