#!/bin/bash
# A/B on the SAME box (boxes differ by ~5 % under the power cap): alternate the variants twice.
# usage: gpu_ab.sh ENVVAR valueA valueB
mkdir -p gpurun_out
VAR=${1:-EVC_GEMM_CTA_GROUP}; A=${2:-1}; B=${3:-2}
for rep in 1 2; do
  for v in $A $B; do
    env $VAR=$v python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-json gpurun_out/ab_prof_${v}.json > gpurun_out/ab_${v}_$rep.json 2> gpurun_out/ab_${v}_$rep.err
    tail -2 gpurun_out/ab_${v}_$rep.err
    python - <<PY
import json
d=json.load(open('gpurun_out/ab_${v}_$rep.json'))
print('$VAR=$v rep=$rep', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'clk', d['clocks']['sm_mhz'])
PY
  done
done
