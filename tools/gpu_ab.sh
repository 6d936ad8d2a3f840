#!/bin/bash
# A/B on the SAME box (boxes differ by ~5 % under the power cap): alternate the variants twice.
mkdir -p gpurun_out
for rep in 1 2; do
  for cg in 1 2; do
    EVC_GEMM_CTA_GROUP=$cg python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ab_cg${cg}_$rep.json 2> gpurun_out/ab_cg${cg}_$rep.err
    python - <<PY
import json
d=json.load(open('gpurun_out/ab_cg${cg}_$rep.json'))
print('cg=$cg rep=$rep', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'clk', d['clocks']['sm_mhz'])
PY
  done
done
