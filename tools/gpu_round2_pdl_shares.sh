#!/bin/bash
# programmatic dependent launch at the share sizes of the 2- and 4-GPU strong-scaling runs (23 / 12 videos) and at 46:
# models/loop.py turns it on for <= 8 videos only (decided in round 1 at 1 and 46 videos); same-box A/B with the env override
mkdir -p gpurun_out
EVC_PDL=1 PYTHONUNBUFFERED=1 timeout 300 python -m pytest -q tests/test_full_model_gpu.py -m gpu --tb=short -p no:cacheprovider -k "shares" > gpurun_out/r2p_shares_pdl1_tests.log 2>&1
tail -3 gpurun_out/r2p_shares_pdl1_tests.log
for v in 12 23 46; do
  for rep in 1 2; do
    for pdl in 0 1; do
      [ $v = 46 ] && [ $rep = 2 ] && continue
      EVC_PDL=$pdl python bench.py --videos $v --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/r2p_pdl${pdl}_b${v}_$rep.json 2> gpurun_out/r2p_pdl${pdl}_b${v}_$rep.err || tail -3 gpurun_out/r2p_pdl${pdl}_b${v}_$rep.err
      python - <<PY
import json
d=json.load(open('gpurun_out/r2p_pdl${pdl}_b${v}_$rep.json'))
print('videos', $v, 'EVC_PDL=$pdl rep $rep', round(d['value'],2), 'fps', round(d['ms_per_step'],1), 'ms/step clk', d['clocks']['sm_mhz'])
PY
    done
  done
done
