#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
  for v in 0 1; do
    for cfgname in b46 b1; do
      if [ $cfgname = b46 ]; then extra=""; else extra="--videos 1 --micro-batch 1"; fi
      env EVC_PDL=$v python bench.py --steps 1 --warmup 1 --no-cpu-baseline $extra > gpurun_out/pdl_${v}_${cfgname}_$rep.json 2> gpurun_out/pdl_${v}_${cfgname}_$rep.err
      tail -2 gpurun_out/pdl_${v}_${cfgname}_$rep.err
      python - <<PY
import json
d=json.load(open('gpurun_out/pdl_${v}_${cfgname}_$rep.json'))
print('EVC_PDL=$v $cfgname rep=$rep', round(d['value'],2), 'fps', round(d['ms_per_step'],1), 'ms/step clk', d['clocks']['sm_mhz'])
PY
    done
  done
done
