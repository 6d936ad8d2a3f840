#!/bin/bash
# Same-box sweep of one environment knob over several values (2 alternating repetitions): prints frames/s and the
# per-kernel-kind milliseconds of one evaluation.  usage: gpu_sweep_env.sh VAR v1 v2 v3 ... [-- bench args]
mkdir -p gpurun_out
VAR=$1; shift
VALS=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do VALS+=("$1"); shift; done
[ "$1" == "--" ] && shift
for rep in 1 2; do
  for v in "${VALS[@]}"; do
    env $VAR=$v python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/sw_${v}_$rep.json 2> gpurun_out/sw_${v}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sw_${v}_$rep.json'))
    print('$VAR=$v rep=$rep', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$VAR=$v failed', e, open('gpurun_out/sw_${v}_$rep.err').read()[-600:])
PY
  done
done
