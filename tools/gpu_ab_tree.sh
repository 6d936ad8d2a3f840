#!/bin/bash
# Same-box A/B of two source trees (e.g. the previous commit exported to _ab_old/ and built there vs. the working tree).
# usage: gpu_ab_tree.sh [other_tree] [bench args...]
mkdir -p gpurun_out
OTHER=${1:-_ab_old}; shift
ROOT=$PWD
for rep in 1 2; do
  for t in $OTHER .; do
    tag=$(echo $t | tr -c 'a-zA-Z0-9\n' '_')
    (cd $t && python bench.py --steps 2 --warmup 2 --no-cpu-baseline "$@" > $ROOT/gpurun_out/abt_${tag}_$rep.json 2> $ROOT/gpurun_out/abt_${tag}_$rep.err)
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/abt_${tag}_$rep.json'))
    print('tree=$t rep=$rep', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('tree=$t failed', e, open('gpurun_out/abt_${tag}_$rep.err').read()[-800:])
PY
  done
done
