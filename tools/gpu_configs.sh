#!/bin/bash
# One short bench line for each BASELINE.json config that fits one GPU (records for profiles/rNN_notes.md).
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/cfg_$name.json 2> gpurun_out/cfg_$name.err; tail -2 gpurun_out/cfg_$name.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/cfg_$name.json'))
    print('$name', round(d['value'],2), 'fps', round(d['ms_per_step'],1), 'ms/step', d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$name FAILED', e)
PY
}
run c3_fpndm20_b512 --videos 512 --micro-batch 64 --sampler fpndm --subsample 20
run c4_ddim10 --sampler ddim --subsample 10
run c4_ddim50 --sampler ddim --subsample 50
run c5_unet_deep_b32 --model unet_deep --videos 32 --micro-batch 32
run c5_unet_deeper_b32 --model unet_deeper --videos 32 --micro-batch 32
run c1_b1 --videos 1 --micro-batch 1
run c2_b6 --videos 6 --micro-batch 6
