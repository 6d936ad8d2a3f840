#!/bin/bash
# the batch sizes a rank gets when bench.py --gpus 4 shards the 46 videos (12 / 11) were never run in round 2 (N = 2 and 8 were):
# parity test of every share size + one-GPU bench lines at those batch sizes
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 400 python -m pytest -q tests/test_full_model_gpu.py -m gpu --tb=short -p no:cacheprovider -k "shares or b46" > gpurun_out/r2y_shares_tests.log 2>&1
tail -5 gpurun_out/r2y_shares_tests.log
for v in 12 11; do
  python bench.py --videos $v --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_b$v.json 2> gpurun_out/r2y_bench_b$v.err || tail -5 gpurun_out/r2y_bench_b$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2y_bench_b$v.json'))
print('videos', $v, round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],1), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3))
PY
done
