#!/bin/bash
# strong-scaling bench on N GPUs of one box (BASELINE configs[1]: 46 videos sharded by video); tools/gpu_scale.sh <N> <tag> [bench args]
# e.g. BASELINE configs[2]: tools/gpu_scale.sh 8 c3 --videos 512 --sampler fpndm --subsample 20
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 3 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${tag}_scale_n$N.json 2> gpurun_out/${tag}_scale_n$N.err
fi
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_scale_n$N.json') if l.startswith('{')][-1])
    print('N=$N', d['scaling'], round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['videos_per_gpu'], 'ms/step', round(d['ms_per_step'],1), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('N=$N failed', e, open('gpurun_out/${tag}_scale_n$N.err').read()[-2000:])
PY
