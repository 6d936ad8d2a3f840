#!/bin/bash
# round 2, call C (2 GPUs): re-run the fixed tests, the two-GPU parity test, and the strong-scaling bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_model_gpu.py tests/test_two_gpu.py tests/test_elementwise_gpu.py tests/test_gemm_gpu.py -m gpu -q -rf 2>&1 | tail -40 > gpurun_out/r2c_pytest.txt; tail -30 gpurun_out/r2c_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c_bench_n2.json'))
    print('N=2 strong', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['videos_per_gpu'], d['scaling'], 'step_frac', round(d['roofline']['step_tensor_frac'],3))
except Exception as e:
    print('N=2 failed', e, open('gpurun_out/r2c_bench_n2.err').read()[-1500:])
PY
