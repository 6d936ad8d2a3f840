#!/bin/bash
# Whole GPU suite (as the driver runs it) + the split-precision mode's accuracy numbers and throughput.
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/suite.log 2>&1
tail -5 gpurun_out/suite.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision fp32 --profile-json gpurun_out/prof_fp32.json > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python - <<'PY'
import json
for n in ("bf16", "fp32"):
    try:
        j = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"], 2), "fps e2e", round(j["e2e"]["value"], 2), j["roofline"]["ms_per_eval_by_kernel"], j["clocks"])
    except Exception as e:
        print(n, "failed", e, open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
