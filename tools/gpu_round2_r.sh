#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 | tee gpurun_out/r2r_pytest.txt
python bench.py --steps 3 --warmup 3 --gpu-eager-context > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err; tail -c 2500 gpurun_out/r2r_bench_default.json; tail -3 gpurun_out/r2r_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err; cat gpurun_out/r2r_bench_reference.json | cut -c1-900
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
