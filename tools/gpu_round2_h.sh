#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_epi_cases.py > gpurun_out/epi_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:evc_gemm -s 3 -c 3 -o gpurun_out/r2h_epi python tools/gpu_epi_cases.py > gpurun_out/epi_ncu.log 2>&1
tail -3 gpurun_out/epi_ncu.log
