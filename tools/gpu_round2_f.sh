#!/bin/bash
mkdir -p gpurun_out
export EVC_LIB=$PWD/build/libevcdiff_prof.so
EPI=1 python tools/gpu_gemm_waits.py 2>&1 | tee gpurun_out/r2f_epi.txt
