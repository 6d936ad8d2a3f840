#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for lib in r2i v1 v4 v41 v5; do
  echo "== $lib"
  EVC_LIB=$PWD/build/libevcdiff_$lib.so GNFUSE=1 python tools/gpu_gemm_waits.py 2>&1 | cut -c1-90 | grep fused
done
done | tee gpurun_out/r2l_gnfuse2.txt
