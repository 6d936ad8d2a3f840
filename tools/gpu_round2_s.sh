#!/bin/bash
mkdir -p gpurun_out
for v in 8 16 32; do echo "== EVC_GN_STREAM=$v"; EVC_GN_STREAM=$v python tools/gpu_copy_ceiling.py 2>&1 | cut -c1-200; done | tee gpurun_out/r2s_gn_stream2.txt
