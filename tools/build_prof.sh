#!/bin/bash
# Builds build/libevcdiff_prof.so: the library with the GEMM wait-time probe and the timing experiments compiled in
# (experiments only, loaded through EVC_LIB=build/libevcdiff_prof.so; the shipped library is never touched).
P=extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --cudart static -shared -DEVC_GEMM_PROF \
  -o build/libevcdiff_prof.so $P/csrc/evc_host.cu $P/csrc/gemm_tc.cu $P/csrc/attn_tc.cu $P/csrc/elementwise.cu $P/csrc/sampler.cu
