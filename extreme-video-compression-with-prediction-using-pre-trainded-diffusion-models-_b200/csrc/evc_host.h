// Host-side helpers shared by the translation units of libevcdiff.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/evcdiff.h"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Resolved through cudaGetDriverEntryPoint so the library does not link libcuda.so and still loads on a
// machine without a driver (the CPU-only test container).
PFN_encodeTiled evc_get_encode_tiled();
int evc_set_error(int code, const char* msg);
// cudaGetLastError() after a launch + launch counter; safe during stream capture.
int evc_check_launch(const char* what);
int evc_num_sms();

static inline int evc_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// 1 unless EVC_PDL=0: kernels of the sampling loop are launched with programmatic dependent launch so that the
// prologue / launch latency of kernel N+1 overlaps the tail of kernel N (also inside captured graphs).
int evc_pdl_enabled();

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t evc_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (evc_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) (void)cudaGetLastError();
  return e;
}
#endif
