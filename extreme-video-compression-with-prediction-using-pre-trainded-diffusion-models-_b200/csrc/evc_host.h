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
