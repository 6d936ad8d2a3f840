/* Minimal C host for the C ABI of libevcdiff.so (include/evcdiff.h): what a cgo / JNI / N-API binding would call.
 * It creates the launch plan of one 3x3 convolution (192 -> 192 channels on a 128x128 image) from caller-owned device
 * buffers and enqueues it on a stream.  Device memory comes from the CUDA runtime here; the library itself never
 * allocates.  Build (after `python __graft_entry__.py build`):
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -L<pkg>/evcdiff/lib -levcdiff -lcudart -o c_abi_demo
 * Without -DEVC_DEMO_WITH_CUDART the file only checks that the header is valid C and that the symbols link. */
#include <stdio.h>
#include <string.h>

#include "evcdiff.h"

#ifdef EVC_DEMO_WITH_CUDART
#include <cuda_runtime_api.h>
#endif

int main(void) {
  printf("evcdiff ABI version %d, sizeof(evc_gemm_desc) = %lld (library: %lld)\n", evc_version(),
         (long long)sizeof(evc_gemm_desc), (long long)evc_struct_size(1));
  if ((long long)sizeof(evc_gemm_desc) != (long long)evc_struct_size(1)) return 2;
#ifdef EVC_DEMO_WITH_CUDART
  const int B = 2, H = 128, W = 128, C = 192, N = 192;
  void *x = NULL, *w = NULL, *y = NULL;
  float* bias = NULL;
  cudaMalloc(&x, (size_t)B * H * W * C * 2);
  cudaMalloc(&w, (size_t)N * 9 * C * 2);
  cudaMalloc(&y, (size_t)B * H * W * N * 2);
  cudaMalloc((void**)&bias, N * sizeof(float));
  cudaMemset(x, 0, (size_t)B * H * W * C * 2);
  cudaMemset(w, 0, (size_t)N * 9 * C * 2);
  cudaMemset(bias, 0, N * sizeof(float));
  evc_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.n_seg = 1;
  d.a[0].ptr = x;
  d.a[0].B = B; d.a[0].H = H; d.a[0].W = W; d.a[0].C = C;
  d.a[0].stride_w = C; d.a[0].stride_h = (int64_t)W * C; d.a[0].stride_b = (int64_t)H * W * C;
  d.taps[0] = 9;
  d.w = w; d.w_rows = N; d.w_k = 9 * C; d.w_batches = 1; d.w_row_stride = 9 * C;
  d.B = B; d.H = H; d.W = W; d.bn = 192;
  d.out = y; d.out_mode = EVC_OUT_BF16_ROWS; d.out_ld = N;
  d.bias = bias; d.alpha = 1.0f;
  evc_gemm_plan* plan = NULL;
  if (evc_gemm_plan_create(&d, &plan) != EVC_OK) {
    fprintf(stderr, "plan: %s\n", evc_last_error());
    return 1;
  }
  if (evc_gemm_plan_launch(plan, NULL, NULL) != EVC_OK) {
    fprintf(stderr, "launch: %s\n", evc_last_error());
    return 1;
  }
  cudaDeviceSynchronize();
  printf("conv3x3 192->192 @128x128: %.1f GFLOP enqueued, %lld kernel launches so far\n", evc_gemm_plan_flops(plan) / 1e9,
         (long long)evc_launch_count());
  evc_gemm_plan_destroy(plan);
#endif
  return 0;
}
