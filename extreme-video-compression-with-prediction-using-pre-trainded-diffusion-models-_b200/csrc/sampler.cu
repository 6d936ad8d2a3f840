// Sampler-state kernels (DDPM / DDIM / denoise / PNDM transfer) and the time-embedding tables.
// The sampler state x_t is NCHW fp32 exactly as the reference keeps it (models/__init__.py:289-335,
// models/pndm.py:19-52); each update also refreshes channels [0,C) of the NHWC bf16 UNet input so the
// next UNet evaluation never needs a separate cat / layout pass.
#include "evc_host.h"
#include "evc_ptx.cuh"

namespace evc {

// One thread per pixel: per channel plane the warp reads/writes 128 contiguous bytes of fp32; the bf16
// NHWC row (C <= 32 values) is written as one short run per thread.
__global__ void sampler_update_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                      const float* __restrict__ noise, float* __restrict__ x_out,
                                      __nv_bfloat16* __restrict__ xin, int B, int C, int HW, int Cpad,
                                      evc_step_coef k) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int p = (int)(i % HW);
    const long long base = (long long)b * C * HW + p;
    __nv_bfloat16* row = xin ? xin + i * Cpad : nullptr;
    for (int c = 0; c < C; ++c) {
      const long long idx = base + (long long)c * HW;
      const float xv = x[idx];
      const float ev = eps[idx];
      float r;
      if (k.mode == 0) {
        // x0 = (1/sqrt(a)) * (x - sqrt(1-a)*eps)
        float x0 = __fmul_rn(k.k0, __fsub_rn(xv, __fmul_rn(k.k1, ev)));
        if (k.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        r = __fadd_rn(__fmul_rn(k.c_x0, x0), __fmul_rn(k.c_x, xv));
        if (k.c_eps != 0.f) r = __fadd_rn(r, __fmul_rn(k.c_eps, ev));
        if (k.c_noise != 0.f) r = __fadd_rn(r, __fmul_rn(k.c_noise, noise[idx]));
      } else {
        r = __fsub_rn(xv, __fmul_rn(k.k1, ev));
      }
      x_out[idx] = r;
      if (row) row[c] = __float2bfloat16_rn(r);
    }
  }
}

struct PndmArgs {
  const float* e[4];
  evc_pndm_coef k;
};

__global__ void pndm_update_kernel(const float* __restrict__ x, PndmArgs a, float* __restrict__ x_out,
                                   float* __restrict__ et_out, __nv_bfloat16* __restrict__ xin, int B, int C, int HW,
                                   int Cpad) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int p = (int)(i % HW);
    const long long base = (long long)b * C * HW + p;
    __nv_bfloat16* row = xin ? xin + i * Cpad : nullptr;
    for (int c = 0; c < C; ++c) {
      const long long idx = base + (long long)c * HW;
      float et = __fmul_rn(a.k.w[0], a.e[0][idx]);
      for (int j = 1; j < a.k.n_e; ++j) et = __fadd_rn(et, __fmul_rn(a.k.w[j], a.e[j][idx]));
      et = __fmul_rn(et, a.k.w_scale);
      const float xv = x[idx];
      // x' = x + d * (p*x - q*et)
      float r = __fadd_rn(xv, __fmul_rn(a.k.d, __fsub_rn(__fmul_rn(a.k.p, xv), __fmul_rn(a.k.q, et))));
      if (a.k.clip) r = fminf(fmaxf(r, -1.f), 1.f);
      x_out[idx] = r;
      if (et_out) et_out[idx] = et;
      if (row) row[c] = __float2bfloat16_rn(r);
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ labels, const float* __restrict__ freqs, int L,
                                          int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * half) return;
  const int l = i / half, j = i % half;
  const float arg = __fmul_rn(labels[l], freqs[j]);
  out[(long long)l * dim + j] = sinf(arg);
  out[(long long)l * dim + half + j] = cosf(arg);
  if ((dim & 1) && j == 0) out[(long long)l * dim + dim - 1] = 0.f;
}

// y[l, n] = act_out( sum_k act_in(x[l,k]) * W[n,k] + b[n] ).  One warp per (n, label tile of 8).
constexpr int kLinTile = 8;
__global__ void linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                  float* __restrict__ y, int L, int K, int N, int act_in, int act_out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int l0 = blockIdx.y * kLinTile;
  if (n >= N) return;
  float acc[kLinTile];
#pragma unroll
  for (int t = 0; t < kLinTile; ++t) acc[t] = 0.f;
  const float* w = W + (long long)n * K;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[k];
#pragma unroll
    for (int t = 0; t < kLinTile; ++t) {
      if (l0 + t < L) {
        float xv = x[(long long)(l0 + t) * K + k];
        if (act_in) xv = xv / (1.f + expf(-xv));
        acc[t] = fmaf(xv, wv, acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kLinTile; ++t) acc[t] = warp_sum(acc[t]);
  if (lane == 0) {
    const float bv = bias ? bias[n] : 0.f;
#pragma unroll
    for (int t = 0; t < kLinTile; ++t) {
      if (l0 + t < L) {
        float v = acc[t] + bv;
        if (act_out) v = v / (1.f + expf(-v));
        y[(long long)(l0 + t) * N + n] = v;
      }
    }
  }
}

// Per-frame PSNR in float64 like city_sender.cal_psnr (:255-258): one block per frame, fixed-order tree reduction.
__global__ void frame_psnr_kernel(const float* __restrict__ a, const float* __restrict__ b, long long frame_elems,
                                  double maxvalue, double* __restrict__ psnr) {
  __shared__ double sred[256];
  const long long f = blockIdx.x;
  const float* pa = a + f * frame_elems;
  const float* pb = b + f * frame_elems;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < frame_elems; i += blockDim.x) {
    const double d = (double)pa[i] - (double)pb[i];
    acc += d * d;
  }
  sred[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sred[threadIdx.x] += sred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) psnr[f] = 10.0 * log10((maxvalue * maxvalue) / (sred[0] / (double)frame_elems));
}

// counts[v] = length of the longest prefix of video v's F frames whose score passes the threshold
// (decide_5to5, city_sender.py:353-374: accept frames until the first failure).
__global__ void accept_prefix_kernel(const double* __restrict__ score, int V, int F, double threshold, int higher_is_better,
                                     int* __restrict__ counts) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int n = 0;
  for (; n < F; ++n) {
    const double s = score[(long long)v * F + n];
    const bool ok = higher_is_better ? (s >= threshold) : (s <= threshold);
    if (!ok) break;
  }
  counts[v] = n;
}

}  // namespace evc

using namespace evc;

static inline int grid_px(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)evc_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int evc_sampler_update(const float* x, const float* eps, const float* noise, float* x_out, void* xin,
                                  int32_t B, int32_t C, int32_t HW, int32_t Cpad, const evc_step_coef* coef_host,
                                  evc_stream_t stream) {
  if (!x || !eps || !x_out || !coef_host || B < 1 || C < 1 || HW < 1 || (xin && C > Cpad))
    return evc_set_error(EVC_ERR_INVALID, "evc_sampler_update: bad arguments");
  if (coef_host->mode == 0 && coef_host->c_noise != 0.f && noise == nullptr)
    return evc_set_error(EVC_ERR_INVALID, "evc_sampler_update: noise required");
  cudaError_t le = evc_launch(sampler_update_kernel, dim3(grid_px((long long)B * HW)), dim3(256), 0, (cudaStream_t)stream, 1,
                              x, eps, noise, x_out, reinterpret_cast<__nv_bfloat16*>(xin), (int)B, (int)C, (int)HW, (int)Cpad,
                              *coef_host);
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch("sampler_update_kernel");
}

extern "C" int evc_pndm_update(const float* x, const float* const* e_host, float* x_out, float* et_out, void* xin,
                               int32_t B, int32_t C, int32_t HW, int32_t Cpad, const evc_pndm_coef* coef_host,
                               evc_stream_t stream) {
  if (!x || !e_host || !x_out || !coef_host || coef_host->n_e < 1 || coef_host->n_e > 4 || B < 1 || C < 1 || HW < 1 ||
      (xin && C > Cpad))
    return evc_set_error(EVC_ERR_INVALID, "evc_pndm_update: bad arguments");
  PndmArgs a;
  for (int j = 0; j < 4; ++j) a.e[j] = j < coef_host->n_e ? e_host[j] : nullptr;
  for (int j = 0; j < coef_host->n_e; ++j)
    if (a.e[j] == nullptr) return evc_set_error(EVC_ERR_INVALID, "evc_pndm_update: null eps pointer");
  a.k = *coef_host;
  cudaError_t le = evc_launch(pndm_update_kernel, dim3(grid_px((long long)B * HW)), dim3(256), 0, (cudaStream_t)stream, 1,
                              x, a, x_out, et_out, reinterpret_cast<__nv_bfloat16*>(xin), (int)B, (int)C, (int)HW, (int)Cpad);
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch("pndm_update_kernel");
}

extern "C" int evc_timestep_embedding(const float* labels, const float* freqs, int32_t L, int32_t dim, float* out,
                                      evc_stream_t stream) {
  if (!labels || !freqs || !out || L < 1 || dim < 2) return evc_set_error(EVC_ERR_INVALID, "evc_timestep_embedding: bad arguments");
  const int n = L * (dim / 2);
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(labels, freqs, L, dim, out);
  return evc_check_launch("timestep_embedding_kernel");
}

extern "C" int evc_linear_f32(const float* x, const float* W, const float* b, float* y, int32_t L, int32_t K, int32_t N,
                              int32_t act_in, int32_t act_out, evc_stream_t stream) {
  if (!x || !W || !y || L < 1 || K < 1 || N < 1) return evc_set_error(EVC_ERR_INVALID, "evc_linear_f32: bad arguments");
  const int warps = 8;
  dim3 grid((N + warps - 1) / warps, (L + kLinTile - 1) / kLinTile);
  linear_f32_kernel<<<grid, warps * 32, 0, (cudaStream_t)stream>>>(x, W, b, y, L, K, N, act_in, act_out);
  return evc_check_launch("linear_f32_kernel");
}

extern "C" int evc_frame_psnr(const float* a, const float* b, int32_t n_frames, int64_t frame_elems, double maxvalue,
                              double* psnr, evc_stream_t stream) {
  if (!a || !b || !psnr || n_frames < 1 || frame_elems < 1) return evc_set_error(EVC_ERR_INVALID, "evc_frame_psnr: bad arguments");
  frame_psnr_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(a, b, frame_elems, maxvalue, psnr);
  return evc_check_launch("frame_psnr_kernel");
}

extern "C" int evc_accept_prefix(const double* score, int32_t n_videos, int32_t n_frames, double threshold,
                                 int32_t higher_is_better, int32_t* counts, evc_stream_t stream) {
  if (!score || !counts || n_videos < 1 || n_frames < 1) return evc_set_error(EVC_ERR_INVALID, "evc_accept_prefix: bad arguments");
  accept_prefix_kernel<<<(n_videos + 127) / 128, 128, 0, (cudaStream_t)stream>>>(score, n_videos, n_frames, threshold,
                                                                                higher_is_better, counts);
  return evc_check_launch("accept_prefix_kernel");
}
