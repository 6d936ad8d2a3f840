// Sampler-state kernels (DDPM / DDIM / denoise / PNDM transfer) and the time-embedding tables.
// The sampler state x_t is NCHW fp32 exactly as the reference keeps it (models/__init__.py:289-335,
// models/pndm.py:19-52); each update also refreshes channels [0,C) of the NHWC bf16 UNet input so the
// next UNet evaluation never needs a separate cat / layout pass.
#include "evc_host.h"
#include "evc_ptx.cuh"

namespace evc {

// The update arithmetic, expression for expression as the reference evaluates it in fp32 (no fused multiply-adds).
__device__ __forceinline__ float step_value(const evc_step_coef& k, float xv, float ev, float nz) {
  if (k.mode == 0) {
    // x0 = (1/sqrt(a)) * (x - sqrt(1-a)*eps)
    float x0 = __fmul_rn(k.k0, __fsub_rn(xv, __fmul_rn(k.k1, ev)));
    if (k.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float r = __fadd_rn(__fmul_rn(k.c_x0, x0), __fmul_rn(k.c_x, xv));
    if (k.c_eps != 0.f) r = __fadd_rn(r, __fmul_rn(k.c_eps, ev));
    if (k.c_noise != 0.f) r = __fadd_rn(r, __fmul_rn(k.c_noise, nz));
    return r;
  }
  return __fsub_rn(xv, __fmul_rn(k.k1, ev));
}

struct PndmArgs {
  const float* e[4];
  evc_pndm_coef k;
};

__device__ __forceinline__ float pndm_value(const evc_pndm_coef& k, float xv, const float (&e)[4], float& et) {
  et = __fmul_rn(k.w[0], e[0]);
#pragma unroll
  for (int j = 1; j < 4; ++j)
    if (j < k.n_e) et = __fadd_rn(et, __fmul_rn(k.w[j], e[j]));
  et = __fmul_rn(et, k.w_scale);
  // x' = x + d * (p*x - q*et)
  float r = __fadd_rn(xv, __fmul_rn(k.d, __fsub_rn(__fmul_rn(k.p, xv), __fmul_rn(k.q, et))));
  if (k.clip) r = fminf(fmaxf(r, -1.f), 1.f);
  return r;
}

// Vectorised form (HW % 4 == 0, 16-byte aligned planes): a thread owns 4 consecutive pixels of one sample.  Per
// channel plane it moves float4s (a warp touches 512 contiguous bytes per load / store instruction), and it writes the
// bf16 NHWC UNet-input rows of its pixels as whole 16-byte chunks: channels [0, CP) with CP = C rounded up to 8, pad
// channels zero (the conditioning frames start at channel CP, see evc_sampler_update in evcdiff.h).
// PNDM = false: DDPM / DDIM / denoise update; PNDM = true: transfer with the fused multistep combination.
template <int CP, bool PNDM>
__global__ void __launch_bounds__(256) state_update_vec_kernel(const float* x, const float* __restrict__ eps,
                                                               const float* __restrict__ noise, PndmArgs pa,
                                                               float* x_out, float* __restrict__ et_out,
                                                               __nv_bfloat16* __restrict__ xin, int B, int C, int HW,
                                                               int Cpad, evc_step_coef k) {
  pdl_wait();
  pdl_trigger();
  const long long groups = (long long)B * HW / 4;
  const bool use_noise = !PNDM && k.mode == 0 && k.c_noise != 0.f;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long long)gridDim.x * blockDim.x) {
    const long long pix = gi * 4;  // first of 4 consecutive pixels (same sample: HW % 4 == 0)
    const int b = (int)(pix / HW);
    const int p = (int)(pix % HW);
    const long long base = (long long)b * C * HW + p;
    uint32_t packed[4][CP / 2];  // [pixel][channel pair]
#pragma unroll
    for (int c = 0; c < CP; c += 2) {
      float r[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = c + h;
        if (cc < C) {
          const long long idx = base + (long long)cc * HW;
          const float4 xv = *reinterpret_cast<const float4*>(x + idx);  // x_out may alias x: plain load, no restrict
          float xa[4] = {xv.x, xv.y, xv.z, xv.w};
          if (PNDM) {
            float ea[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < pa.k.n_e) {
                const float4 ev = __ldg(reinterpret_cast<const float4*>(pa.e[j] + idx));
                ea[0][j] = ev.x; ea[1][j] = ev.y; ea[2][j] = ev.z; ea[3][j] = ev.w;
              } else {
                ea[0][j] = ea[1][j] = ea[2][j] = ea[3][j] = 0.f;
              }
            }
            float et[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) r[h][q] = pndm_value(pa.k, xa[q], ea[q], et[q]);
            if (et_out) *reinterpret_cast<float4*>(et_out + idx) = make_float4(et[0], et[1], et[2], et[3]);
          } else {
            const float4 ev = __ldg(reinterpret_cast<const float4*>(eps + idx));
            float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_noise) nv = __ldg(reinterpret_cast<const float4*>(noise + idx));
            r[h][0] = step_value(k, xa[0], ev.x, nv.x);
            r[h][1] = step_value(k, xa[1], ev.y, nv.y);
            r[h][2] = step_value(k, xa[2], ev.z, nv.z);
            r[h][3] = step_value(k, xa[3], ev.w, nv.w);
          }
          *reinterpret_cast<float4*>(x_out + idx) = make_float4(r[h][0], r[h][1], r[h][2], r[h][3]);
        } else {
          r[h][0] = r[h][1] = r[h][2] = r[h][3] = 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) packed[q][c / 2] = pack_bf16x2(r[0][q], r[1][q]);
    }
    if (xin != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4* row = reinterpret_cast<uint4*>(xin + (pix + q) * Cpad);
#pragma unroll
        for (int v = 0; v < CP / 8; ++v)
          row[v] = make_uint4(packed[q][4 * v], packed[q][4 * v + 1], packed[q][4 * v + 2], packed[q][4 * v + 3]);
      }
    }
  }
}

// Scalar form for shapes the vectorised kernel does not take (HW % 4 != 0, misaligned views, C > 32): one thread per
// pixel; same row convention (channels [C, CP) are written as zero).
template <bool PNDM>
__global__ void state_update_scalar_kernel(const float* x, const float* __restrict__ eps,
                                           const float* __restrict__ noise, PndmArgs pa, float* x_out,
                                           float* __restrict__ et_out, __nv_bfloat16* __restrict__ xin, int B, int C, int HW,
                                           int Cpad, int CP, evc_step_coef k) {
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
      float r;
      if (PNDM) {
        float e[4] = {0.f, 0.f, 0.f, 0.f}, et;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < pa.k.n_e) e[j] = pa.e[j][idx];
        r = pndm_value(pa.k, x[idx], e, et);
        if (et_out) et_out[idx] = et;
      } else {
        r = step_value(k, x[idx], eps[idx], (k.mode == 0 && k.c_noise != 0.f) ? noise[idx] : 0.f);
      }
      x_out[idx] = r;
      if (row) row[c] = __float2bfloat16_rn(r);
    }
    if (row)
      for (int c = C; c < CP && c < Cpad; ++c) row[c] = __float2bfloat16_rn(0.f);
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

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool PNDM>
static int launch_state_update(const float* x, const float* eps, const float* noise, const PndmArgs& pa, float* x_out,
                               float* et_out, void* xin, int B, int C, int HW, int Cpad, const evc_step_coef& k,
                               cudaStream_t stream, const char* what) {
  const int CP = (C + 7) / 8 * 8;
  if (xin != nullptr && (CP > Cpad || (Cpad % 8) != 0))
    return evc_set_error(EVC_ERR_INVALID, "state update: the UNet-input row must hold roundup8(C) channels (Cpad % 8 == 0)");
  __nv_bfloat16* xi = reinterpret_cast<__nv_bfloat16*>(xin);
  bool vec = (HW % 4) == 0 && CP <= 32 && aligned16(x) && aligned16(x_out) && aligned16(xin) &&
             (et_out == nullptr || aligned16(et_out));
  if (PNDM) {
    for (int j = 0; j < pa.k.n_e; ++j) vec = vec && aligned16(pa.e[j]);
  } else {
    vec = vec && aligned16(eps) && (noise == nullptr || aligned16(noise));
  }
  cudaError_t le;
  if (vec) {
    const dim3 grid(grid_px((long long)B * HW / 4));
    auto go = [&](auto kern) {
      return evc_launch(kern, grid, dim3(256), 0, stream, 1, x, eps, noise, pa, x_out, et_out, xi, B, C, HW, Cpad, k);
    };
    if (CP <= 8) le = go(state_update_vec_kernel<8, PNDM>);
    else if (CP <= 16) le = go(state_update_vec_kernel<16, PNDM>);
    else if (CP <= 24) le = go(state_update_vec_kernel<24, PNDM>);
    else le = go(state_update_vec_kernel<32, PNDM>);
  } else {
    le = evc_launch(state_update_scalar_kernel<PNDM>, dim3(grid_px((long long)B * HW)), dim3(256), 0, stream, 1, x, eps,
                    noise, pa, x_out, et_out, xi, B, C, HW, Cpad, CP, k);
  }
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch(what);
}

extern "C" int evc_sampler_update(const float* x, const float* eps, const float* noise, float* x_out, void* xin,
                                  int32_t B, int32_t C, int32_t HW, int32_t Cpad, const evc_step_coef* coef_host,
                                  evc_stream_t stream) {
  if (!x || !eps || !x_out || !coef_host || B < 1 || C < 1 || HW < 1 || (xin && C > Cpad))
    return evc_set_error(EVC_ERR_INVALID, "evc_sampler_update: bad arguments");
  if (coef_host->mode == 0 && coef_host->c_noise != 0.f && noise == nullptr)
    return evc_set_error(EVC_ERR_INVALID, "evc_sampler_update: noise required");
  PndmArgs pa;
  memset(&pa, 0, sizeof(pa));
  return launch_state_update<false>(x, eps, noise, pa, x_out, nullptr, xin, B, C, HW, Cpad, *coef_host,
                                    (cudaStream_t)stream, "sampler_update_kernel");
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
  evc_step_coef k;
  memset(&k, 0, sizeof(k));
  return launch_state_update<true>(x, nullptr, nullptr, a, x_out, et_out, xin, B, C, HW, Cpad, k, (cudaStream_t)stream,
                                   "pndm_update_kernel");
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
