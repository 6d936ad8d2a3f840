// HBM-bound kernels of the sampling path: GroupNorm statistics, fused normalise/AdaGN/SiLU, FIR x2
// resampling, row softmax, layout packing.  All activations are NHWC bf16; every thread moves 16-byte
// vectors (8 channels) so a warp touches whole 128-byte lines along the channel axis.
#include <stdlib.h>

#include "evc_host.h"
#include "evc_ptx.cuh"

namespace evc {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
// x * sigmoid(x) with two MUFU ops (ex2, rcp) and no range fix-up code: exp(-x) = 2^(-x*log2(e)); for x -> -inf the
// product is x * rcp(inf) = -0, for x -> +inf it is x * rcp(1) = x.
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}
template <bool SILU, bool PAIR>
__device__ __forceinline__ uint4 gn_affine_act(const uint4& raw, const float (&a)[8], const float (&bb)[8]) {
  float f[8];
  unpack8(raw, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], a[j], bb[j]);
  if (SILU) {
    if (PAIR) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) silu_pair(f[j], f[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
    }
  }
  return pack8(f);
}
// split-precision pair (hi, lo) <-> fp32: x = hi + lo, hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
  float g[8];
  hi = pack8(f);
  float h[8];
  unpack8(hi, h);
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = f[j] - h[j];
  lo = pack8(g);
}
__device__ __forceinline__ void add8(const uint4& lo, float (&f)[8]) {
  float g[8];
  unpack8(lo, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] += g[j];
}

// read-only 16-byte load that allocates in L1 (FIR stencils: every input vector is re-read by 4-9 neighbouring threads)
__device__ __forceinline__ uint4 ld_ro16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// streaming 16-byte load (read once: do not pollute L1)
__device__ __forceinline__ uint4 ld_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: per-(sample, channel) sum and sum of squares.  Deterministic two-level reduction:
// every block writes its partial sums to `partials`, takes a ticket, and the last block of a sample adds the
// partials in a fixed order (no floating-point atomics, so graph replays are bit-identical).
// grid = (chunks, B), block = (nvec, rows) with nvec = C/8 channel vectors.
// ---------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int HW, int C,
                                long long* __restrict__ stats, int c_total, int c_off, int pix_per_block,
                                float* __restrict__ partials, unsigned* __restrict__ tickets, long long x_lo_off) {
  extern __shared__ float sred[];  // [rows][nvec][16]
  __shared__ unsigned s_last;
  pdl_wait();
  pdl_trigger();
  const int nvec = blockDim.x;
  const int v = threadIdx.x;
  const int r = threadIdx.y;
  const int b = blockIdx.y;
  const int chunks = gridDim.x;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  const __nv_bfloat16* xb = x + (long long)b * HW * ldx + v * 8;
  {
    const int rows = blockDim.y;
    int p = p_begin + r;
    for (; p + 3 * rows < p_end; p += 4 * rows) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_nc16(xb + (long long)(p + k * rows) * ldx);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(u[k], f);
        if (x_lo_off != 0) add8(ld_nc16(xb + x_lo_off + (long long)(p + k * rows) * ldx), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += f[j];
          ss[j] = fmaf(f[j], f[j], ss[j]);
        }
      }
    }
    for (; p < p_end; p += rows) {
      float f[8];
      unpack8(ld_nc16(xb + (long long)p * ldx), f);
      if (x_lo_off != 0) add8(ld_nc16(xb + x_lo_off + (long long)p * ldx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        ss[j] = fmaf(f[j], f[j], ss[j]);
      }
    }
  }
  float* mine = sred + ((size_t)r * nvec + v) * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = s[j];
    mine[8 + j] = ss[j];
  }
  __syncthreads();
  float* part = partials + ((long long)b * chunks + blockIdx.x) * (long long)C * 2;
  if (r == 0) {
    for (int rr = 1; rr < blockDim.y; ++rr) {
      const float* o = sred + ((size_t)rr * nvec + v) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += o[j];
        ss[j] += o[8 + j];
      }
    }
    float4* dst = reinterpret_cast<float4*>(part + v * 16);
    dst[0] = make_float4(s[0], ss[0], s[1], ss[1]);
    dst[1] = make_float4(s[2], ss[2], s[3], ss[3]);
    dst[2] = make_float4(s[4], ss[4], s[5], ss[5]);
    dst[3] = make_float4(s[6], ss[6], s[7], ss[7]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_last = (atomicAdd(tickets + b, 1u) == (unsigned)(chunks - 1)) ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  // last block of sample b: fixed-order sum over the chunks, 2*C outputs spread over the whole block
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const float* pb = partials + (long long)b * chunks * (long long)C * 2;
  for (int i = tid; i < 2 * C; i += nthr) {
    float acc = 0.f;
    for (int k = 0; k < chunks; ++k) acc += __ldcg(pb + (long long)k * C * 2 + i);
    stats[((long long)b * c_total + c_off) * 2 + i] = __float2ll_rn(acc * 1048576.f);
  }
  if (tid == 0) tickets[b] = 0u;
}

// ---------------------------------------------------------------------------------------------
// Normalise + (AdaGN | affine) + SiLU over the virtual concat [x0 | x1]; writes one contiguous tensor.
// grid = (chunks, B), block = (nvec, rows): a thread owns 8 fixed channels (its y = x*a + b coefficients live in
// registers) and streams pixels with 4 independent 16-byte loads in flight.  Group statistics are rebuilt per
// block from the per-channel sums: one warp per group, shuffle tree (deterministic).
// ---------------------------------------------------------------------------------------------
template <bool SILU>
__global__ void __launch_bounds__(256, 3) gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                       const __nv_bfloat16* __restrict__ x1, int C1, int HW,
                                                       const long long* __restrict__ stats0,
                                                       const long long* __restrict__ stats1, int groups, float eps,
                                                       const float* __restrict__ ss, int adagn,
                                                       __nv_bfloat16* __restrict__ y, int pix_per_block,
                                                       const __nv_bfloat16* __restrict__ x0_lo,
                                                       const __nv_bfloat16* __restrict__ x1_lo,
                                                       __nv_bfloat16* __restrict__ y_lo) {
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float2 s_ch[2048];  // per-channel (sum, sum of squares)
  pdl_wait();
  pdl_trigger();
  const int C = C0 + C1;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int v = threadIdx.x;  // channel vector
  const int c = v * 8;
  const bool first = (c < C0);
  const int ld = first ? C0 : C1;
  const int rows = blockDim.y;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  const int p0 = p_begin + threadIdx.y;
  const __nv_bfloat16* src = (first ? x0 + c : x1 + (c - C0)) + ((long long)b * HW + p0) * ld;
  __nv_bfloat16* dst = y + ((long long)b * HW + p0) * C + c;
  const long long sstep = (long long)rows * ld, dstep = (long long)rows * C;
  int n = (p0 < p_end) ? (p_end - p0 + rows - 1) / rows : 0;  // pixels this thread handles
  const bool split = (y_lo != nullptr);
  // the first four pixel loads go out before the statistics are touched: their latency hides the prologue
  uint4 cur[4];
  if (!split && n >= 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) cur[k] = ld_nc16(src + k * sstep);
  }
  // prologue, one global round trip: per-channel sums -> shared; the (scale, shift) rows are fetched meanwhile
  for (int cc = tid; cc < C; cc += nthr) {
    const longlong2 st = *reinterpret_cast<const longlong2*>(
        (cc < C0) ? stats0 + ((long long)b * C0 + cc) * 2 : stats1 + ((long long)b * C1 + (cc - C0)) * 2);
    s_ch[cc] = make_float2((float)((double)st.x * (1.0 / 1048576.0)), (float)((double)st.y * (1.0 / 1048576.0)));
  }
  float gam[8], bet[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(ss + c), g1 = *reinterpret_cast<const float4*>(ss + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(ss + C + c), b1 = *reinterpret_cast<const float4*>(ss + C + c + 4);
    gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
    bet[0] = b0.x; bet[1] = b0.y; bet[2] = b0.z; bet[3] = b0.w; bet[4] = b1.x; bet[5] = b1.y; bet[6] = b1.z; bet[7] = b1.w;
  }
  __syncthreads();
  // group statistics: one thread per group, channels summed in index order (deterministic)
  for (int gi = tid; gi < groups; gi += nthr) {
    float s = 0.f, q = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const float2 t = s_ch[gi * cpg + j];
      s += t.x;
      q += t.y;
    }
    const float inv_n = 1.f / ((float)cpg * (float)HW);
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    s_mean[gi] = mean;
    s_rstd[gi] = rsqrtf(var + eps);
  }
  __syncthreads();
  float a[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    a[j] = s_rstd[g] * (adagn ? (1.f + gam[j]) : gam[j]);
    bb[j] = bet[j] - s_mean[g] * a[j];
  }
  if (split) {
    // split-precision path (accuracy mode, not tuned): x = hi + lo in, (hi, lo) out
    const __nv_bfloat16* src_lo = (first ? x0_lo + c : x1_lo + (c - C0)) + ((long long)b * HW + p0) * ld;
    __nv_bfloat16* dst_lo = y_lo + ((long long)b * HW + p0) * C + c;
    for (int p = p0; p < p_end; p += rows) {
      float f[8];
      unpack8(ld_nc16(src), f);
      add8(ld_nc16(src_lo), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(f[j], a[j], bb[j]);
        f[j] = SILU ? t / (1.f + expf(-t)) : t;
      }
      uint4 hi, lo;
      split8(f, hi, lo);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst_lo) = lo;
      src += sstep; src_lo += sstep; dst += dstep; dst_lo += dstep;
    }
    return;
  }
  // software pipeline: the loads of the next four pixels are in flight while the current four are normalised
  while (n >= 4) {
    uint4 nxt[4];
    const bool more = (n >= 8);
    if (more) {
#pragma unroll
      for (int k = 0; k < 4; ++k) nxt[k] = ld_nc16(src + (4 + k) * sstep);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + k * dstep) = gn_affine_act<SILU, true>(cur[k], a, bb);
    src += 4 * sstep;
    dst += 4 * dstep;
    n -= 4;
    if (more) {
#pragma unroll
      for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
    }
  }
  for (; n > 0; --n) {
    *reinterpret_cast<uint4*>(dst) = gn_affine_act<SILU, true>(ld_nc16(src), a, bb);
    src += sstep;
    dst += dstep;
  }
}

// Streaming form of gn_apply_kernel (bf16 mode): many small blocks instead of one resident wave.  A block covers
// `pix_per_block` pixels of one sample (16 or 32 per thread in batches of four 16-byte loads, all four in flight before
// the first is used; no cross-batch prefetch -- the other resident blocks cover the gap) and recomputes the group statistics from the per-channel sums in its prologue (4.5 KB from L2 per
// ~60 KB of payload); six blocks per SM hide that round trip behind each other's streams.  Same arithmetic as
// gn_apply_kernel, bit for bit.  Measured against torch's elementwise kernels on the same tensors
// (tools/gpu_copy_ceiling.py, profiles/r02_notes.md).
template <bool SILU>
__global__ void __launch_bounds__(256, 4) gn_apply_stream_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                              const __nv_bfloat16* __restrict__ x1, int C1, int HW,
                                                              const long long* __restrict__ stats0,
                                                              const long long* __restrict__ stats1, int groups, float eps,
                                                              const float* __restrict__ ss, int adagn,
                                                              __nv_bfloat16* __restrict__ y, int pix_per_block) {
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float2 s_ch[2048];  // per-channel (sum, sum of squares)
  pdl_wait();
  pdl_trigger();
  const int C = C0 + C1;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int c = threadIdx.x * 8;
  const bool first = (c < C0);
  const int ld = first ? C0 : C1;
  const int rows = blockDim.y;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  const int p0 = p_begin + threadIdx.y;
  const __nv_bfloat16* src = (first ? x0 + c : x1 + (c - C0)) + ((long long)b * HW + p0) * ld;
  __nv_bfloat16* dst = y + ((long long)b * HW + p0) * C + c;
  const long long sstep = (long long)rows * ld, dstep = (long long)rows * C;
  // the first batch of pixel loads goes out before the statistics are touched
  uint4 cur[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (p0 + k * rows < p_end) cur[k] = ld_nc16(src + k * sstep);
  for (int cc = tid; cc < C; cc += nthr) {
    const longlong2 st = *reinterpret_cast<const longlong2*>(
        (cc < C0) ? stats0 + ((long long)b * C0 + cc) * 2 : stats1 + ((long long)b * C1 + (cc - C0)) * 2);
    s_ch[cc] = make_float2((float)((double)st.x * (1.0 / 1048576.0)), (float)((double)st.y * (1.0 / 1048576.0)));
  }
  float gam[8], bet[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(ss + c), g1 = *reinterpret_cast<const float4*>(ss + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(ss + C + c), b1 = *reinterpret_cast<const float4*>(ss + C + c + 4);
    gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
    bet[0] = b0.x; bet[1] = b0.y; bet[2] = b0.z; bet[3] = b0.w; bet[4] = b1.x; bet[5] = b1.y; bet[6] = b1.z; bet[7] = b1.w;
  }
  __syncthreads();
  for (int gi = tid; gi < groups; gi += nthr) {  // one thread per group, channels summed in index order (deterministic)
    float s = 0.f, q = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const float2 t = s_ch[gi * cpg + j];
      s += t.x;
      q += t.y;
    }
    const float inv_n = 1.f / ((float)cpg * (float)HW);
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    s_mean[gi] = mean;
    s_rstd[gi] = rsqrtf(var + eps);
  }
  __syncthreads();
  float a[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    a[j] = s_rstd[g] * (adagn ? (1.f + gam[j]) : gam[j]);
    bb[j] = bet[j] - s_mean[g] * a[j];
  }
  for (int p = p0; p < p_end; p += 4 * rows) {
    if (p != p0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (p + k * rows < p_end) cur[k] = ld_nc16(src + k * sstep);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (p + k * rows < p_end) *reinterpret_cast<uint4*>(dst + k * dstep) = gn_affine_act<SILU, true>(cur[k], a, bb);
    src += 4 * sstep;
    dst += 4 * dstep;
  }
}

// ---------------------------------------------------------------------------------------------
// FIR [1,3,3,1] x2 resampling, separable, zero borders (upfirdn2d modes of upsample_2d / downsample_2d).
//   up:   out[2i]   = (x[i-1] + 3 x[i]) / 4,  out[2i+1] = (3 x[i] + x[i+1]) / 4      (per axis)
//   down: out[i]    = (x[2i-1] + 3 x[2i] + 3 x[2i+1] + x[2i+2]) / 8                  (per axis)
// One thread = one output pixel x 8 channels.
// ---------------------------------------------------------------------------------------------
// Loads / stores of one 8-channel vector; in split-precision mode the value is hi + lo and is written back as a pair.
template <bool SPLIT>
__device__ __forceinline__ void fir_load(const __nv_bfloat16* p, long long lo_off, float (&f)[8]) {
  unpack8(ld_ro16(p), f);
  if (SPLIT) add8(ld_ro16(p + lo_off), f);
}
template <bool SPLIT>
__device__ __forceinline__ void fir_store(__nv_bfloat16* p, long long lo_off, const float (&f)[8]) {
  if (SPLIT) {
    uint4 hi, lo;
    split8(f, hi, lo);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + lo_off) = lo;
  } else {
    *reinterpret_cast<uint4*>(p) = pack8(f);
  }
}

// One thread = one INPUT pixel x 8 channels -> the 2x2 output pixels it expands to (3x3 input neighbourhood,
// 9 loads for 4 stores instead of 16).  Horizontal pass first: L = (x[i-1] + 3 x[i]) / 4, R = (3 x[i] + x[i+1]) / 4.
template <bool SPLIT>
__global__ void fir_up_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                              int C, long long x_lo_off, long long y_lo_off) {
  pdl_wait();
  pdl_trigger();
  const int nvec = C / 8;
  const int OW = 2 * W;
  const long long total = (long long)B * H * W * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long t = i / nvec;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const int b = (int)(t / H);
    const __nv_bfloat16* xb = x + (long long)b * H * W * C + v * 8;
    float hl[3][8], hr[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = iy - 1 + r;
      if (yy < 0 || yy >= H) {
#pragma unroll
        for (int j = 0; j < 8; ++j) hl[r][j] = hr[r][j] = 0.f;
        continue;
      }
      const __nv_bfloat16* row = xb + (long long)yy * W * C;
      float c0[8], cm[8], cp[8];
      fir_load<SPLIT>(row + (long long)ix * C, x_lo_off, c0);
      if (ix > 0) fir_load<SPLIT>(row + (long long)(ix - 1) * C, x_lo_off, cm);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) cm[j] = 0.f;
      }
      if (ix + 1 < W) fir_load<SPLIT>(row + (long long)(ix + 1) * C, x_lo_off, cp);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) cp[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hl[r][j] = 0.25f * cm[j] + 0.75f * c0[j];
        hr[r][j] = 0.75f * c0[j] + 0.25f * cp[j];
      }
    }
    float o[8];
    __nv_bfloat16* yb = y + (((long long)b * 2 * H + 2 * iy) * OW + 2 * ix) * C + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * hl[0][j] + 0.75f * hl[1][j];
    fir_store<SPLIT>(yb, y_lo_off, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * hr[0][j] + 0.75f * hr[1][j];
    fir_store<SPLIT>(yb + C, y_lo_off, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.75f * hl[1][j] + 0.25f * hl[2][j];
    fir_store<SPLIT>(yb + (long long)OW * C, y_lo_off, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.75f * hr[1][j] + 0.25f * hr[2][j];
    fir_store<SPLIT>(yb + (long long)OW * C + C, y_lo_off, o);
  }
}

template <bool SPLIT>
__global__ void fir_down_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H,
                                int W, int C, long long x_lo_off, long long y_lo_off) {
  pdl_wait();
  pdl_trigger();
  const int nvec = C / 8;
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)B * OH * OW * nvec;
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long t = i / nvec;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* xb = x + (long long)b * H * W * C + v * 8;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = 2 * oy - 1 + a;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int xx = 2 * ox - 1 + c;
        if (xx < 0 || xx >= W) continue;
        const float w = k[a] * k[c];
        float f[8];
        fir_load<SPLIT>(xb + ((long long)yy * W + xx) * C, x_lo_off, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, f[j], acc[j]);
      }
    }
    fir_store<SPLIT>(y + (((long long)b * OH + oy) * OW + ox) * C + v * 8, y_lo_off, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused res-block prologue of the up / down blocks (layerspp.py:598-611): from ONE read of the block input
// x = [x0 | x1] produce  y_act = FIR(SiLU(AdaGN(x)))  and  y_raw = FIR(x)  (the skip path), instead of
// gn_apply -> fir(h) + fir(x) (three reads, one full-size intermediate).  A thread owns 8 channels of one pixel
// column and walks down a strip of rows with a two-row rolling window of horizontally filtered values, so every
// input vector is loaded once per column neighbourhood (3 loads per input row when up-sampling, 8 per output row
// when down-sampling).  Zero padding applies after the activation, exactly like upfirdn2d on h.
// grid = (column blocks, row strips, B), block = (nvec, PX).
// ---------------------------------------------------------------------------------------------
struct GnFirArgs {
  const __nv_bfloat16 *x0, *x1;
  int C0, C1, H, W;
  const long long *stats0, *stats1;
  int groups;
  float eps;
  const float* ss;
  int adagn;
  __nv_bfloat16 *y_act, *y_raw0, *y_raw1;
  int strip;  // rows per block: input rows (up) / output rows (down)
};

// V channels (8 or 4) per thread: 16- / 8-byte vectors.  V = 4 halves the register footprint (twice the resident warps),
// which is what the load-latency-bound walk needs at the high-resolution levels.
template <int V>
__device__ __forceinline__ void ldv(const __nv_bfloat16* p, float (&f)[V]) {
  if (V == 8) {
    const uint4 u = ld_ro16(p);
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4 % V] = bf16_lo(u.z); f[5 % V] = bf16_hi(u.z); f[6 % V] = bf16_lo(u.w); f[7 % V] = bf16_hi(u.w);
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  }
}
template <int V>
__device__ __forceinline__ void stv(__nv_bfloat16* p, const float (&f)[V]) {
  if (V == 8) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4 % V], f[5 % V]); u.w = pack_bf16x2(f[6 % V], f[7 % V]);
    *reinterpret_cast<uint4*>(p) = u;
  } else {
    uint2 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
}
template <int V>
__device__ __forceinline__ void actv(const float (&x)[V], const float (&a)[V], const float (&bb)[V], float (&y)[V]) {
#pragma unroll
  for (int j = 0; j < V; ++j) y[j] = fmaf(x[j], a[j], bb[j]);
#pragma unroll
  for (int j = 0; j < V; j += 2) silu_pair(y[j], y[j + 1]);
}

template <bool UP, int V>
__global__ void __launch_bounds__(256, (V == 8) ? 2 : 3) gn_fir_kernel(const GnFirArgs g) {
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float2 s_ch[2048];
  pdl_wait();
  pdl_trigger();
  const int C = g.C0 + g.C1, H = g.H, W = g.W;
  const int b = blockIdx.z;
  const int cpg = C / g.groups;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int c = threadIdx.x * V;
  // ---- GroupNorm coefficients of this thread's channels (same arithmetic as gn_apply_kernel)
  for (int cc = tid; cc < C; cc += nthr) {
    const longlong2 st = *reinterpret_cast<const longlong2*>(
        (cc < g.C0) ? g.stats0 + ((long long)b * g.C0 + cc) * 2 : g.stats1 + ((long long)b * g.C1 + (cc - g.C0)) * 2);
    s_ch[cc] = make_float2((float)((double)st.x * (1.0 / 1048576.0)), (float)((double)st.y * (1.0 / 1048576.0)));
  }
  float gam[V], bet[V];
#pragma unroll
  for (int j = 0; j < V; j += 4) {
    const float4 g0 = *reinterpret_cast<const float4*>(g.ss + c + j);
    const float4 b0 = *reinterpret_cast<const float4*>(g.ss + C + c + j);
    gam[j] = g0.x; gam[j + 1] = g0.y; gam[j + 2] = g0.z; gam[j + 3] = g0.w;
    bet[j] = b0.x; bet[j + 1] = b0.y; bet[j + 2] = b0.z; bet[j + 3] = b0.w;
  }
  __syncthreads();
  for (int gi = tid; gi < g.groups; gi += nthr) {  // blocks can be narrower than the group count
    float s = 0.f, q = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const float2 t = s_ch[gi * cpg + j];
      s += t.x;
      q += t.y;
    }
    const float inv_n = 1.f / ((float)cpg * (float)(H * W));
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    s_mean[gi] = mean;
    s_rstd[gi] = rsqrtf(var + g.eps);
  }
  __syncthreads();
  float a[V], bb[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int gi = (c + j) / cpg;
    a[j] = s_rstd[gi] * (g.adagn ? (1.f + gam[j]) : gam[j]);
    bb[j] = bet[j] - s_mean[gi] * a[j];
  }
  const bool first = (c < g.C0);
  const int Cs = first ? g.C0 : g.C1;      // channels of the source / raw destination tensor of this thread
  const int cs = first ? c : c - g.C0;     // channel offset inside it
  const __nv_bfloat16* xs = (first ? g.x0 : g.x1) + (long long)b * H * W * Cs + cs;
  __nv_bfloat16* yr = first ? g.y_raw0 : g.y_raw1;
  const int px = blockIdx.x * blockDim.y + threadIdx.y;  // input column (up) / output column (down)

  if (UP) {
    if (px >= W) return;
    const int OW = 2 * W;
    const int r0 = blockIdx.y * g.strip, r1 = min(H, r0 + g.strip);
    __nv_bfloat16* ya = g.y_act + (long long)b * 4 * H * W * C + c;
    yr += (long long)b * 4 * H * W * Cs + cs;
    // horizontally filtered row r: L = (x[px-1] + 3 x[px]) / 4, R = (3 x[px] + x[px+1]) / 4, raw and activated
    float pl[V], pr[V], pal[V], par[V];  // previous row
    auto hrow = [&](int r, float (&l)[V], float (&rr)[V], float (&al)[V], float (&ar)[V]) {
      if (r < 0 || r >= H) {
#pragma unroll
        for (int j = 0; j < V; ++j) l[j] = rr[j] = al[j] = ar[j] = 0.f;
        return;
      }
      const __nv_bfloat16* row = xs + ((long long)r * W + px) * Cs;
      float m[V], c0[V], p[V], am[V], ac[V], ap[V];
      const bool hm = px > 0, hp = px + 1 < W;
      ldv<V>(row, c0);
      if (hm) ldv<V>(row - Cs, m);
      if (hp) ldv<V>(row + Cs, p);
      actv<V>(c0, a, bb, ac);
      if (hm) actv<V>(m, a, bb, am);
      if (hp) actv<V>(p, a, bb, ap);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (!hm) m[j] = am[j] = 0.f;
        if (!hp) p[j] = ap[j] = 0.f;
        l[j] = 0.25f * m[j] + 0.75f * c0[j];
        rr[j] = 0.75f * c0[j] + 0.25f * p[j];
        al[j] = 0.25f * am[j] + 0.75f * ac[j];
        ar[j] = 0.75f * ac[j] + 0.25f * ap[j];
      }
    };
    auto emit = [&](int oy, float wp, float wc, const float (&cl)[V], const float (&cr)[V], const float (&cal)[V],
                    const float (&car)[V]) {
      float o[V];
      const long long pix = ((long long)oy * OW + 2 * px);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = wp * pl[j] + wc * cl[j];
      stv<V>(yr + pix * Cs, o);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = wp * pr[j] + wc * cr[j];
      stv<V>(yr + (pix + 1) * Cs, o);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = wp * pal[j] + wc * cal[j];
      stv<V>(ya + pix * C, o);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = wp * par[j] + wc * car[j];
      stv<V>(ya + (pix + 1) * C, o);
    };
    hrow(r0 - 1, pl, pr, pal, par);
    // step r (r0 .. r1): rows r-1 and r are known -> output rows 2r-1 (if r > r0) and 2r (if r < r1)
    for (int r = r0; r <= r1; ++r) {
      float cl[V], cr[V], cal[V], car[V];
      hrow(r, cl, cr, cal, car);
      if (r > r0) emit(2 * r - 1, 0.75f, 0.25f, cl, cr, cal, car);
      if (r < r1) emit(2 * r, 0.25f, 0.75f, cl, cr, cal, car);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        pl[j] = cl[j];
        pr[j] = cr[j];
        pal[j] = cal[j];
        par[j] = car[j];
      }
    }
  } else {
    const int OH = H / 2, OW = W / 2;
    if (px >= OW) return;
    const int o0 = blockIdx.y * g.strip, o1 = min(OH, o0 + g.strip);
    __nv_bfloat16* ya = g.y_act + (long long)b * OH * OW * C + c;
    yr += (long long)b * OH * OW * Cs + cs;
    const bool hm = px > 0, hp = 2 * px + 2 < W;
    const __nv_bfloat16* col = xs + (long long)(2 * px) * Cs;
    // horizontally filtered rows r, r+1 at output column px: (x[2px-1] + 3 x[2px] + 3 x[2px+1] + x[2px+2]) / 8;
    // all eight loads are issued before any arithmetic
    auto hrow2 = [&](int r, float (&h)[V], float (&ah)[V], float (&hh)[V], float (&ahh)[V]) {
      float v[2][4][V];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool ok = (r + k >= 0) && (r + k < H);
        const __nv_bfloat16* row = col + (long long)(r + k) * W * Cs;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const bool in = ok && (t != 0 || hm) && (t != 3 || hp);
          if (in) ldv<V>(row + (t - 1) * Cs, v[k][t]);
          else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[k][t][j] = 0.f;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool ok = (r + k >= 0) && (r + k < H);
        float (&dh)[V] = k ? hh : h;
        float (&dah)[V] = k ? ahh : ah;
#pragma unroll
        for (int j = 0; j < V; ++j) dh[j] = dah[j] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const bool in = ok && (t != 0 || hm) && (t != 3 || hp);
          const float w = (t == 0 || t == 3) ? 0.125f : 0.375f;
          float av[V];
          if (in) {
            actv<V>(v[k][t], a, bb, av);
#pragma unroll
            for (int j = 0; j < V; ++j) {
              dh[j] = fmaf(w, v[k][t][j], dh[j]);
              dah[j] = fmaf(w, av[j], dah[j]);
            }
          }
        }
      }
    };
    float h0[V], ah0[V], h1[V], ah1[V];  // rows 2oy-1 and 2oy
    hrow2(2 * o0 - 1, h0, ah0, h1, ah1);
    for (int oy = o0; oy < o1; ++oy) {
      float h2[V], ah2[V], h3[V], ah3[V];
      hrow2(2 * oy + 1, h2, ah2, h3, ah3);
      float o[V];
      const long long pix = (long long)oy * OW + px;
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = 0.125f * h0[j] + 0.375f * h1[j] + 0.375f * h2[j] + 0.125f * h3[j];
      stv<V>(yr + pix * Cs, o);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = 0.125f * ah0[j] + 0.375f * ah1[j] + 0.375f * ah2[j] + 0.125f * ah3[j];
      stv<V>(ya + pix * C, o);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        h0[j] = h2[j];
        ah0[j] = ah2[j];
        h1[j] = h3[j];
        ah1[j] = ah3[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Down-sampling form of the fused prologue with every input element loaded and activated exactly ONCE (the walk above
// activates it twice: neighbouring output columns share two of their four input columns, and SiLU costs 1.5 MUFU ops).
// A thread owns 8 channels of one INPUT column and walks down the rows: the vertical half of the separable filter runs
// in registers (carry = .125 r[2oy-1] + .375 r[2oy] from the previous step, this step adds .375 r[2oy+1] + .125 r[2oy+2]),
// the horizontal half exchanges the vertically reduced rows (raw and activated, fp32) through a double-buffered
// shared-memory row: one __syncthreads per output row, and all threads share the 2 * OXB * nvb output vectors.
// block = (nvb channel vectors, NC = 2*OXB + 2 input columns incl. the halo); grid = (column blocks * channel chunks,
// row strips, B); dynamic shared memory 2 buffers x 2 planes x NC x (nvb * 8) floats.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(416, 2) gn_fir_down_kernel(const GnFirArgs g, int OXB, int CB, int n_chunks) {
  extern __shared__ float s_x[];
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float2 s_ch[2048];
  pdl_wait();
  pdl_trigger();
  const int C = g.C0 + g.C1, H = g.H, W = g.W, OH = H / 2, OW = W / 2;
  const int b = blockIdx.z;
  const int cpg = C / g.groups;
  const int nvb = blockDim.x, NC = blockDim.y;
  const int vx = threadIdx.x, cx = threadIdx.y;
  const int tid = cx * nvb + vx, nthr = nvb * NC;
  const int chunk = blockIdx.x % n_chunks, xblk = blockIdx.x / n_chunks;
  const int c = chunk * CB + vx * 8;  // first channel of this thread
  // ---- GroupNorm coefficients (same arithmetic as gn_apply_kernel / gn_fir_kernel)
  for (int cc = tid; cc < C; cc += nthr) {
    const longlong2 st = *reinterpret_cast<const longlong2*>(
        (cc < g.C0) ? g.stats0 + ((long long)b * g.C0 + cc) * 2 : g.stats1 + ((long long)b * g.C1 + (cc - g.C0)) * 2);
    s_ch[cc] = make_float2((float)((double)st.x * (1.0 / 1048576.0)), (float)((double)st.y * (1.0 / 1048576.0)));
  }
  float a[8], bb[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(g.ss + c), g1 = *reinterpret_cast<const float4*>(g.ss + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(g.ss + C + c), b1 = *reinterpret_cast<const float4*>(g.ss + C + c + 4);
    a[0] = g0.x; a[1] = g0.y; a[2] = g0.z; a[3] = g0.w; a[4] = g1.x; a[5] = g1.y; a[6] = g1.z; a[7] = g1.w;
    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
  }
  __syncthreads();
  for (int gi = tid; gi < g.groups; gi += nthr) {
    float sm = 0.f, q = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const float2 t = s_ch[gi * cpg + j];
      sm += t.x;
      q += t.y;
    }
    const float inv_n = 1.f / ((float)cpg * (float)(H * W));
    const float mean = sm * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    s_mean[gi] = mean;
    s_rstd[gi] = rsqrtf(var + g.eps);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gi = (c + j) / cpg;
    const float gam = s_rstd[gi] * (g.adagn ? (1.f + a[j]) : a[j]);
    bb[j] = bb[j] - s_mean[gi] * gam;
    a[j] = gam;
  }
  const bool first = (c < g.C0);
  const int Cs = first ? g.C0 : g.C1;
  const int cs = first ? c : c - g.C0;
  const int ox0 = xblk * OXB;
  const int gx = 2 * ox0 - 1 + cx;  // this thread's input column
  const bool col_ok = (gx >= 0) && (gx < W);
  const __nv_bfloat16* xs = (first ? g.x0 : g.x1) + ((long long)b * H * W + (col_ok ? gx : 0)) * Cs + cs;
  const int o0 = blockIdx.y * g.strip, o1 = min(OH, o0 + g.strip);
  const int pitch = nvb * 8;                 // floats per column
  const int plane = NC * pitch;              // floats per (raw | act) plane
  auto load_row = [&](int r) -> uint4 {
    if (!col_ok || r < 0 || r >= H) return make_uint4(0u, 0u, 0u, 0u);
    return ld_nc16(xs + (long long)r * W * Cs);
  };
  // rows outside the image contribute zeros to BOTH outputs (upfirdn2d pads the activated tensor, layerspp.py:604-611)
  auto row_ok = [&](int r) { return col_ok && r >= 0 && r < H; };
  float cr[8], ca[8];  // carry: .125 r[2oy-1] + .375 r[2oy], raw and activated
  {
    const uint4 u0 = load_row(2 * o0 - 1), u1 = load_row(2 * o0);
    float r0[8], r1[8], a0[8], a1[8];
    unpack8(u0, r0);
    unpack8(u1, r1);
    const bool k0 = row_ok(2 * o0 - 1), k1 = row_ok(2 * o0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = fmaf(r0[j], a[j], bb[j]);
      a1[j] = fmaf(r1[j], a[j], bb[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      silu_pair(a0[j], a0[j + 1]);
      silu_pair(a1[j], a1[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cr[j] = 0.125f * r0[j] + 0.375f * r1[j];
      ca[j] = (k0 ? 0.125f * a0[j] : 0.f) + (k1 ? 0.375f * a1[j] : 0.f);
    }
  }
  // phase-2 role of this thread: output vector (which plane, output column, channel vector)
  const int items = 2 * OXB * nvb;
  const bool p2 = tid < items;
  const int which = tid / (OXB * nvb);
  const int rem = tid - which * (OXB * nvb);
  const int p2ox = rem / nvb, p2vx = rem - p2ox * nvb;
  const int p2c = chunk * CB + p2vx * 8;
  const bool p2first = p2c < g.C0;
  __nv_bfloat16* p2dst;
  long long p2ld;
  if (which == 0) {  // raw (skip path): FIR(x0) / FIR(x1) as separate tensors
    const int Cd = p2first ? g.C0 : g.C1;
    p2dst = (p2first ? g.y_raw0 : g.y_raw1) + (long long)b * OH * OW * Cd + (p2first ? p2c : p2c - g.C0);
    p2ld = Cd;
  } else {
    p2dst = g.y_act + (long long)b * OH * OW * C + p2c;
    p2ld = C;
  }
  const bool p2ok = p2 && (ox0 + p2ox) < OW;
  uint4 n0 = load_row(2 * o0 + 1), n1 = load_row(2 * o0 + 2);
  int buf = 0;
  for (int oy = o0; oy < o1; ++oy) {
    float r0[8], r1[8], a0[8], a1[8];
    unpack8(n0, r0);
    unpack8(n1, r1);
    const bool k0 = row_ok(2 * oy + 1), k1 = row_ok(2 * oy + 2);
    if (oy + 1 < o1) {  // next step's rows are in flight while this one is reduced and exchanged
      n0 = load_row(2 * oy + 3);
      n1 = load_row(2 * oy + 4);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = fmaf(r0[j], a[j], bb[j]);
      a1[j] = fmaf(r1[j], a[j], bb[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      silu_pair(a0[j], a0[j + 1]);
      silu_pair(a1[j], a1[j + 1]);
    }
    float vr[8], va[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x0 = k0 ? a0[j] : 0.f, x1 = k1 ? a1[j] : 0.f;
      vr[j] = cr[j] + 0.375f * r0[j] + 0.125f * r1[j];
      va[j] = ca[j] + 0.375f * x0 + 0.125f * x1;
      cr[j] = 0.125f * r0[j] + 0.375f * r1[j];
      ca[j] = 0.125f * x0 + 0.375f * x1;
    }
    // column cx, two float4 planes per 8-channel vector (conflict-free 16-byte accesses along vx)
    float* sb = s_x + buf * 2 * plane + cx * pitch + vx * 4;
    *reinterpret_cast<float4*>(sb) = make_float4(vr[0], vr[1], vr[2], vr[3]);
    *reinterpret_cast<float4*>(sb + nvb * 4) = make_float4(vr[4], vr[5], vr[6], vr[7]);
    *reinterpret_cast<float4*>(sb + plane) = make_float4(va[0], va[1], va[2], va[3]);
    *reinterpret_cast<float4*>(sb + plane + nvb * 4) = make_float4(va[4], va[5], va[6], va[7]);
    __syncthreads();
    if (p2ok) {
      const float* sp = s_x + buf * 2 * plane + which * plane + (2 * p2ox) * pitch + p2vx * 4;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float w = (t == 0 || t == 3) ? 0.125f : 0.375f;
        const float4 lo = *reinterpret_cast<const float4*>(sp + t * pitch);
        const float4 hi = *reinterpret_cast<const float4*>(sp + t * pitch + nvb * 4);
        o[0] = fmaf(w, lo.x, o[0]); o[1] = fmaf(w, lo.y, o[1]); o[2] = fmaf(w, lo.z, o[2]); o[3] = fmaf(w, lo.w, o[3]);
        o[4] = fmaf(w, hi.x, o[4]); o[5] = fmaf(w, hi.y, o[5]); o[6] = fmaf(w, hi.z, o[6]); o[7] = fmaf(w, hi.w, o[7]);
      }
      *reinterpret_cast<uint4*>(p2dst + ((long long)oy * OW + ox0 + p2ox) * p2ld) = pack8(o);
    }
    buf ^= 1;
  }
}

__global__ void nearest_up2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H,
                                   int W, int C) {
  const int nvec = C / 8;
  const int OH = 2 * H, OW = 2 * W;
  const long long total = (long long)B * OH * OW * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long t = i / nvec;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    const uint4 u = *reinterpret_cast<const uint4*>(x + (((long long)b * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 8);
    *reinterpret_cast<uint4*>(y + (((long long)b * OH + oy) * OW + ox) * C + v * 8) = u;
  }
}

// ---------------------------------------------------------------------------------------------
// Row softmax: one warp per row, fp32 in, bf16 out.  cols % 4 == 0.
// ---------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, long long rows,
                                    int cols, __nv_bfloat16* __restrict__ P_lo) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + row * cols;
  __nv_bfloat16* p = P + row * cols;
  float m = -INFINITY;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(s + c);
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(s + c);
    sum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(s + c);
    const float e0 = __expf(v.x - m) * inv, e1 = __expf(v.y - m) * inv, e2 = __expf(v.z - m) * inv,
                e3 = __expf(v.w - m) * inv;
    uint2 o;
    o.x = pack_bf16x2(e0, e1);
    o.y = pack_bf16x2(e2, e3);
    *reinterpret_cast<uint2*>(p + c) = o;
    if (P_lo != nullptr) {
      uint2 l;
      l.x = pack_bf16x2(e0 - bf16_lo(o.x), e1 - bf16_hi(o.x));
      l.y = pack_bf16x2(e2 - bf16_lo(o.y), e3 - bf16_hi(o.y));
      *reinterpret_cast<uint2*>(P_lo + row * cols + c) = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32/fp64 -> channel slice of an NHWC bf16 buffer.  One thread per pixel; reads are coalesced per
// channel plane, writes are one short contiguous run per pixel.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_nchw_kernel(const T* __restrict__ src, int B, int C, int HW, float scale, float shift,
                                 __nv_bfloat16* __restrict__ dst, int Cpad, int c_off,
                                 __nv_bfloat16* __restrict__ dst_lo) {
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int p = (int)(i % HW);
    const T* s = src + (long long)b * C * HW + p;
    __nv_bfloat16* d = dst + i * Cpad + c_off;
    // evaluated in the source precision with separate multiply / add, like the reference's `2 * X - 1.`
    for (int c = 0; c < C; ++c) {
      const T v = s[(long long)c * HW] * (T)scale;
      const float f = (float)(v + (T)shift);
      const __nv_bfloat16 h = __float2bfloat16_rn(f);
      d[c] = h;
      if (dst_lo != nullptr) dst_lo[i * Cpad + c_off + c] = __float2bfloat16_rn(f - __bfloat162float(h));
    }
  }
}

__global__ void inverse_transform_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = (x[i] + 1.f) / 2.f;
    y[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

// frames in [0,1] fp32 -> uint8 round(255 x) (round half to even, like numpy / torch round), 4 values per thread
__global__ void frames_to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, long long n) {
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uchar4 o;
    o.x = (unsigned char)fminf(fmaxf(rintf(v.x * 255.f), 0.f), 255.f);
    o.y = (unsigned char)fminf(fmaxf(rintf(v.y * 255.f), 0.f), 255.f);
    o.z = (unsigned char)fminf(fmaxf(rintf(v.z * 255.f), 0.f), 255.f);
    o.w = (unsigned char)fminf(fmaxf(rintf(v.w * 255.f), 0.f), 255.f);
    reinterpret_cast<uchar4*>(y)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    y[i] = (unsigned char)fminf(fmaxf(rintf(x[i] * 255.f), 0.f), 255.f);
  }
}

}  // namespace evc

using namespace evc;

static inline int grid_for(long long work_items, int block) {
  long long g = (work_items + block - 1) / block;
  const long long cap = (long long)evc_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int evc_gn_stats_workspace(int32_t B, int32_t HW, int32_t C, int64_t* bytes) {
  if (!bytes || B < 1 || HW < 1 || C < 8) return evc_set_error(EVC_ERR_INVALID, "evc_gn_stats_workspace: bad arguments");
  // tickets (B * 4 bytes, rounded to 256) + partial sums for the largest grid evc_gn_stats can pick
  const long long chunks = (long long)evc_num_sms() * 4 + 1;
  const long long per = ((long long)B * 4 + 255) / 256 * 256;
  long long c = chunks < HW ? chunks : HW;
  *bytes = per + (long long)B * c * C * 2 * (long long)sizeof(float);
  return EVC_OK;
}

static int gn_stats_impl(const void* x, const void* x_lo, int64_t ldx, int32_t B, int32_t HW, int32_t C, int64_t* stats,
                         int32_t c_total, int32_t c_off, void* workspace, int64_t workspace_bytes, evc_stream_t stream);
extern "C" int evc_gn_stats(const void* x, int64_t ldx, int32_t B, int32_t HW, int32_t C, int64_t* stats,
                            int32_t c_total, int32_t c_off, void* workspace, int64_t workspace_bytes,
                            evc_stream_t stream) {
  return gn_stats_impl(x, nullptr, ldx, B, HW, C, stats, c_total, c_off, workspace, workspace_bytes, stream);
}
extern "C" int evc_gn_stats_split(const void* x, const void* x_lo, int64_t ldx, int32_t B, int32_t HW, int32_t C,
                                  int64_t* stats, int32_t c_total, int32_t c_off, void* workspace,
                                  int64_t workspace_bytes, evc_stream_t stream) {
  if (!x_lo) return evc_set_error(EVC_ERR_INVALID, "evc_gn_stats_split: missing residual plane");
  return gn_stats_impl(x, x_lo, ldx, B, HW, C, stats, c_total, c_off, workspace, workspace_bytes, stream);
}
static int gn_stats_impl(const void* x, const void* x_lo, int64_t ldx, int32_t B, int32_t HW, int32_t C, int64_t* stats,
                         int32_t c_total, int32_t c_off, void* workspace, int64_t workspace_bytes, evc_stream_t stream) {
  if (!x || !stats || !workspace || B < 1 || HW < 1 || C < 8 || (C % 8) || (ldx % 8) || (c_off % 8))
    return evc_set_error(EVC_ERR_INVALID, "evc_gn_stats: bad arguments");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return evc_set_error(EVC_ERR_INVALID, "evc_gn_stats: x / workspace not 16B aligned");
  const int nvec = C / 8;
  if (nvec > 1024) return evc_set_error(EVC_ERR_INVALID, "evc_gn_stats: C too large");
  int rows = 256 / nvec;
  if (rows < 1) rows = 1;
  if (rows > HW) rows = HW;
  // aim for ~4 blocks per SM overall, at least `rows` pixels per block
  const int sms = evc_num_sms();
  int chunks = (sms * 4 + B - 1) / B;
  int max_chunks = (HW + rows - 1) / rows;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int ppb = (HW + chunks - 1) / chunks;
  chunks = (HW + ppb - 1) / ppb;
  const long long tick_bytes = ((long long)B * 4 + 255) / 256 * 256;
  const long long need = tick_bytes + (long long)B * chunks * C * 2 * (long long)sizeof(float);
  if (workspace_bytes < need) return evc_set_error(EVC_ERR_INVALID, "evc_gn_stats: workspace too small");
  unsigned* tickets = reinterpret_cast<unsigned*>(workspace);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + tick_bytes);
  dim3 block(nvec, rows), grid(chunks, B);
  const size_t smem = (size_t)rows * nvec * 16 * sizeof(float);
  cudaError_t le = evc_launch(gn_stats_kernel, grid, block, smem, (cudaStream_t)stream, 1,
                              reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx, (int)HW, (int)C,
                              reinterpret_cast<long long*>(stats), (int)c_total, (int)c_off, ppb, partials, tickets,
                              (long long)(x_lo ? reinterpret_cast<const __nv_bfloat16*>(x_lo) - reinterpret_cast<const __nv_bfloat16*>(x) : 0));
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch("gn_stats_kernel");
}

static int gn_apply_impl(const void* x0, const void* x0_lo, int32_t C0, const void* x1, const void* x1_lo, int32_t C1,
                         int32_t B, int32_t HW, const int64_t* stats0, const int64_t* stats1, int32_t groups, float eps,
                         const float* ss, int32_t adagn, int32_t silu, void* y, void* y_lo, evc_stream_t stream);

extern "C" int evc_gn_apply(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int32_t HW,
                            const int64_t* stats0, const int64_t* stats1, int32_t groups, float eps, const float* ss,
                            int32_t adagn, int32_t silu, void* y, evc_stream_t stream) {
  return gn_apply_impl(x0, nullptr, C0, x1, nullptr, C1, B, HW, stats0, stats1, groups, eps, ss, adagn, silu, y, nullptr,
                       stream);
}

extern "C" int evc_gn_apply_split(const void* x0, const void* x0_lo, int32_t C0, const void* x1, const void* x1_lo,
                                  int32_t C1, int32_t B, int32_t HW, const int64_t* stats0, const int64_t* stats1,
                                  int32_t groups, float eps, const float* ss, int32_t adagn, int32_t silu, void* y,
                                  void* y_lo, evc_stream_t stream) {
  if (!x0_lo || !y_lo || (x1 != nullptr && x1_lo == nullptr))
    return evc_set_error(EVC_ERR_INVALID, "evc_gn_apply_split: missing residual planes");
  return gn_apply_impl(x0, x0_lo, C0, x1, x1_lo, C1, B, HW, stats0, stats1, groups, eps, ss, adagn, silu, y, y_lo, stream);
}

static int gn_apply_impl(const void* x0, const void* x0_lo, int32_t C0, const void* x1, const void* x1_lo, int32_t C1,
                         int32_t B, int32_t HW, const int64_t* stats0, const int64_t* stats1, int32_t groups, float eps,
                         const float* ss, int32_t adagn, int32_t silu, void* y, void* y_lo, evc_stream_t stream) {
  if (!x0 || !stats0 || (x1 != nullptr && stats1 == nullptr) || !ss || !y || B < 1 || HW < 1 || C0 < 8 || (C0 % 8) || (C1 % 8) || (x1 == nullptr && C1 != 0) ||
      groups < 1 || ((C0 + C1) % groups))
    return evc_set_error(EVC_ERR_INVALID, "evc_gn_apply: bad arguments");
  const int C = C0 + C1;
  const int nvec = C / 8;
  if (nvec > 256 || groups > 64) return evc_set_error(EVC_ERR_INVALID, "evc_gn_apply: C > 2048 or groups > 64");
  if ((reinterpret_cast<uintptr_t>(stats0) & 15) || (reinterpret_cast<uintptr_t>(stats1) & 15) ||
      (reinterpret_cast<uintptr_t>(ss) & 15))
    return evc_set_error(EVC_ERR_INVALID, "evc_gn_apply: stats / ss must be 16-byte aligned");
  int rows = 256 / nvec;
  if (rows < 1) rows = 1;
  const int sms = evc_num_sms();
  static int waves = -1;
  if (waves < 0) {
    const char* e = getenv("EVC_GN_WAVES");
    waves = e ? atoi(e) : 1;
  }
  int chunks = (sms * 6 + B - 1) / B;
  if (waves > 0) {
    // whole waves of resident blocks (3 per SM, __launch_bounds__): B * chunks just below waves * 3 * sms
    chunks = (sms * 3 * waves) / B;
  }
  const int max_chunks = (HW + rows - 1) / rows;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int ppb = (HW + chunks - 1) / chunks;
  if (waves > 0) {
    // the rounding of pixels-per-block may only lower the number of blocks
    while (ppb > 1 && (long long)((HW + ppb - 2) / (ppb - 1)) * B <= (long long)sms * 3 * waves) --ppb;
  }
  chunks = (HW + ppb - 1) / ppb;
  static int stream_env = -1;
  if (stream_env < 0) {
    const char* e = getenv("EVC_GN_STREAM");
    stream_env = e ? atoi(e) : 1;
  }
  // streaming form for large tensors (same-box sweep, profiles/r02_notes.md): 32 pixels per thread if that still gives
  // >= 16 blocks per SM, else 16, else the one-wave kernel below (better under ~140 MB).  EVC_GN_STREAM=0 disables it,
  // EVC_GN_STREAM=n > 1 forces 4n pixels per thread.
  int sn = 0;
  if (stream_env > 1) {
    sn = stream_env;
  } else if (stream_env == 1) {
    for (int n = 8; n >= 4 && sn == 0; n >>= 1)
      if ((long long)B * ((HW + rows * 4 * n - 1) / (rows * 4 * n)) >= 16LL * sms) sn = n;
  }
  if (sn > 0 && y_lo == nullptr) {
    const int sppb = rows * 4 * sn;
    dim3 sgrid((HW + sppb - 1) / sppb, B), sblock(nvec, rows);
    cudaError_t se = evc_launch(silu ? gn_apply_stream_kernel<true> : gn_apply_stream_kernel<false>, sgrid, sblock, 0,
                                (cudaStream_t)stream, 1, reinterpret_cast<const __nv_bfloat16*>(x0), (int)C0,
                                reinterpret_cast<const __nv_bfloat16*>(x1), (int)C1, (int)HW,
                                reinterpret_cast<const long long*>(stats0), reinterpret_cast<const long long*>(stats1),
                                (int)groups, eps, ss, (int)adagn, reinterpret_cast<__nv_bfloat16*>(y), sppb);
    if (se != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(se));
    return evc_check_launch("gn_apply_stream_kernel");
  }
  dim3 grid(chunks, B), block(nvec, rows);
  cudaError_t le = evc_launch(silu ? gn_apply_kernel<true> : gn_apply_kernel<false>, grid, block, 0, (cudaStream_t)stream, 1,
                              reinterpret_cast<const __nv_bfloat16*>(x0), (int)C0, reinterpret_cast<const __nv_bfloat16*>(x1),
                              (int)C1, (int)HW, reinterpret_cast<const long long*>(stats0),
                              reinterpret_cast<const long long*>(stats1), (int)groups, eps, ss, (int)adagn,
                              reinterpret_cast<__nv_bfloat16*>(y), ppb, reinterpret_cast<const __nv_bfloat16*>(x0_lo),
                              reinterpret_cast<const __nv_bfloat16*>(x1_lo), reinterpret_cast<__nv_bfloat16*>(y_lo));
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch("gn_apply_kernel");
}

extern "C" int evc_gn_fir(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int32_t H, int32_t W,
                          const int64_t* stats0, const int64_t* stats1, int32_t groups, float eps, const float* ss,
                          int32_t adagn, int32_t up, void* y_act, void* y_raw0, void* y_raw1, evc_stream_t stream) {
  if (!x0 || !stats0 || !ss || !y_act || !y_raw0 || B < 1 || H < 1 || W < 1 || C0 < 8 || (C0 % 8) || (C1 % 8) ||
      (x1 == nullptr) != (C1 == 0) || (x1 != nullptr && (!stats1 || !y_raw1)) || groups < 1 || groups > 64 ||
      ((C0 + C1) % groups) || (!up && ((H | W) & 1)))
    return evc_set_error(EVC_ERR_INVALID, "evc_gn_fir: bad arguments");
  const int C = C0 + C1, nvec = C / 8;
  if (nvec > 256) return evc_set_error(EVC_ERR_INVALID, "evc_gn_fir: C > 2048");
  const void* ptrs[7] = {x0, x1, stats0, stats1, ss, y_act, y_raw0};
  for (int i = 0; i < 7; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return evc_set_error(EVC_ERR_INVALID, "evc_gn_fir: pointers must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(y_raw1) & 15) return evc_set_error(EVC_ERR_INVALID, "evc_gn_fir: pointers must be 16-byte aligned");
  GnFirArgs g;
  g.x0 = reinterpret_cast<const __nv_bfloat16*>(x0);
  g.x1 = reinterpret_cast<const __nv_bfloat16*>(x1);
  g.C0 = C0; g.C1 = C1; g.H = H; g.W = W;
  g.stats0 = reinterpret_cast<const long long*>(stats0);
  g.stats1 = reinterpret_cast<const long long*>(stats1);
  g.groups = groups; g.eps = eps; g.ss = ss; g.adagn = adagn;
  g.y_act = reinterpret_cast<__nv_bfloat16*>(y_act);
  g.y_raw0 = reinterpret_cast<__nv_bfloat16*>(y_raw0);
  g.y_raw1 = reinterpret_cast<__nv_bfloat16*>(y_raw1);
  static int down_env = -1;
  if (down_env < 0) {
    const char* e = getenv("EVC_GN_FIR_DOWN");  // 0: the two-activations-per-input walk (A/B)
    down_env = e ? atoi(e) : 1;
  }
  if (!up && down_env) {
    // channel chunk per block: the largest multiple of 8 that divides C and is <= 96
    int CB = 8;
    for (int cb = 96; cb >= 8; cb -= 8)
      if (C % cb == 0) { CB = cb; break; }
    const int nvb = CB / 8, n_chunks = C / CB;
    const int OW = W / 2, OH = H / 2;
    int OXB = (416 / nvb - 2) / 2;
    if (OXB > OW) OXB = OW;
    if (OXB < 1) OXB = 1;
    const int NC = 2 * OXB + 2;
    const int xb = (OW + OXB - 1) / OXB;
    int strip = 16;
    while (strip > 4 && (long long)xb * n_chunks * ((OH + strip - 1) / strip) * B < 2ll * evc_num_sms()) strip >>= 1;
    if (strip > OH) strip = OH;
    g.strip = strip;
    const size_t smem = (size_t)2 * 2 * NC * CB * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t ae = cudaFuncSetAttribute(gn_fir_down_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (ae != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(ae));
      attr_set = true;
    }
    dim3 grid(xb * n_chunks, (OH + strip - 1) / strip, B), block(nvb, NC);
    cudaError_t le = evc_launch(gn_fir_down_kernel, grid, block, smem, (cudaStream_t)stream, 1, g, OXB, CB, n_chunks);
    if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
    return evc_check_launch("gn_fir_down_kernel");
  }
  // 4 channels per thread (twice the resident warps) wherever the channel count allows a <= 256-wide block row
  static int vec_env = -1;
  if (vec_env < 0) {
    const char* e = getenv("EVC_GN_FIR_VEC");
    vec_env = e ? atoi(e) : 4;
  }
  const int V = (vec_env == 4 && (C % 4) == 0 && (C0 % 4) == 0 && C / 4 <= 256) ? 4 : 8;
  const int nv = C / V;
  int pxb = 256 / nv;
  if (pxb < 1) pxb = 1;
  const int cols = up ? W : W / 2, rows = up ? H : H / 2;
  if (pxb > cols) pxb = cols;
  // strips: enough blocks for ~4 per SM, at least 4 rows each (the window warm-up costs 1 (up) / 2 (down) rows)
  const int xb = (cols + pxb - 1) / pxb;
  int strip = 16;
  while (strip > 4 && (long long)xb * ((rows + strip - 1) / strip) * B < 4ll * evc_num_sms()) strip >>= 1;
  if (strip > rows) strip = rows;
  g.strip = strip;
  dim3 grid(xb, (rows + strip - 1) / strip, B), block(nv, pxb);
  void (*kernel)(GnFirArgs) = up ? (V == 4 ? gn_fir_kernel<true, 4> : gn_fir_kernel<true, 8>)
                                 : (V == 4 ? gn_fir_kernel<false, 4> : gn_fir_kernel<false, 8>);
  cudaError_t le = evc_launch(kernel, grid, block, 0, (cudaStream_t)stream, 1, g);
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch("gn_fir_kernel");
}

static int fir_impl(const void* x, const void* x_lo, void* y, void* y_lo, int32_t B, int32_t H, int32_t W, int32_t C,
                    int32_t up, evc_stream_t stream) {
  if (!x || !y || B < 1 || H < 1 || W < 1 || C < 8 || (C % 8) || (!up && ((H | W) & 1)) || ((x_lo != nullptr) != (y_lo != nullptr)))
    return evc_set_error(EVC_ERR_INVALID, "evc_fir_resample: bad arguments");
  const long long work = up ? (long long)B * H * W : (long long)B * (H / 2) * (W / 2);  // threads: input px (up) / output px
  const long long items = work * (C / 8);
  const int grid = grid_for(items, 256);
  const bool split = (x_lo != nullptr);
  const long long xo = split ? reinterpret_cast<const __nv_bfloat16*>(x_lo) - reinterpret_cast<const __nv_bfloat16*>(x) : 0;
  const long long yo = split ? reinterpret_cast<__nv_bfloat16*>(y_lo) - reinterpret_cast<__nv_bfloat16*>(y) : 0;
  auto kern = up ? (split ? fir_up_kernel<true> : fir_up_kernel<false>) : (split ? fir_down_kernel<true> : fir_down_kernel<false>);
  cudaError_t le = evc_launch(kern, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 1,
                              reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), (int)B, (int)H,
                              (int)W, (int)C, xo, yo);
  if (le != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(le));
  return evc_check_launch("fir_resample");
}

extern "C" int evc_fir_resample(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, int32_t up,
                                evc_stream_t stream) {
  return fir_impl(x, nullptr, y, nullptr, B, H, W, C, up, stream);
}

extern "C" int evc_fir_resample_split(const void* x, const void* x_lo, void* y, void* y_lo, int32_t B, int32_t H,
                                      int32_t W, int32_t C, int32_t up, evc_stream_t stream) {
  if (!x_lo || !y_lo) return evc_set_error(EVC_ERR_INVALID, "evc_fir_resample_split: missing residual planes");
  return fir_impl(x, x_lo, y, y_lo, B, H, W, C, up, stream);
}

extern "C" int evc_nearest_up2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C,
                               evc_stream_t stream) {
  if (!x || !y || B < 1 || H < 1 || W < 1 || C < 8 || (C % 8))
    return evc_set_error(EVC_ERR_INVALID, "evc_nearest_up2: bad arguments");
  const long long items = (long long)B * 4 * H * W * (C / 8);
  nearest_up2_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C);
  return evc_check_launch("nearest_up2_kernel");
}

extern "C" int evc_softmax_rows_split(const float* S, void* P, void* P_lo, int64_t rows, int32_t cols,
                                      evc_stream_t stream);
extern "C" int evc_softmax_rows(const float* S, void* P, int64_t rows, int32_t cols, evc_stream_t stream) {
  return evc_softmax_rows_split(S, P, nullptr, rows, cols, stream);
}
extern "C" int evc_softmax_rows_split(const float* S, void* P, void* P_lo, int64_t rows, int32_t cols,
                                      evc_stream_t stream) {
  if (!S || !P || rows < 1 || cols < 4 || (cols % 4)) return evc_set_error(EVC_ERR_INVALID, "evc_softmax_rows: bad arguments");
  const int warps = 8;
  const long long grid = (rows + warps - 1) / warps;
  if (grid > 0x7fffffffLL) return evc_set_error(EVC_ERR_INVALID, "evc_softmax_rows: too many rows");
  softmax_rows_kernel<<<(unsigned)grid, warps * 32, 0, (cudaStream_t)stream>>>(S, reinterpret_cast<__nv_bfloat16*>(P),
                                                                              rows, cols, reinterpret_cast<__nv_bfloat16*>(P_lo));
  return evc_check_launch("softmax_rows_kernel");
}

extern "C" int evc_pack_nchw_split(const void* src, int32_t src_is_f64, int32_t B, int32_t C, int32_t HW, float scale,
                                   float shift, void* dst, void* dst_lo, int32_t Cpad, int32_t c_off, evc_stream_t stream);
extern "C" int evc_pack_nchw(const void* src, int32_t src_is_f64, int32_t B, int32_t C, int32_t HW, float scale,
                             float shift, void* dst, int32_t Cpad, int32_t c_off, evc_stream_t stream) {
  return evc_pack_nchw_split(src, src_is_f64, B, C, HW, scale, shift, dst, nullptr, Cpad, c_off, stream);
}
extern "C" int evc_pack_nchw_split(const void* src, int32_t src_is_f64, int32_t B, int32_t C, int32_t HW, float scale,
                                   float shift, void* dst, void* dst_lo, int32_t Cpad, int32_t c_off, evc_stream_t stream) {
  if (!src || !dst || B < 1 || C < 1 || HW < 1 || c_off < 0 || c_off + C > Cpad)
    return evc_set_error(EVC_ERR_INVALID, "evc_pack_nchw: bad arguments");
  const int grid = grid_for((long long)B * HW, 256);
  if (src_is_f64)
    pack_nchw_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(src), B, C, HW,
                                                                     scale, shift, reinterpret_cast<__nv_bfloat16*>(dst),
                                                                     Cpad, c_off, reinterpret_cast<__nv_bfloat16*>(dst_lo));
  else
    pack_nchw_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(src), B, C, HW, scale,
                                                                    shift, reinterpret_cast<__nv_bfloat16*>(dst), Cpad,
                                                                    c_off, reinterpret_cast<__nv_bfloat16*>(dst_lo));
  return evc_check_launch("pack_nchw_kernel");
}

extern "C" int evc_fill_zero(void* p, int64_t bytes, evc_stream_t stream) {
  if (!p || bytes < 0) return evc_set_error(EVC_ERR_INVALID, "evc_fill_zero: bad arguments");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(e));
  return EVC_OK;
}

extern "C" int evc_inverse_transform(const float* x, float* frames, int64_t n, evc_stream_t stream) {
  if (!x || !frames || n < 1) return evc_set_error(EVC_ERR_INVALID, "evc_inverse_transform: bad arguments");
  inverse_transform_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, frames, n);
  return evc_check_launch("inverse_transform_kernel");
}

extern "C" int evc_frames_to_uint8(const float* frames, uint8_t* out, int64_t n, evc_stream_t stream) {
  if (!frames || !out || n < 1 || (reinterpret_cast<uintptr_t>(frames) & 15) || (reinterpret_cast<uintptr_t>(out) & 3))
    return evc_set_error(EVC_ERR_INVALID, "evc_frames_to_uint8: bad arguments (frames 16-byte, out 4-byte aligned)");
  frames_to_uint8_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(frames, out, n);
  return evc_check_launch("frames_to_uint8_kernel");
}
