// Fused self-attention forward on tcgen05:  O = softmax(scale * Q K^T) V  without the N x N matrix in HBM.
// Replaces, per head, the reference's einsum -> softmax -> einsum (models/better/layerspp.py:239-243,
// models/unet.py:114-119) and round 1's unfused QK^T GEMM + row-softmax + PV GEMM.
//
// One CTA per (128-query tile, head x column slice, sample), ONE pass over the keys (64 keys per tile) with an online
// softmax:
//   S = Q K^T -> row maximum of the tile; P = exp2((S - m) * scale*log2 e) (bf16), l += row sums, O += P V
//   end:  O /= l -> bf16
// The running reference m of a row is only raised when the tile's maximum exceeds it by more than 2^8 (lazy
// rescaling): P then stays below 2^8 -- harmless for bf16 (relative precision does not depend on magnitude) and for
// the fp32 accumulators -- and the O accumulator (TMEM) is multiplied by exp2((m_old - m_new) c) by the row's own
// thread, after the previous P V has completed and before the next one is issued.  With the scores of this model
// that happens in the first tiles of a row at most (round 1 recomputed S in a second pass instead: 1/3 more MMA work
// and twice the number of dependent tile hand-offs).
// V operand, two layouts: V^T (B, C, N) -- K-major B operand, what round 1's transposed-store projection wrote -- or V
// rows (B, N, ld) straight from a fused q|k|v projection: the 64-key x 64-channel boxes land as MN-major SWIZZLE_128B
// tiles (tcgen05 instruction descriptor bit 16), so the separate V projection launch and its scattered 2-byte stores
// disappear.
// Head dims above 384 (models/unet.py 'deeper': one head of 768): the O columns of a head are split over n_dsplit CTAs
// (each recomputes S) and Q no longer stays resident: every pipeline stage carries a 64-channel chunk of Q and of K.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = softmax (thread = query row).
// TMEM: S double buffer 2 x 64 columns at [0,128), O accumulator dv columns at [128, 128+dv), dv <= 384.
#include "evc_host.h"
#include "evc_ptx.cuh"

namespace evc {

constexpr int kAttnThreads = 256;
constexpr int kBQ = 128;  // queries per CTA
constexpr int kBK = 64;   // keys per tile
constexpr int kAttnMaxStages = 6;

struct alignas(64) AttnParams {
  CUtensorMap q_map;   // (2C | ld, N, B) box (64, 128, 1)
  CUtensorMap k_map;   // same tensor, box (64, 64, 1)
  CUtensorMap v_map;   // V^T (Np, C, B) box (64 keys, dv_box rows, 1)  |  V rows (ld, N, B) box (64 channels, 64 keys, 1)
  int N, C, heads, d;
  int dv;              // O columns per CTA: d / n_dsplit
  int n_dsplit;        // CTAs per (query tile, head); > 1 only for d > 384
  int dv_box;          // columns per PV MMA (dv if dv <= 256 else dv/2)
  int n_dchunks;       // dv / dv_box
  int v_mn;            // 1: V rows (MN-major B operand), 0: V^T (K-major)
  int q_stream;        // 1: Q is not resident, every QK stage holds [Q chunk 128x64 | K chunk 64x64]
  int num_stages;
  unsigned st_bytes;   // bytes per ring stage
  float scale_log2e;   // scale * log2(e)
  __nv_bfloat16* out;
  long long out_ld;    // elements between consecutive query rows
};

// MN-major operand tile, SWIZZLE_128B: rows of 64 consecutive MN elements (128 B) per K index, 8 K rows per 1024-byte
// atom (stride-byte-offset), the next 64 MN elements `lbo` bytes further (leading-byte-offset).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kAttnThreads, 1) evc_attn_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // warp-uniform warp index (see gemm_tc.cu): the producer / MMA warps run convergent loops, one elected lane issues
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBQ, head = blockIdx.y / p.n_dsplit, ds = blockIdx.y - head * p.n_dsplit, b = blockIdx.z;
  const int d = p.d, kc_n = d / 64, dv = p.dv;
  const int T = p.N / kBK;  // key tiles (N % 64 == 0; a 64-query sample uses half of the 128-row tile, TMA zero-fills the rest)
  // shared memory: Q [kc_n][128x64] (resident mode) | ring [stages][st_bytes] | P [2][128x64] | barriers
  const uint32_t q_bytes = p.q_stream ? 0u : kBQ * d * 2u;
  const uint32_t st_bytes = p.st_bytes;
  const uint32_t v_bytes = kBK * dv * 2u;
  const uint32_t sQ = base;
  const uint32_t sRing = sQ + q_bytes;
  const uint32_t sP = sRing + p.num_stages * st_bytes;
  const uint32_t bar = sP + 2u * (kBQ * kBK * 2u);
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kAttnMaxStages + s); };
  auto s_full = [&](int s) { return bar + 8u * (1 + 2 * kAttnMaxStages + s); };
  auto s_empty = [&](int s) { return bar + 8u * (3 + 2 * kAttnMaxStages + s); };
  auto p_full = [&](int s) { return bar + 8u * (5 + 2 * kAttnMaxStages + s); };
  auto p_empty = [&](int s) { return bar + 8u * (7 + 2 * kAttnMaxStages + s); };
  const uint32_t o_full = bar + 8u * (9 + 2 * kAttnMaxStages);
  const uint32_t tmem_slot = bar + 8u * (10 + 2 * kAttnMaxStages);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    tma_prefetch_desc(&p.v_map);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full(s), 1);
      mbar_init(s_empty(s), 4);
      mbar_init(p_full(s), 4);
      mbar_init(p_empty(s), 1);
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const uint32_t tmem_S = tmem_base;         // 2 x 64 columns
  const uint32_t tmem_O = tmem_base + 128u;  // dv columns

  pdl_wait();
  pdl_trigger();
  const int cq = head * d;          // channel offset of this head's q
  const int ck = p.C + head * d;    // ... and k inside the fused [q | k (| v)] rows
  const int cv = head * d + ds * dv;  // first V channel of this CTA (relative to the V tensor / V^T rows)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (!p.q_stream && elect_one()) {
      mbar_expect_tx(q_full, q_bytes);
      for (int kc = 0; kc < kc_n; ++kc) tma_load_3d(&p.q_map, sQ + kc * (kBQ * 128u), q_full, cq + kc * 64, q0, b);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
    };
    auto load_k = [&](int j) {
      if (p.q_stream) {
        for (int kc = 0; kc < kc_n; ++kc) {  // one stage per 64-channel chunk: [Q chunk | K chunk]
          mbar_wait(kv_empty(stage), phase ^ 1u);
          const uint32_t dst = sRing + stage * st_bytes;
          if (elect_one()) {
            mbar_expect_tx(kv_full(stage), (kBQ + kBK) * 128u);
            tma_load_3d(&p.q_map, dst, kv_full(stage), cq + kc * 64, q0, b);
            tma_load_3d(&p.k_map, dst + kBQ * 128u, kv_full(stage), ck + kc * 64, j * kBK, b);
          }
          __syncwarp();
          advance();
        }
        return;
      }
      mbar_wait(kv_empty(stage), phase ^ 1u);
      const uint32_t dst = sRing + stage * st_bytes;
      if (elect_one()) {
        mbar_expect_tx(kv_full(stage), kBK * d * 2u);
        for (int kc = 0; kc < kc_n; ++kc)
          tma_load_3d(&p.k_map, dst + kc * (kBK * 128u), kv_full(stage), ck + kc * 64, j * kBK, b);
      }
      __syncwarp();
      advance();
    };
    auto load_v = [&](int j) {
      mbar_wait(kv_empty(stage), phase ^ 1u);
      const uint32_t dst = sRing + stage * st_bytes;
      if (elect_one()) {
        mbar_expect_tx(kv_full(stage), v_bytes);
        if (p.v_mn) {  // 64-channel x 64-key boxes of the row-major V: [chunk][64 keys][128 B]
          for (int c = 0; c < dv / 64; ++c)
            tma_load_3d(&p.v_map, dst + c * (kBK * 128u), kv_full(stage), cv + c * 64, j * kBK, b);
        } else {
          for (int dc = 0; dc < p.n_dchunks; ++dc)
            tma_load_3d(&p.v_map, dst + dc * (p.dv_box * 128u), kv_full(stage), j * kBK, cv + dc * p.dv_box, b);
        }
      }
      __syncwarp();
      advance();
    };
    load_k(0);  // K_0, then (K_{j+1}, V_j)
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) load_k(j + 1);
      load_v(j);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane, always the same)
    const uint32_t idesc_s = umma_idesc_bf16(128u, kBK);
    // bit 16: B operand MN-major (V rows); otherwise both operands K-major
    const uint32_t idesc_o = umma_idesc_bf16(128u, static_cast<uint32_t>(p.dv_box)) | (p.v_mn ? (1u << 16) : 0u);
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
    };
    int t = 0;  // S tile counter: buffer t & 1, phase (t >> 1) & 1
    auto issue_s = [&]() {
      if (p.q_stream) {
        mbar_wait(s_empty(t & 1), ((t >> 1) & 1u) ^ 1u);
        for (int kc = 0; kc < kc_n; ++kc) {
          mbar_wait(kv_full(stage), phase);
          tc_fence_after();
          const uint32_t st = sRing + stage * st_bytes;
          if (elect_one()) {
            const uint64_t da = umma_desc_sw128(st);
            const uint64_t db = umma_desc_sw128(st + kBQ * 128u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_S + (t & 1) * kBK, da + 2u * k, db + 2u * k, idesc_s, (kc | k) != 0 ? 1u : 0u);
            umma_commit(kv_empty(stage));
            if (kc + 1 == kc_n) umma_commit(s_full(t & 1));
          }
          __syncwarp();
          advance();
        }
        ++t;
        return;
      }
      mbar_wait(kv_full(stage), phase);
      mbar_wait(s_empty(t & 1), ((t >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t kt = sRing + stage * st_bytes;
      if (elect_one()) {
        for (int kc = 0; kc < kc_n; ++kc) {
          const uint64_t da = umma_desc_sw128(sQ + kc * (kBQ * 128u));
          const uint64_t db = umma_desc_sw128(kt + kc * (kBK * 128u));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_S + (t & 1) * kBK, da + 2u * k, db + 2u * k, idesc_s, (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit(kv_empty(stage));
        umma_commit(s_full(t & 1));
      }
      __syncwarp();
      advance();
      ++t;
    };
    if (!p.q_stream) mbar_wait(q_full, 0);
    issue_s();                              // S_0
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_s();             // S_{j+1} runs while the softmax warps turn S_j into P_j
      mbar_wait(kv_full(stage), phase);     // V_j
      mbar_wait(p_full(j & 1), (j >> 1) & 1u);
      tc_fence_after();
      const uint32_t vt = sRing + stage * st_bytes;
      const uint64_t da = umma_desc_sw128(sP + (j & 1) * (kBQ * kBK * 2u));
      if (elect_one()) {
        for (int dc = 0; dc < p.n_dchunks; ++dc) {
          if (p.v_mn) {
            // [chunk of 64 channels][64 keys][128 B]: 64-channel blocks 8192 B apart, 16 keys = two 1024-byte atoms
            const uint64_t db = umma_desc_sw128_mn(vt + dc * (p.dv_box / 64) * (kBK * 128u), kBK * 128u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_O + dc * p.dv_box, da + 2u * k, db + 128u * k, idesc_o, (j | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t db = umma_desc_sw128(vt + dc * (p.dv_box * 128u));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_O + dc * p.dv_box, da + 2u * k, db + 2u * k, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(kv_empty(stage));
        umma_commit(p_empty(j & 1));
        if (j + 1 == T) umma_commit(o_full);
      }
      __syncwarp();
      advance();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / epilogue (thread = query row)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float mc = -INFINITY;  // running reference of this row, already multiplied by scale * log2(e)
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      mbar_wait(s_full(j & 1), (j >> 1) & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_S + lane_off + (j & 1) * kBK, v0);
      tmem_ld_32x32(tmem_S + lane_off + (j & 1) * kBK + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(j & 1));
      float mt = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mt = fmaxf(mt, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
      mt *= p.scale_log2e;  // scale > 0
      // lazy rescaling: keep the old reference unless this tile exceeds it by more than 8 (P would pass 2^8)
      const bool raise = mt > mc + 8.f;
      if (__any_sync(0xffffffffu, raise)) {  // warp-uniform: tcgen05.ld / st are warp collectives
        if (j > 0) {
          // the previous P V has completed (p_empty of its buffer) and the next one waits for this thread's p_full
          mbar_wait(p_empty((j - 1) & 1), ((j - 1) >> 1) & 1u);
          tc_fence_after();
          const float alpha = raise ? exp2f(mc - mt) : 1.f;  // 0 for the first raise from -inf cannot occur: j > 0
          l *= alpha;
          for (int c0 = 0; c0 < dv; c0 += 32) {
            uint32_t o[32];
            tmem_ld_32x32(tmem_O + lane_off + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tmem_O + lane_off + c0, o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        if (raise) mc = mt;
      }
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v0[i]), p.scale_log2e, -mc)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v0[i + 1]), p.scale_log2e, -mc)));
        l += e0 + e1;
        pk[i >> 1] = pack_bf16x2(e0, e1);
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v1[i]), p.scale_log2e, -mc)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v1[i + 1]), p.scale_log2e, -mc)));
        l += e0 + e1;
        pk[16 + (i >> 1)] = pack_bf16x2(e0, e1);
      }
      mbar_wait(p_empty(j & 1), ((j >> 1) & 1u) ^ 1u);  // the MMA that read this P buffer two tiles ago is done
      // K-major [128 x 64] bf16 tile, SWIZZLE_128B: 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
      const uint32_t prow = sP + (j & 1) * (kBQ * kBK * 2u) + row * 128u;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t a = prow + ((static_cast<uint32_t>(c) ^ (static_cast<uint32_t>(row) & 7u)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                     "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                     : "memory");
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(j & 1));
    }
    // epilogue: O / l -> bf16 rows
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    __nv_bfloat16* orow = p.out + ((long long)b * p.N + q0 + row) * p.out_ld + cv;
    const bool row_valid = (q0 + row) < p.N;
    for (int c0 = 0; c0 < dv; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_off + c0, v);
      tmem_ld_wait();
      if (!row_valid) continue;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[i + 0]) * inv, __uint_as_float(v[i + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c0 + i) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace evc

using namespace evc;

struct evc_attn_plan {
  AttnParams p;
  dim3 grid;
  int smem_bytes;
  double flops;
};

static int encode3(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                   uint32_t b0, uint32_t b1) {
  PFN_encodeTiled enc = evc_get_encode_tiled();
  if (enc == nullptr) return evc_set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t strides[2] = {s1, s2};
  uint32_t box[3] = {b0, b1, 1};
  uint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return evc_set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (attention)");
  return EVC_OK;
}

extern "C" int evc_attn_plan_create(const evc_attn_desc* a, evc_attn_plan** out) {
  if (!a || !out) return evc_set_error(EVC_ERR_INVALID, "null argument");
  *out = nullptr;
  const bool v_mn = (a->vT == nullptr);
  if (!a->qk || (v_mn && !a->v) || !a->out || a->B < 1 || a->heads < 1 || a->C < 64 || (a->C % a->heads))
    return evc_set_error(EVC_ERR_INVALID, "evc_attn_plan_create: bad arguments");
  const int d = a->C / a->heads;
  // O columns per CTA: the whole head up to 384 (one or two PV MMAs per key tile); above, the smallest split into
  // slices of <= 256 columns (multiples of 64) -- Q and K then stream through the ring in 64-channel chunks
  int n_dsplit = 1;
  if (d > 384) {
    while (n_dsplit <= 16 && ((d % n_dsplit) != 0 || ((d / n_dsplit) % 64) != 0 || d / n_dsplit > 256)) ++n_dsplit;
    if (n_dsplit > 16) n_dsplit = 0;
  }
  if ((a->N % kBK) != 0 || (d % 64) != 0 || n_dsplit == 0 || (d > 256 && d <= 384 && (d % 128) != 0))
    return evc_set_error(EVC_ERR_UNSUPPORTED,
                         "fused attention needs N % 64 == 0, head dim % 64 == 0 (% 128 between 256 and 384)");
  const int64_t v_ld = v_mn ? a->v_ld : a->vT_ld;
  const void* v_ptr = v_mn ? a->v : a->vT;
  if ((a->qk_ld % 8) || (v_ld % 8) || (a->out_ld % 8) || (reinterpret_cast<uintptr_t>(a->qk) & 15) ||
      (reinterpret_cast<uintptr_t>(v_ptr) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15))
    return evc_set_error(EVC_ERR_INVALID, "evc_attn_plan_create: 16-byte alignment required");
  evc_attn_plan* pl = new evc_attn_plan();
  AttnParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.C = a->C; p.heads = a->heads; p.d = d;
  p.n_dsplit = n_dsplit;
  p.dv = d / n_dsplit;
  p.dv_box = p.dv <= 256 ? p.dv : p.dv / 2;
  p.n_dchunks = p.dv / p.dv_box;
  p.v_mn = v_mn ? 1 : 0;
  p.q_stream = d > 384 ? 1 : 0;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.out_ld = a->out_ld;
  int rc = encode3(&p.q_map, a->qk, 2 * (uint64_t)a->C, a->N, a->B, (uint64_t)a->qk_ld * 2, (uint64_t)a->qk_ld * 2 * a->N, 64, kBQ);
  if (rc == EVC_OK)
    rc = encode3(&p.k_map, a->qk, 2 * (uint64_t)a->C, a->N, a->B, (uint64_t)a->qk_ld * 2, (uint64_t)a->qk_ld * 2 * a->N, 64, kBK);
  if (rc == EVC_OK) {
    if (v_mn)
      rc = encode3(&p.v_map, a->v, a->C, a->N, a->B, (uint64_t)a->v_ld * 2, (uint64_t)a->v_ld * 2 * a->N, 64, kBK);
    else
      rc = encode3(&p.v_map, a->vT, a->N, a->C, a->B, (uint64_t)a->vT_ld * 2, (uint64_t)a->vT_ld * 2 * a->C, kBK, p.dv_box);
  }
  if (rc != EVC_OK) {
    delete pl;
    return rc;
  }
  const int q_bytes = p.q_stream ? 0 : kBQ * d * 2, p_bytes = 2 * kBQ * kBK * 2;
  int st_bytes = p.q_stream ? (kBQ + kBK) * 128 : kBK * d * 2;
  if (kBK * p.dv * 2 > st_bytes) st_bytes = kBK * p.dv * 2;
  int stages = (227 * 1024 - 1024 - 512 - q_bytes - p_bytes) / st_bytes;
  if (stages > kAttnMaxStages) stages = kAttnMaxStages;
  if (stages < 2) {
    delete pl;
    return evc_set_error(EVC_ERR_UNSUPPORTED, "fused attention: head dim too large for shared memory");
  }
  p.num_stages = stages;
  p.st_bytes = static_cast<unsigned>(st_bytes);
  pl->smem_bytes = 1024 + q_bytes + stages * st_bytes + p_bytes + 512;
  pl->grid = dim3((a->N + kBQ - 1) / kBQ, a->heads * n_dsplit, a->B);
  pl->flops = 4.0 * a->B * (double)a->N * a->N * a->C;  // QK^T + PV
  *out = pl;
  return EVC_OK;
}

extern "C" int evc_attn_plan_launch(const evc_attn_plan* pl, evc_stream_t stream) {
  if (!pl) return evc_set_error(EVC_ERR_INVALID, "null plan");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(evc_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaError_t e = evc_launch(evc_attn_kernel, pl->grid, dim3(kAttnThreads), pl->smem_bytes, (cudaStream_t)stream, 1, pl->p);
  if (e != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(e));
  return evc_check_launch("evc_attn_kernel");
}

extern "C" void evc_attn_plan_destroy(evc_attn_plan* pl) { delete pl; }
extern "C" double evc_attn_plan_flops(const evc_attn_plan* pl) { return pl ? pl->flops : 0.0; }
