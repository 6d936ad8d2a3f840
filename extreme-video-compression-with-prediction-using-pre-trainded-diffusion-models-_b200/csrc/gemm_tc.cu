// Implicit-GEMM convolution / batched GEMM on the sm_100a tensor cores.
//
// One persistent, warp-specialised kernel: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer,
// warp 2 = TMEM allocator, warps 4..11 = epilogue.
//   * A operand: NHWC bf16 activations fetched by 4-D TMA boxes (64 channels x TW x TH x TB pixels = one
//     128 x 64 K-major SWIZZLE_128B tile).  A 3x3 convolution is nine shifted boxes; TMA out-of-bounds
//     zero fill implements the zero padding, so no im2col buffer ever exists in HBM.
//   * B operand: packed bf16 weights (N, K_total) fetched by 3-D TMA boxes (64 x BN x 1).
//   * accumulators: fp32 in TMEM, two stages of 256 columns so the epilogue of tile i overlaps the main
//     loop of tile i+1.
//   * epilogue, bf16 row outputs of whole 128-row tiles (every layer of the models): TMEM -> registers -> +bias
//     +residual (fetched by TMA) -> bf16 -> 64 B-swizzled 32 x 32 blocks in shared memory -> TMA stores; the fused
//     GroupNorm statistics are column sums read back from those blocks; with gn_ss the whole tile is parked in a
//     shared-memory slot and normalised + activated one tile later, once its sample's statistics are complete.
//     Other outputs (fp32, transposed, ragged toy shapes, split precision): registers -> global per thread.
// Replaces cuDNN conv2d / cuBLAS einsum of the reference (models/better/layers.py:89-113, 521-544;
// models/better/layerspp.py:239-243; models/unet.py:49-63, 114-119) and, fused, the GroupNorm reduction and the
// AdaGN + SiLU pass after Conv_0 (models/better/layerspp.py:520-527, 611-613).
#include <stdlib.h>

#include "evc_host.h"
#include "evc_ptx.cuh"

namespace evc {

constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 64 * 2;  // one A stage: 128 rows x 64 bf16
#ifndef EVC_EPI_WARPS
#define EVC_EPI_WARPS 8
#endif
// epilogue warps: kEpiGroups per TMEM lane quarter, each taking every kEpiGroups-th 32-column chunk of the tile.  The
// epilogue is latency-bound (TMEM loads, shared-memory round trips, shuffles), so it wants warps, not registers
constexpr int kEpiThreads = 32 * EVC_EPI_WARPS;
constexpr int kEpiGroups = kEpiThreads / 128;
constexpr int kChunkStride = 32 * kEpiGroups;  // columns between two chunks of one warp
constexpr int kThreads = 128 + kEpiThreads;
static_assert(EVC_EPI_WARPS % 4 == 0 && EVC_EPI_WARPS >= 4 && EVC_EPI_WARPS <= 16, "whole groups of four epilogue warps");
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages

constexpr int kMaxVSeg = 9;

// Fused GroupNorm apply: number of tile waits that gave up (see gn_pass2).  Read by evc_gemm_fault_count().
__device__ unsigned int g_gn_wait_faults = 0;

#ifdef EVC_GEMM_PROF
// Wait-time probe (tools/gpu_gemm_waits.py; never compiled into the shipped library):
// [0] MMA warp total, [1] MMA waits on full (operands), [2] MMA waits on tempty (accumulator),
// [3] producer total, [4] producer waits on empty, [5] epilogue total, [6] epilogue waits on tfull, [7] CTAs
__device__ unsigned long long g_prof[16];  // [8] epilogue: prefetch + bar.sync, [9] tcgen05.ld + wait, [10] math + stores + stats
#define PROF_DECL(...) __VA_ARGS__
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(acc, v) acc += clock64() - v
#define PROF_OUT(i, acc) atomicAdd(&g_prof[i], (unsigned long long)(acc))
#else
#define PROF_DECL(...)
#define PROF_T0(v)
#define PROF_ADD(acc, v)
#define PROF_OUT(i, acc)
#endif

// Split-precision ("fp32-tolerance") mode: every bf16 tensor has a second bf16 plane holding the rounding residual
// (x = hi + lo carries 16 mantissa bits) and a product is evaluated as hi*hi + hi*lo + lo*hi in the fp32 accumulator.
// In the K loop this is simply three "virtual segments" per source segment: (A_hi, W_hi), (A_hi, W_lo), (A_lo, W_hi).
struct alignas(64) GemmParams {
  CUtensorMap a_map[6];  // [0,3) hi planes of the source segments, [3,6) lo planes
  CUtensorMap b_map[2];  // W hi, W lo
  CUtensorMap out_map;   // bf16 row output as a 2-D tensor (N, B*H*W), 32 x 32 boxes, 64 B swizzle (TMA-store epilogue)
  int tma_out;           // 0: per-thread global stores; 1 / 2: TMA stores through 1 / 2 staging buffers per warp
  CUtensorMap resid_map; // residual as a 2-D tensor (N, B*H*W), 64 x 128 boxes, 128 B swizzle
  int resid_tma;         // 1: the tile's residual rows are fetched by TMA (needs tma_out), 0: cp.async per thread
  int off_resid, off_stat, off_stage;  // byte offsets of the epilogue scratch areas behind the barrier block
  int off_coef;          // fused GroupNorm apply: [2][BN] (a, b) coefficient pairs written by the coordinator warp
  // fused GroupNorm apply: the epilogue keeps the accumulator in TMEM until every tile of the sample has contributed
  // its statistics (per-sample ticket), then writes SiLU(GN(acc + bias) * gamma' + beta') instead of the raw value
  int gn_fuse;
  const float* gn_ss;   // [gamma' (N) | beta' (N)] of the current label
  int* gn_ticket;       // [B], zeroed by the caller before the launch
  float gn_eps, gn_inv_n;
  int gn_cpg, gn_adagn, tiles_per_sample;
#ifdef EVC_GEMM_PROF
  int exp_alt;           // timing experiment (wrong results): alternate the accumulator between consecutive MMAs
  int exp_skip;          // timing experiment (wrong results): skip operand loads, see the producer loop
#endif
  int n_seg;             // number of virtual segments
  int seg_a[kMaxVSeg];   // a_map index
  int seg_b[kMaxVSeg];   // b_map index
  int seg_koff[kMaxVSeg];  // first W column of the source segment
  int seg_taps[kMaxVSeg];
  int seg_kb[kMaxVSeg];  // 64-channel blocks per tap (last one may be partial: TMA zero-fills the A columns)
  int seg_c[kMaxVSeg];   // channels per tap
  int B, H, W;
  int TW, TH, TB;
  int tiles_x, tiles_y, tiles_b, tiles_n;
  unsigned mul_tiles_x, mul_tiles_y, mul_tiles_n, mul_split_k;  // fast_div multipliers
  int BN, N;
  int b_batched;
  int num_stages;
  int rows_valid;
  int total_kb;
  int out_mode;
  long long out_ld, out_bs;
  void* out;
  void* out_lo;  // split mode: residual plane of a bf16 output (NULL otherwise)
  const float* bias;
  const __nv_bfloat16* resid;
  const __nv_bfloat16* resid_lo;
  long long resid_ld;
  float alpha;
  unsigned tx_bytes;
  int resid_smem;  // 1: the epilogue stages residual rows in shared memory (cp.async prefetch)
  long long* stats;  // optional fused GroupNorm statistics: (B, N, 2) fixed-point [sum, sumsq] of the stored bf16 values
  int stats_combine;  // 1: the four epilogue warps of a tile belong to one sample (TW*TH == 128)
  int sample_rows;    // TW*TH
  int stride;         // convolution stride (1 or 2): input coordinate = stride * output coordinate + tap offset
  int m_tiles;        // tiles_x * tiles_y * tiles_b
  // split-K (few M tiles: small batches, 8x8 / 16x16 levels): a work unit is (tile, K slice); every unit writes its fp32
  // partial tile to `sk_ws`, the unit that arrives last on the tile's ticket adds the slices in the fixed order
  // 0..split_k-1 (deterministic whatever the arrival order) and runs the normal epilogue on the sum
  int split_k;
  float* sk_ws;            // [split_k][m_tiles_padded][sk_ld columns][128 rows] fp32
  int* sk_ticket;          // [total_tiles * CG], zero before the first launch; reset by the last arriver
  long long sk_plane;      // elements per K-slice plane
  int sk_ld;               // tiles_n * BN
  int sk_coop;             // 1: every unit is resident (grid == units): each K slice's CTA finishes the chunks it owns
};

// Work unit `tile` of a CTA (CG = 1) or CTA pair (CG = 2): N tile tn = tile % tiles_n, M tile(s) CG*(tile/tiles_n)+rank.
// n / d for the tile bookkeeping without the ~25-instruction integer division: m = ceil(2^32 / d) (0: d == 1 or the
// host could not prove n * d < 2^32, then the plain division is used).  Every warp decodes every tile.
__device__ __forceinline__ int fast_div(int n, int d, unsigned m) {
  return m != 0u ? static_cast<int>(__umulhi(static_cast<unsigned>(n), m)) : (d == 1 ? n : n / d);
}
template <int CG>
__device__ __forceinline__ void decode_tile(const GemmParams& p, int tile, int rank, int& x0, int& y0, int& b0,
                                            int& n0) {
  const int tq = fast_div(tile, p.tiles_n, p.mul_tiles_n);
  int tn = tile - tq * p.tiles_n;
  int tm = tq * CG + rank;  // may be == m_tiles for the peer of the last pair: fully out of range
  int t2 = fast_div(tm, p.tiles_x, p.mul_tiles_x);
  int tx = tm - t2 * p.tiles_x;
  int tb = fast_div(t2, p.tiles_y, p.mul_tiles_y);
  int ty = t2 - tb * p.tiles_y;
  x0 = tx * p.TW;
  y0 = ty * p.TH;
  b0 = tb * p.TB;
  n0 = tn * p.BN;
}

// Column sums over the 32 rows held by a warp: v[j] is column j of this lane's row; afterwards v[0] of lane l is
// the sum of column l over all 32 lanes.  31 shuffles (16+8+4+2+1) instead of 32 five-step reductions.
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}
constexpr float kStatScale = 1048576.f;  // 2^20 fixed point: integer atomics are order-independent => deterministic
__device__ __forceinline__ void stat_add(long long* dst, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(__float2ll_rn(v * kStatScale)));
}

__device__ __forceinline__ void store_chunk(const GemmParams& p, const float (&f)[32], int ncols, int n, long long pix,
                                            int b, long long pin, void* out_base) {
  // f[0..ncols) are final values for output columns n..n+ncols of pixel row `pix` (global row index),
  // `pin` = pixel index inside sample b.
  if (p.out_mode == EVC_OUT_BF16_ROWS) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_base) + pix * p.out_ld + n;
    if (n + ncols <= p.N && (p.N & 7) == 0 && (p.out_ld & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < ncols) {
          uint4 u;
          u.x = pack_bf16x2(f[j + 0], f[j + 1]);
          u.y = pack_bf16x2(f[j + 2], f[j + 3]);
          u.z = pack_bf16x2(f[j + 4], f[j + 5]);
          u.w = pack_bf16x2(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = u;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols && n + j < p.N) o[j] = __float2bfloat16_rn(f[j]);
    }
  } else if (p.out_mode == EVC_OUT_F32_ROWS) {
    float* o = reinterpret_cast<float*>(out_base) + pix * p.out_ld + n;
    if (n + ncols <= p.N && (p.N & 3) == 0 && (p.out_ld & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols && n + j < p.N) o[j] = f[j];
    }
  } else if (p.out_mode == EVC_OUT_BF16_T) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_base) + (long long)b * p.out_bs + pin;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols && n + j < p.N) o[(long long)(n + j) * p.out_ld] = __float2bfloat16_rn(f[j]);
  } else {
    float* o = reinterpret_cast<float*>(out_base) + (long long)b * p.out_bs + pin;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols && n + j < p.N) o[(long long)(n + j) * p.out_ld] = f[j];
  }
}

template <int CG, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) evc_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();  // the host sizes the allocation without alignment slack
  // warp index through a shuffle: the compiler then knows it is warp-uniform, so the producer / MMA loops below are
  // convergent code whose descriptors live in uniform registers (a `lane == 0` loop makes ptxas wrap every
  // UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;  // 0 = leader (issues the MMAs)
  const int unit = blockIdx.x / CG;
  const int num_units = gridDim.x / CG;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN / CG) * 128u;  // each CTA of a pair stages half of the B tile
  const uint32_t stage_bytes = kABytes + b_bytes;
  const uint32_t bar_base = base + p.num_stages * stage_bytes;
  // barrier slots (8 B each): full[8] | empty[8] | tmem_full[2] | tmem_empty[2] | tmem base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  const uint32_t resid_bar = bar_base + 8u * (2 * kMaxStages + 5);  // TMA-fetched residual tile landed
  // fused GroupNorm apply (slot s = tile parity): statistics of the tile issued | coefficients ready | coefficients read
  auto gn_stat_bar = [&](int s) { return bar_base + 200u + 8u * s; };
  auto gn_coef_bar = [&](int s) { return bar_base + 216u + 8u * s; };
  auto gn_free_bar = [&](int s) { return bar_base + 232u + 8u * s; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_seg; ++s) tma_prefetch_desc(&p.a_map[p.seg_a[s]]);
    tma_prefetch_desc(&p.b_map[0]);
    if (SPLIT) tma_prefetch_desc(&p.b_map[1]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar(s), CG);  // one arrival per producing CTA (the leader's barrier collects both)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG * (kEpiThreads / 32));  // one arrival per epilogue warp of every CTA
    }
    mbar_init(resid_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(gn_stat_bar(s), kEpiThreads / 32);
      mbar_init(gn_coef_bar(s), 1);
      mbar_init(gn_free_bar(s), kEpiThreads / 32);
    }
    fence_mbar_init();
  }
  if (CG == 2) cluster_sync_all();  // peer barriers initialised before anyone signals them; both CTAs alive
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2sm(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);  // warp-uniform for the compiler

  const int m_pairs = (p.m_tiles + CG - 1) / CG;
  const int total_tiles = m_pairs * p.tiles_n;
  const int total_units = total_tiles * p.split_k;  // split_k == 1: a unit is a tile
  pdl_wait();     // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp walks the loop, one
    // elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    PROF_DECL(long long w_empty = 0;)
    PROF_T0(t_prod);
    for (int u = unit; u < total_units; u += num_units) {
      const int tile = fast_div(u, p.split_k, p.mul_split_k);
      const int ks = u - tile * p.split_k;
      const int kb_lo = (ks * p.total_kb) / p.split_k, kb_hi = ((ks + 1) * p.total_kb) / p.split_k;
      int x0, y0, b0, n0;
      decode_tile<CG>(p, tile, rank, x0, y0, b0, n0);
      const int zb = p.b_batched ? b0 : 0;
      const int nb = n0 + rank * (p.BN / CG);  // this CTA's half of the B tile
      int kb = 0;
      for (int s = 0; s < p.n_seg; ++s) {
        const int taps = p.seg_taps[s];
        const CUtensorMap* am = &p.a_map[p.seg_a[s]];
        const CUtensorMap* bm = &p.b_map[p.seg_b[s]];
        for (int t = 0; t < taps; ++t) {
          const int ktap = p.seg_koff[s] + t * p.seg_c[s];
          const int dy = (taps == 9) ? (t / 3 - 1) : 0;
          const int dx = (taps == 9) ? (t % 3 - 1) : 0;
          for (int c = 0; c < p.seg_kb[s]; ++c, ++kb) {
            if (kb < kb_lo || kb >= kb_hi) continue;  // another unit's K slice
            PROF_T0(t_e);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            PROF_ADD(w_empty, t_e);
            const uint32_t sa = base + stage * stage_bytes;
#ifdef EVC_GEMM_PROF
            // timing experiments (wrong results): 1 = A only for the first tap of each row of taps, 4 = A only for tap 0,
            // 2 = B only for the first K block of a unit -- how much of the K-block time is operand delivery?
            const bool ld_a = !(((p.exp_skip & 1) && (t % 3) != 0) || ((p.exp_skip & 4) && t != 0));
            const bool ld_b = !((p.exp_skip & 2) && kb != kb_lo);
            const uint32_t txb = (ld_a ? static_cast<uint32_t>(p.rows_valid) * 128u : 0u) + (ld_b ? b_bytes : 0u);
#else
            const bool ld_a = true, ld_b = true;
            const uint32_t txb = p.tx_bytes;
#endif
            if (elect_one()) {
              if (CG == 2) {
                // both CTAs' bytes land on the leader's barrier; the peer contributes a plain (remote) arrival
                if (rank == 0) mbar_expect_tx(full_bar(stage), txb * 2u);
                else mbar_arrive_cluster(full_bar(stage), 0);
                if (ld_a) tma_load_4d_2sm(am, sa, full_bar(stage), c * 64, x0 * p.stride + dx, y0 * p.stride + dy, b0);
                if (ld_b) tma_load_3d_2sm(bm, sa + kABytes, full_bar(stage), ktap + c * 64, nb, zb);
              } else {
                mbar_expect_tx(full_bar(stage), txb);
                if (ld_a) tma_load_4d(am, sa, full_bar(stage), c * 64, x0 * p.stride + dx, y0 * p.stride + dy, b0);
                if (ld_b) tma_load_3d(bm, sa + kABytes, full_bar(stage), ktap + c * 64, nb, zb);
              }
            }
            __syncwarp();
            if (++stage == p.num_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    PROF_DECL(long long tot_prod = 0;)
    PROF_ADD(tot_prod, t_prod);
    if (lane == 0) {
      PROF_OUT(3, tot_prod);
      PROF_OUT(4, w_empty);
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only; whole warp walks
    // the loop with warp-uniform descriptors, one elected lane -- always the same one -- issues MMAs and commits)
    const uint32_t idesc = umma_idesc_bf16(128u * CG, static_cast<uint32_t>(p.BN));
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    PROF_DECL(long long w_full = 0; long long w_tempty = 0;)
    PROF_T0(t_mma);
    for (int u = unit; u < total_units; u += num_units, ++it) {
      const int ks = u % p.split_k;
      const int kb_lo = (ks * p.total_kb) / p.split_k, kb_hi = ((ks + 1) * p.total_kb) / p.split_k;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      PROF_T0(t_te);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      PROF_ADD(w_tempty, t_te);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kAccStride;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        PROF_T0(t_f);
        mbar_wait(full_bar(stage), phase);
        PROF_ADD(w_full, t_f);
        tc_fence_after();
        const uint32_t sa = base + stage * stage_bytes;
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sa + kABytes);
        if (elect_one()) {
#ifdef EVC_GEMM_PROF
          if (!(p.exp_skip & 8))  // timing experiment: no MMAs at all (commits only): pure operand-delivery rate
#endif
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
#ifdef EVC_GEMM_PROF
            const uint32_t td = tmem_d ^ ((p.exp_alt && (k & 1)) ? static_cast<uint32_t>(kAccStride) : 0u);
#else
            const uint32_t td = tmem_d;
#endif
            const uint32_t accum = (kb != kb_lo || k != 0) ? 1u : 0u;
            if (CG == 2) umma_bf16_2sm(td, da + 2u * k, db + 2u * k, idesc, accum);
            else umma_bf16(td, da + 2u * k, db + 2u * k, idesc, accum);
          }
          if (CG == 2) {
            umma_commit_2sm(empty_bar(stage), 3);  // frees this stage in both CTAs
            if (kb == kb_hi - 1) umma_commit_2sm(tfull_bar(acc), 3);
          } else {
            umma_commit(empty_bar(stage));
            if (kb == kb_hi - 1) umma_commit(tfull_bar(acc));
          }
        }
        __syncwarp();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    PROF_DECL(long long tot_mma = 0;)
    PROF_ADD(tot_mma, t_mma);
    if (lane == 0) {
      PROF_OUT(0, tot_mma);
      PROF_OUT(1, w_full);
      PROF_OUT(2, w_tempty);
      PROF_OUT(7, 1);
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ fused GroupNorm apply: ticket publisher.
    // Warps 2 and 3 keep everything that waits for global memory off the epilogue warps.  This one turns "the epilogue
    // has issued the statistics of tile k" (gn_stat_bar) into one ticket for the tile's sample, as early as possible: a
    // sample is complete for everybody only when its last ticket is out.
    if (!SPLIT && p.gn_fuse) {
      int it = 0;
      for (int u = unit; u < total_units; u += num_units, ++it) {
        int x0, y0, b0, n0;
        decode_tile<CG>(p, u, rank, x0, y0, b0, n0);
        if (b0 >= p.B) continue;  // the peer CTA of an odd last pair has no tile
        mbar_wait(gn_stat_bar(it & 1), static_cast<uint32_t>(it >> 1) & 1u);
        if (lane == 0) {
          // the statistics atomics of the epilogue threads happen before their arrivals on gn_stat_bar, which this
          // thread has observed: the fence makes them visible device-wide before the ticket is
          __threadfence();
          asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p.gn_ticket + b0), "r"(1) : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ fused GroupNorm apply: coefficients.
    // For each tile of this CTA: wait until every tile of its sample has published a ticket, fetch the channel sums of
    // the tile's groups once (one L2 round trip), turn them into per-column (a, b) with y = a * x + b, and hand them
    // to the epilogue's pass 2 (gn_coef_bar).  A sample's tiles are spread over two rounds of the persistent grid, so
    // the latency of this chain (statistics -> ticket -> poll -> sums -> coefficients) is paid once per round by the
    // CTAs that hold a sample's early tiles: it is kept as short as measured variants allow (profiles/r02_notes.md).
    if (!SPLIT && p.gn_fuse) {
      uint8_t* tail3 = smem_raw + (bar_base - smem_u32(smem_raw));
      float2* scoef_all = reinterpret_cast<float2*>(tail3 + p.off_coef);
      float2* sraw = scoef_all + 2 * p.BN;                          // raw sums of <= BN + 2 * 48 channels
      float* sgn = reinterpret_cast<float*>(sraw + p.BN + 96);      // [gamma' (N) | beta' (N)] of this launch's label
      for (int j = lane; j < 2 * p.N; j += 32) sgn[j] = __ldg(p.gn_ss + j);
      __syncwarp();
      int it = 0;
      for (int u = unit; u < total_units; u += num_units, ++it) {
        int x0, y0, b0, n0;
        decode_tile<CG>(p, u, rank, x0, y0, b0, n0);
        if (b0 >= p.B) continue;
        const int slot = it & 1;
        const uint32_t par = (it >> 1) & 1u;
        mbar_wait(gn_stat_bar(slot), par);  // nothing to poll for before this CTA's own tile has reported
        if (lane == 0) {
          int seen;
          const long long t0 = clock64();
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.gn_ticket + b0) : "memory");
            if (seen < p.tiles_per_sample) {
              __nanosleep(32);
              // The plan guarantees (host side, evc_gemm_plan_create) that no CTA waits for a third tile of a sample
              // it owns and that the whole grid fits on the device at once, so the other tiles are being computed and
              // this wait ends; foreign work holding SMs only delays it.  After ~10 s at any clock the caller broke
              // the contract (tickets not zeroed): count a fault and carry on with the statistics that are there --
              // wrong numbers for this sample, reported through evc_gemm_fault_count(), instead of a hung or trapped
              // context.
              if (clock64() - t0 > (1ll << 34)) {
                atomicAdd(&g_gn_wait_faults, 1u);
                break;
              }
            }
          } while (seen < p.tiles_per_sample);
        }
        __syncwarp();
        mbar_wait(gn_free_bar(slot), par ^ 1u);  // pass 2 of the tile two iterations back has read this buffer
        const int c_lo = (n0 / p.gn_cpg) * p.gn_cpg;
        const int c_end = min(n0 + p.BN, p.N);
        const int c_hi = min(p.N, ((c_end + p.gn_cpg - 1) / p.gn_cpg) * p.gn_cpg);
        for (int c = c_lo + lane; c < c_hi; c += 32) {
          const longlong2 st = __ldcg(reinterpret_cast<const longlong2*>(p.stats + ((long long)b0 * p.N + c) * 2));
          sraw[c - c_lo] = make_float2((float)((double)st.x * (1.0 / 1048576.0)), (float)((double)st.y * (1.0 / 1048576.0)));
        }
        __syncwarp();
        float2* scoef = scoef_all + slot * p.BN;
        for (int j = lane; j < p.BN; j += 32) {
          const int c = n0 + j;
          float a = 0.f, bb = 0.f;
          if (c < p.N) {
            const int g0 = (c / p.gn_cpg) * p.gn_cpg - c_lo;
            float sm = 0.f, qq = 0.f;
            for (int k = 0; k < p.gn_cpg; ++k) {
              const float2 t = sraw[g0 + k];
              sm += t.x;
              qq += t.y;
            }
            const float mean = sm * p.gn_inv_n;
            const float var = fmaxf(qq * p.gn_inv_n - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.gn_eps);
            a = rstd * (p.gn_adagn ? 1.f + sgn[c] : sgn[c]);
            bb = sgn[p.N + c] - mean * a;
          }
          scoef[j] = make_float2(a, bb);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gn_coef_bar(slot));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // Per tile: (1) while the main loop of this tile is still running, stage the bias slice in smem and
    // prefetch this thread's residual row with cp.async (each thread only ever reads the row it copied, so no
    // cross-thread synchronisation is needed for it); (2) wait for the accumulator; (3) TMEM -> registers ->
    // +bias +residual, *alpha -> global.
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int grp = (warp - 4) >> 2;   // which 32-column chunks this warp handles: grp, grp + kEpiGroups, ...
    const int row = q * 32 + lane;
    const int e = threadIdx.x - 128;  // 0..kEpiThreads-1
    const int dx = row % p.TW;
    const int dy = (row / p.TW) % p.TH;
    const int db = row / (p.TW * p.TH);
    uint8_t* gsm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* tail = gsm + (bar_base - base);
    float* sbias_all = reinterpret_cast<float*>(tail + 256);
    const uint32_t res_pitch = static_cast<uint32_t>(p.BN) * 2u + 16u;
    uint8_t* sres = tail + p.off_resid + static_cast<size_t>(row) * res_pitch;
    const uint32_t sres_u32 = bar_base + p.off_resid + static_cast<uint32_t>(row) * res_pitch;
    // TMA-fetched residual: [BN/64 panels][128 rows][128 B], 128 B swizzle
    const uint8_t* sres_t = tail + p.off_resid + row * 128;
    float* sstat = reinterpret_cast<float*>(tail + p.off_stat);  // [4 warps][BN][2] floats
    // TMA-store staging: [8 warps][tma_out buffers][32 rows x 64 B], 64 B swizzle
    uint8_t* stg_base = tail + p.off_stage + static_cast<uint32_t>(warp - 4) * static_cast<uint32_t>(p.tma_out) * 2048u;
    const uint32_t stg_base_u32 = bar_base + p.off_stage + static_cast<uint32_t>(warp - 4) * static_cast<uint32_t>(p.tma_out) * 2048u;
    int stg_i = 0;
    // fused GroupNorm apply: two tile-sized slots (bf16, TMA-store layout) replace the per-warp staging buffers
    uint8_t* stg_base0 = tail + p.off_stage;
    bool gn_pending = false;
    int gn_n0 = 0;
    long long gn_pix = 0;
    const float2* scoef_all = reinterpret_cast<const float2*>(tail + p.off_coef);
    auto gn_pass2 = [&](int pn0, long long ppix, int pit) {  // pit: iteration index of the parked tile
      const int slot_idx = pit & 1;
      mbar_wait(gn_coef_bar(slot_idx), static_cast<uint32_t>(pit >> 1) & 1u);  // coordinator: (a, b) of this tile's columns
      const float2* scoef = scoef_all + slot_idx * p.BN;
      uint8_t* slot = stg_base0 + static_cast<uint32_t>(slot_idx) * static_cast<uint32_t>(p.BN) * 256u;
      const uint32_t slot_u32 = bar_base + p.off_stage + static_cast<uint32_t>(slot_idx) * static_cast<uint32_t>(p.BN) * 256u;
      const int sw = (lane >> 1) & 3;
      for (int c0 = grp * 32; c0 < p.BN; c0 += kChunkStride) {
        const uint32_t boff = static_cast<uint32_t>((c0 >> 5) * 4 + q) * 2048u;
        uint8_t* srow = slot + boff + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4* ptr = reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4));
          const uint4 u = *ptr;
          float y[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                        bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 ab = scoef[c0 + 8 * j + k];
            y[k] = fmaf(y[k], ab.x, ab.y);
          }
#pragma unroll
          for (int k = 0; k < 8; k += 2) silu_pair(y[k], y[k + 1]);
          uint4 o;
          o.x = pack_bf16x2(y[0], y[1]);
          o.y = pack_bf16x2(y[2], y[3]);
          o.z = pack_bf16x2(y[4], y[5]);
          o.w = pack_bf16x2(y[6], y[7]);
          *ptr = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&p.out_map, slot_u32 + boff, pn0 + c0, static_cast<int>(ppix) + q * 32);
          bulk_commit();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(gn_free_bar(slot_idx));  // this warp is done with the coefficient buffer
    };
    // the bias vector, pre-multiplied by alpha (tiles_n * BN floats, zero beyond N), is staged once per CTA
    if (p.bias != nullptr)
      for (int j = e; j < p.tiles_n * p.BN; j += kEpiThreads) sbias_all[j] = (j < p.N) ? __ldg(p.bias + j) * p.alpha : 0.f;
    epi_bar();
    int it = 0;
    PROF_DECL(long long w_tfull = 0; long long w_pre = 0; long long w_ld = 0; long long w_rest = 0; long long w_bw = 0;
              long long w_st = 0; long long w_stat = 0;)
    PROF_T0(t_epi);
    int rit = 0;  // residual tiles fetched by TMA so far (phase of resid_bar)
    volatile int* sk_flag = reinterpret_cast<volatile int*>(tail + 192);  // free bytes of the barrier block
    for (int u = unit; u < total_units; u += num_units, ++it) {
      const int tile = fast_div(u, p.split_k, p.mul_split_k);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      int x0, y0, b0, n0;
      decode_tile<CG>(p, tile, rank, x0, y0, b0, n0);
      const int b = b0 + db;
      const bool valid = (row < p.rows_valid) && (b < p.B);
      const long long pin = (long long)(y0 + dy) * p.W + (x0 + dx);
      const long long pix = (long long)b * p.H * p.W + pin;
      const bool full_n = (n0 + p.BN <= p.N);
      const bool resid_fast = (p.resid != nullptr) && p.resid_smem && full_n;
      const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      bool from_ws = false;
      const float* ws_row = nullptr;
      int sk_ks = 0;
      int* sk_tk = nullptr;
#ifdef EVC_GEMM_PROF
      if (p.exp_skip & 16) {  // timing experiment: the epilogue only hands the accumulator back
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
          else mbar_arrive(tempty_bar(acc));
        }
        continue;
      }
#endif
      if (p.split_k > 1) {
        // ---- split-K, phase A (every unit): accumulator -> this K slice's fp32 plane, one ticket per (tile, CTA rank)
        const int ks = u - tile * p.split_k;
        const long long tm = (long long)(tile / p.tiles_n) * CG + rank;
        // plane layout [M tile][column][128 rows]: the 32 lanes of a warp (= 32 rows) write / read 128 contiguous bytes
        float* wrow = p.sk_ws + (long long)ks * p.sk_plane + (tm * p.sk_ld + n0) * 128 + row;
        ws_row = p.sk_ws + (tm * p.sk_ld + n0) * 128 + row;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        for (int c0 = grp * 32; c0 < p.BN; c0 += kChunkStride) {  // BN % 32 == 0 (checked by the host)
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(wrow + (c0 + j) * 128, __uint_as_float(v[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
          else mbar_arrive(tempty_bar(acc));
        }
        __threadfence();
        epi_bar();
        sk_ks = ks;
        sk_tk = p.sk_ticket + (long long)tile * CG + rank;
        if (p.sk_coop) {
          // every slice of the tile is being computed right now (grid == units, one CTA per SM): wait for all of them,
          // then this CTA finishes the 32-column chunks it owns (chunk index % split_k == ks) -- the S-fold reduction
          // and the epilogue run on S SMs instead of on the last arriver alone.  The ticket counts arrivals (0..S) and
          // then departures (S..2S); the last one to leave zeroes it for the next launch.
          if (e == 0) {
            atomicAdd(sk_tk, 1);
            int seen;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(sk_tk) : "memory");
            } while (seen < p.split_k);
          }
          epi_bar();
        } else {
          if (e == 0) {
            const int seen = atomicAdd(sk_tk, 1);
            const int last = (seen == p.split_k - 1) ? 1 : 0;
            if (last) *sk_tk = 0;  // every slice has arrived: ready for the next launch
            *sk_flag = last;
          }
          epi_bar();
          if (*sk_flag == 0) continue;  // another unit adds the slices up
          __threadfence();
        }
        from_ws = true;
      }
      const bool sk_owned_only = from_ws && p.sk_coop;
      // (1) prefetch
      PROF_T0(t_pre);
      epi_bar();  // everyone is done with the previous tile's residual rows and statistics scratch
      const float* sbias = sbias_all + n0;
      if (p.resid_tma) {
        if (warp == 4 && elect_one()) {
          const int panels = (p.BN + 63) >> 6;
          const long long tile_pix = ((long long)b0 * p.H + y0) * p.W + x0;
          mbar_expect_tx(resid_bar, static_cast<uint32_t>(panels) * 16384u);
          for (int pn = 0; pn < panels; ++pn)
            tma_load_2d(&p.resid_map, bar_base + p.off_resid + pn * 16384, resid_bar, n0 + 64 * pn,
                        static_cast<int>(tile_pix));
        }
      } else if (resid_fast && valid) {
        const __nv_bfloat16* r = p.resid + pix * p.resid_ld + n0;
        for (int c0 = grp * 32; c0 < p.BN; c0 += kChunkStride)  // only the chunks this thread will consume
#pragma unroll
          for (int j = c0; j < c0 + 32; j += 8)
            if (j < p.BN)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sres_u32 + j * 2), "l"(r + j) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      PROF_ADD(w_pre, t_pre);
      // (2) accumulator ready
      PROF_T0(t_tf);
      if (!from_ws) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
      }
      PROF_ADD(w_tfull, t_tf);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (p.resid_tma) {
        mbar_wait(resid_bar, static_cast<uint32_t>(rit) & 1u);
        ++rit;
      }
      if (!SPLIT && p.gn_fuse) {
        // ---- fused GroupNorm apply.  Pass 1 (this tile): bf16(acc + bias) -> this tile's slot in shared memory
        // (TMA-store layout: [chunk][row quarter][32 rows x 64 B, 64 B swizzle]) + column sums -> statistics + one
        // ticket for the sample; the accumulator is released at once.  Pass 2 (the PREVIOUS tile, one tile-time
        // later, when its sample is almost certainly complete): normalise + SiLU in place, TMA stores.  Every warp
        // only ever touches its own blocks of a slot, so the slots need no cross-warp synchronisation.
        const bool tile_ok = (b0 < p.B);  // the peer CTA of an odd last pair has no tile
        const int sw = (lane >> 1) & 3;
        uint8_t* slot = stg_base0 + static_cast<uint32_t>(it & 1) * static_cast<uint32_t>(p.BN) * 256u;
        if (elect_one()) bulk_wait_read<0>();  // the stores that read this slot were issued a whole tile ago
        __syncwarp();
        for (int c0 = grp * 32; c0 < p.BN; c0 += kChunkStride) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          if (c0 + kChunkStride >= p.BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
              else mbar_arrive(tempty_bar(acc));
            }
          }
          uint8_t* blk = slot + static_cast<uint32_t>((c0 >> 5) * 4 + q) * 2048u;
          uint8_t* srow = blk + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0v = *reinterpret_cast<const float4*>(sbias + c0 + 8 * j);
            const float4 b1v = *reinterpret_cast<const float4*>(sbias + c0 + 8 * j + 4);
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) + b0v.x, __uint_as_float(v[8 * j + 1]) + b0v.y);
            u.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) + b0v.z, __uint_as_float(v[8 * j + 3]) + b0v.w);
            u.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) + b1v.x, __uint_as_float(v[8 * j + 5]) + b1v.y);
            u.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) + b1v.z, __uint_as_float(v[8 * j + 7]) + b1v.w);
            *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = u;
          }
          __syncwarp();
          const int cp = lane & 15, par = lane >> 4;
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint8_t* sb = blk + ((cp & 3) << 2);
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const int rr = 2 * r + par;
            const uint32_t u = *reinterpret_cast<const uint32_t*>(sb + rr * 64 + (((cp >> 2) ^ (r & 3)) << 4));
            const float a0 = bf16_lo(u), a1 = bf16_hi(u);
            s0 += a0;
            s1 += a1;
            q0 = fmaf(a0, a0, q0);
            q1 = fmaf(a1, a1, q1);
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
          q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
          if (lane < 16) *reinterpret_cast<float4*>(sstat + (q * p.BN + c0 + 2 * cp) * 2) = make_float4(s0, q0, s1, q1);
        }
        if (grp * 32 >= p.BN) {  // this warp had no chunk: it still owes the accumulator release
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        epi_bar();
        if (tile_ok) {
          for (int i = e; i < 2 * p.BN; i += kEpiThreads) {
            const int col = i >> 1;
            if (n0 + col < p.N) {
              const float t = sstat[i] + sstat[2 * p.BN + i] + sstat[4 * p.BN + i] + sstat[6 * p.BN + i];
              stat_add(p.stats + ((long long)b0 * p.N + n0 + col) * 2 + (i & 1), t);
            }
          }
        }
        __syncwarp();
        if (tile_ok && lane == 0) mbar_arrive(gn_stat_bar(it & 1));  // the coordinator publishes the sample ticket
        if (gn_pending) gn_pass2(gn_n0, gn_pix, it - 1);
        gn_pending = tile_ok;
        gn_n0 = n0;
        gn_pix = ((long long)b0 * p.H + y0) * p.W + x0;
        continue;
      }
      bool released = from_ws;  // split-K: the accumulator was handed back in phase A
      for (int c0 = grp * 32; c0 < p.BN; c0 += kChunkStride) {
        if (sk_owned_only && ((c0 >> 5) % p.split_k) != sk_ks) continue;  // another slice's CTA finishes this chunk
        const int ncols = min(32, p.BN - c0);
        uint32_t v[32];
        PROF_T0(t_ld);
        if (from_ws) {
          // split-K, phase B: slices added in the fixed order 0, 1, ..., split_k-1
          float sacc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sacc[j] = 0.f;
          for (int ks = 0; ks < p.split_k; ++ks) {
            const float* src = ws_row + (long long)ks * p.sk_plane + c0 * 128;
#pragma unroll
            for (int j = 0; j < 32; ++j) sacc[j] += __ldcg(src + j * 128);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(sacc[j]);
        } else {
          if (ncols == 32)
            tmem_ld_32x32(taddr + c0, v);
          else
            tmem_ld_32x16(taddr + c0, v);
          tmem_ld_wait();
        }
        PROF_ADD(w_ld, t_ld);
        PROF_T0(t_rest);
        if (!from_ws && c0 + kChunkStride >= p.BN) {
          released = true;
          // last TMEM read of this tile is complete: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        const int n = n0 + c0;
        const float al = p.alpha;  // the staged bias is pre-multiplied: out = acc * alpha + bias * alpha + resid * alpha
        if (!SPLIT && p.tma_out != 0) {
          // ---- staged path (bf16 rows, full 128-row tiles, BN % 32 == 0): acc * alpha + bias' (+ resid * alpha) ->
          // swizzled smem -> one TMA store per 32 x 32 chunk; rows / columns outside the tensor are clipped by the TMA
          // unit.  Eight columns at a time, straight from the TMEM registers.
          const uint32_t boff = (p.tma_out == 2 ? static_cast<uint32_t>(stg_i & 1) : 0u) * 2048u;
          ++stg_i;
          PROF_T0(t_bw);
          if (elect_one()) {  // the TMA unit must have finished reading this buffer (store issued two / one chunks ago)
            if (p.tma_out == 2) bulk_wait_read<1>();
            else bulk_wait_read<0>();
          }
          __syncwarp();
          PROF_ADD(w_bw, t_bw);
          PROF_T0(t_st);
          uint8_t* srow = stg_base + boff + lane * 64;
          const int sw = (lane >> 1) & 3;
          const bool has_bias = p.bias != nullptr;
          const int rmode = p.resid_tma ? 1 : (resid_fast ? (valid ? 2 : 0) : ((p.resid != nullptr && valid) ? 3 : 0));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float g[8];
            if (has_bias) {
              const float4 b0v = *reinterpret_cast<const float4*>(sbias + c0 + 8 * j);
              const float4 b1v = *reinterpret_cast<const float4*>(sbias + c0 + 8 * j + 4);
              g[0] = b0v.x; g[1] = b0v.y; g[2] = b0v.z; g[3] = b0v.w;
              g[4] = b1v.x; g[5] = b1v.y; g[6] = b1v.z; g[7] = b1v.w;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) g[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = fmaf(__uint_as_float(v[8 * j + k]), al, g[k]);
            if (rmode == 1 || rmode == 2) {
              const int col = c0 + 8 * j;
              const uint4 u = (rmode == 1)
                  ? *reinterpret_cast<const uint4*>(sres_t + (col >> 6) * 16384 + ((((col & 63) >> 3) ^ (row & 7)) << 4))
                  : *reinterpret_cast<const uint4*>(sres + col * 2);
              g[0] = fmaf(bf16_lo(u.x), al, g[0]);
              g[1] = fmaf(bf16_hi(u.x), al, g[1]);
              g[2] = fmaf(bf16_lo(u.y), al, g[2]);
              g[3] = fmaf(bf16_hi(u.y), al, g[3]);
              g[4] = fmaf(bf16_lo(u.z), al, g[4]);
              g[5] = fmaf(bf16_hi(u.z), al, g[5]);
              g[6] = fmaf(bf16_lo(u.w), al, g[6]);
              g[7] = fmaf(bf16_hi(u.w), al, g[7]);
            } else if (rmode == 3) {
              const __nv_bfloat16* r = p.resid + pix * p.resid_ld + n + 8 * j;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (n + 8 * j + k < p.N) g[k] = fmaf(__bfloat162float(r[k]), al, g[k]);
            }
            uint4 u;
            u.x = pack_bf16x2(g[0], g[1]);
            u.y = pack_bf16x2(g[2], g[3]);
            u.z = pack_bf16x2(g[4], g[5]);
            u.w = pack_bf16x2(g[6], g[7]);
            *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = u;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {  // always the same lane: bulk groups are per thread
            const long long tile_pix = ((long long)b0 * p.H + y0) * p.W + x0;  // tile rows are consecutive pixels
            tma_store_2d(&p.out_map, stg_base_u32 + boff, n, static_cast<int>(tile_pix) + q * 32);
            bulk_commit();
          }
          PROF_ADD(w_st, t_st);
          PROF_T0(t_stat);
          if (p.stats != nullptr) {
            // column sums of the values as stored, read back from the staging buffer: lane -> (row parity, column
            // pair); 16 conflict-free 4-byte loads per lane, then one exchange between the two parities
            const int cp = lane & 15, par = lane >> 4;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            const uint8_t* sb = stg_base + boff + ((cp & 3) << 2);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const int rr = 2 * r + par;
              const uint32_t u = *reinterpret_cast<const uint32_t*>(sb + rr * 64 + (((cp >> 2) ^ (r & 3)) << 4));
              const float a0 = bf16_lo(u), a1 = bf16_hi(u);
              s0 += a0;
              s1 += a1;
              q0 = fmaf(a0, a0, q0);
              q1 = fmaf(a1, a1, q1);
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
            q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
            const int col = c0 + 2 * cp;
            if (p.stats_combine) {
              if (lane < 16) *reinterpret_cast<float4*>(sstat + (q * p.BN + col) * 2) = make_float4(s0, q0, s1, q1);
            } else {
              const int bw = b0 + (q * 32) / p.sample_rows;  // all 32 rows of this warp belong to one sample
              if (bw < p.B && lane < 16) {
                long long* dst = p.stats + ((long long)bw * p.N + n0 + col) * 2;
                if (n0 + col < p.N) {
                  stat_add(dst, s0);
                  stat_add(dst + 1, q0);
                }
                if (n0 + col + 1 < p.N) {
                  stat_add(dst + 2, s1);
                  stat_add(dst + 3, q1);
                }
              }
            }
          }
          PROF_ADD(w_stat, t_stat);
          PROF_ADD(w_rest, t_rest);
          continue;
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = (j < ncols) ? __uint_as_float(v[j]) * al : 0.f;
        if (valid) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const float4 bv = *reinterpret_cast<const float4*>(sbias + c0 + j);
                f[j + 0] += bv.x;
                f[j + 1] += bv.y;
                f[j + 2] += bv.z;
                f[j + 3] += bv.w;
              }
            }
          }
          if (resid_fast) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (j < ncols) {
                const uint4 u = *reinterpret_cast<const uint4*>(sres + (c0 + j) * 2);
                f[j + 0] = fmaf(bf16_lo(u.x), al, f[j + 0]);
                f[j + 1] = fmaf(bf16_hi(u.x), al, f[j + 1]);
                f[j + 2] = fmaf(bf16_lo(u.y), al, f[j + 2]);
                f[j + 3] = fmaf(bf16_hi(u.y), al, f[j + 3]);
                f[j + 4] = fmaf(bf16_lo(u.z), al, f[j + 4]);
                f[j + 5] = fmaf(bf16_hi(u.z), al, f[j + 5]);
                f[j + 6] = fmaf(bf16_lo(u.w), al, f[j + 6]);
                f[j + 7] = fmaf(bf16_hi(u.w), al, f[j + 7]);
              }
            }
          } else if (p.resid != nullptr) {
            const __nv_bfloat16* r = p.resid + pix * p.resid_ld + n;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols && n + j < p.N) f[j] = fmaf(__bfloat162float(r[j]), al, f[j]);
            if (SPLIT && p.resid_lo != nullptr) {
              const __nv_bfloat16* r2 = p.resid_lo + pix * p.resid_ld + n;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols && n + j < p.N) f[j] = fmaf(__bfloat162float(r2[j]), al, f[j]);
            }
          }
          store_chunk(p, f, ncols, n, pix, b, pin, p.out);
          if (SPLIT && p.out_lo != nullptr) {
            float g[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = f[j] - __bfloat162float(__float2bfloat16_rn(f[j]));
            store_chunk(p, g, ncols, n, pix, b, pin, p.out_lo);
          }
        }
        if (p.stats != nullptr) {
          // statistics of the values as stored (bf16-rounded); rows outside the tensor contribute zero
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            // bf16 mode: the value as stored; split mode: hi + lo represents f to 2^-17, use f itself
            const float r = (valid && j < ncols && n + j < p.N)
                                ? ((SPLIT && p.out_lo != nullptr) ? f[j] : __bfloat162float(__float2bfloat16_rn(f[j])))
                                : 0.f;
            f[j] = r;
            sq[j] = r * r;
          }
          warp_transpose_sum(f, lane);
          warp_transpose_sum(sq, lane);
          if (p.stats_combine) {
            if (lane < ncols) {
              sstat[(q * p.BN + c0 + lane) * 2 + 0] = f[0];
              sstat[(q * p.BN + c0 + lane) * 2 + 1] = sq[0];
            }
          } else {
            const int bw = b0 + (q * 32) / p.sample_rows;  // all 32 rows of this warp belong to one sample
            if (bw < p.B && q * 32 < p.rows_valid && lane < ncols && n + lane < p.N) {
              long long* dst = p.stats + ((long long)bw * p.N + n + lane) * 2;
              stat_add(dst, f[0]);
              stat_add(dst + 1, sq[0]);
            }
          }
        }
        PROF_ADD(w_rest, t_rest);
      }
      if (!released) {  // this warp had no chunk in this tile (BN <= 32 * grp)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
          else mbar_arrive(tempty_bar(acc));
        }
      }
      if (p.stats != nullptr && p.stats_combine) {
        epi_bar();
        if (b0 < p.B) {
          for (int i = e; i < 2 * p.BN; i += kEpiThreads) {
            const int col = i >> 1;
            if (n0 + col < p.N && !(sk_owned_only && ((col >> 5) % p.split_k) != sk_ks)) {
              const float t = sstat[i] + sstat[2 * p.BN + i] + sstat[4 * p.BN + i] + sstat[6 * p.BN + i];
              stat_add(p.stats + ((long long)b0 * p.N + n0 + col) * 2 + (i & 1), t);
            }
          }
        }
      }
      if (sk_owned_only) {  // departure: the last CTA to leave the tile zeroes its ticket for the next launch
        epi_bar();
        if (e == 0 && atomicAdd(sk_tk, 1) == 2 * p.split_k - 1) *sk_tk = 0;
      }
    }
    if (!SPLIT && p.gn_fuse && gn_pending) gn_pass2(gn_n0, gn_pix, it - 1);  // `it` = tiles done
    __syncwarp();
    // the staging buffers must have been read before the CTA's shared memory goes away; the writes themselves complete
    // with the grid
    if (p.tma_out != 0 && elect_one()) bulk_wait_read<0>();
#ifdef EVC_GEMM_PROF
    if (e == 0 && rank == 0) {
      long long tot_epi = 0;
      PROF_ADD(tot_epi, t_epi);
      PROF_OUT(5, tot_epi);
      PROF_OUT(6, w_tfull);
      PROF_OUT(8, w_pre);
      PROF_OUT(9, w_ld);
      PROF_OUT(10, w_rest);
      PROF_OUT(11, w_bw);
      PROF_OUT(12, w_st);
      PROF_OUT(13, w_stat);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its B half
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace evc

// ===================================================================================== host side
using namespace evc;

struct evc_gemm_plan {
  GemmParams p;
  int cg;  // 1: one CTA per 128-row tile; 2: CTA pair (cta_group::2), 256-row tile, B tile split across the pair
  int grid;
  int smem_bytes;
  bool split;  // split-precision operands: the <CG, true> instantiation
  double flops;
};

static cudaError_t ensure_gemm_attrs() {
  static bool attr_set = false;
  if (attr_set) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  void (*kernels[4])(GemmParams) = {evc_gemm_kernel<1, false>, evc_gemm_kernel<2, false>, evc_gemm_kernel<1, true>,
                                    evc_gemm_kernel<2, true>};
  for (int i = 0; i < 4 && e == cudaSuccess; ++i)
    e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) attr_set = true;
  return e;
}

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int conv_stride = 1, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = evc_get_encode_tiled();
  if (enc == nullptr) return evc_set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  if (rank == 4) estr[1] = estr[2] = (uint32_t)conv_stride;  // W and H traversal stride of a strided convolution
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] ptr %p", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
             (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
             ptr);
    return evc_set_error(EVC_ERR_CUDA, buf);
  }
  return EVC_OK;
}

extern "C" int evc_gemm_plan_create(const evc_gemm_desc* d, evc_gemm_plan** out_plan) {
  if (d == nullptr || out_plan == nullptr) return evc_set_error(EVC_ERR_INVALID, "null argument");
  *out_plan = nullptr;
  if (d->n_seg < 1 || d->n_seg > 3) return evc_set_error(EVC_ERR_INVALID, "n_seg must be 1..3");
  if (d->bn < 16 || d->bn > 256 || (d->bn % 16) != 0) return evc_set_error(EVC_ERR_INVALID, "bn must be 16..256, %16");
  if (d->B < 1 || d->H < 1 || d->W < 1) return evc_set_error(EVC_ERR_INVALID, "bad output extent");
  if (d->out == nullptr || d->w == nullptr) return evc_set_error(EVC_ERR_INVALID, "null out / w");
  if (d->out_mode < 0 || d->out_mode > 3) return evc_set_error(EVC_ERR_INVALID, "bad out_mode");
  if (d->w_batches != 1 && d->w_batches != d->B) return evc_set_error(EVC_ERR_INVALID, "w_batches must be 1 or B");

  evc_gemm_plan* pl = new evc_gemm_plan();
  GemmParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.H = d->H;
  p.W = d->W;
  p.b_batched = (d->w_batches > 1) ? 1 : 0;
  // M tile = TW x TH x TB pixels (<= 128)
  int TW = d->W < 128 ? d->W : 128;
  int TH = d->H < (128 / TW) ? d->H : (128 / TW);
  if (TH < 1) TH = 1;
  int TB = p.b_batched ? 1 : 128 / (TW * TH);
  if (TB < 1) TB = 1;
  if ((d->W % TW) != 0 || (d->H % TH) != 0 || (128 % TW) != 0 || TW * TH * TB > 128) {
    delete pl;
    return evc_set_error(EVC_ERR_INVALID, "W/H must tile into 128-pixel boxes (powers of two)");
  }
  p.TW = TW;
  p.TH = TH;
  p.TB = TB;
  p.rows_valid = TW * TH * TB;
  p.tiles_x = d->W / TW;
  p.tiles_y = d->H / TH;
  p.tiles_b = (d->B + TB - 1) / TB;
  p.BN = d->bn;
  p.N = d->w_rows;
  p.tiles_n = (d->w_rows + d->bn - 1) / d->bn;
  p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  // CTA pairs pay off when there are enough 256-row tiles to fill the 74 SM pairs; they need a shared B operand and
  // an N tile that splits into two halves of whole 8-row swizzle atoms.
  int cg = 1;
  {
    const bool can_pair = !p.b_batched && (d->bn % 16) == 0;  // each CTA stages bn/2 rows = whole 8-row swizzle atoms
    if (d->cta_group == 2) {
      if (!can_pair) {
        delete pl;
        return evc_set_error(EVC_ERR_INVALID, "cta_group 2 needs shared weights and bn % 16 == 0");
      }
      cg = 2;
    } else if (d->cta_group == 0 && can_pair) {
      // same-box A/B on B200 under the 1 kW cap (profiles/r01_notes.md): pairs are ~1.5 % faster end to end --
      // fewer B-tile bytes per FLOP lowers power, which raises the sustained clock
      // r02 (gpurun_out r2j, same box, forced pairs vs the old "at least two rounds of pair tiles" rule): pairs also win
      // 10-30 % on launches with few tiles (8x8 / 16x16 levels, small batches) -- half the B-operand traffic per SM and
      // more pipeline stages in the same shared memory -- unless the M tiles cannot be paired up without waste
      if (p.m_tiles >= 8 && ((p.m_tiles & 1) == 0 || p.m_tiles >= 23)) cg = 2;
    }
  }
  const int split_k = d->split_k > 1 ? d->split_k : 1;
  if (split_k > 1 && d->cta_group == 0) cg = 1;  // few tiles: nothing to pair
  pl->cg = cg;

  const int cs = d->stride <= 1 ? 1 : d->stride;
  if (cs != 1 && cs != 2) {
    delete pl;
    return evc_set_error(EVC_ERR_INVALID, "stride must be 1 or 2");
  }
  p.stride = cs;
  int total_kb = 0;
  long long ktot = 0;
  // split precision: all-or-nothing (every A segment and W carry a residual plane)
  const bool split = (d->w_lo != nullptr);
  for (int s = 0; s < d->n_seg; ++s) {
    if ((d->a_lo[s].ptr != nullptr) != split) {
      delete pl;
      return evc_set_error(EVC_ERR_INVALID, "split precision needs a_lo for every segment and w_lo (or none of them)");
    }
  }
  if (!split && (d->out_lo != nullptr || d->resid_lo != nullptr)) {
    delete pl;
    return evc_set_error(EVC_ERR_INVALID, "out_lo / resid_lo without split-precision operands");
  }
  int nv = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    if (d->taps[s] != 1 && d->taps[s] != 9) {
      delete pl;
      return evc_set_error(EVC_ERR_INVALID, "taps must be 1 or 9");
    }
    for (int plane = 0; plane < (split ? 2 : 1); ++plane) {
      const evc_tensor4& a = plane == 0 ? d->a[s] : d->a_lo[s];
      if (a.ptr == nullptr || (a.C % 8) != 0 || a.C < 8 || a.W != d->W * cs || a.H != d->H * cs || a.B != d->B ||
          a.C != d->a[s].C) {
        delete pl;
        return evc_set_error(EVC_ERR_INVALID, "A segment: C % 8 != 0 or extent mismatch");
      }
      if ((reinterpret_cast<uintptr_t>(a.ptr) & 15) || (a.stride_w % 8) || (a.stride_h % 8) || (a.stride_b % 8)) {
        delete pl;
        return evc_set_error(EVC_ERR_INVALID, "A segment: pointer/strides must be 16-byte aligned");
      }
      uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
      uint64_t strides[3] = {(uint64_t)a.stride_w * 2, (uint64_t)a.stride_h * 2, (uint64_t)a.stride_b * 2};
      uint32_t box[4] = {64, (uint32_t)(TW * cs), (uint32_t)(TH * cs), (uint32_t)TB};
      int rc = encode_map(&p.a_map[plane * 3 + s], a.ptr, 4, dims, strides, box, cs);
      if (rc != EVC_OK) {
        delete pl;
        return rc;
      }
    }
    const int C = d->a[s].C;
    // (A_hi, W_hi) [, (A_hi, W_lo), (A_lo, W_hi)]
    const int combos[3][2] = {{0, 0}, {0, 1}, {1, 0}};
    for (int c = 0; c < (split ? 3 : 1); ++c) {
      p.seg_a[nv] = combos[c][0] * 3 + s;
      p.seg_b[nv] = combos[c][1];
      p.seg_koff[nv] = (int)ktot;
      p.seg_taps[nv] = d->taps[s];
      p.seg_kb[nv] = (C + 63) / 64;
      p.seg_c[nv] = C;
      total_kb += d->taps[s] * ((C + 63) / 64);
      ++nv;
    }
    ktot += (long long)d->taps[s] * C;
  }
  p.n_seg = nv;
  if (ktot != d->w_k) {
    delete pl;
    return evc_set_error(EVC_ERR_INVALID, "w_k does not match sum(taps*C) of the A segments");
  }
  p.total_kb = total_kb;
  {
    if ((reinterpret_cast<uintptr_t>(d->w) & 15) || (d->w_row_stride % 8) || (d->w_batches > 1 && (d->w_batch_stride % 8))) {
      delete pl;
      return evc_set_error(EVC_ERR_INVALID, "W: pointer/strides must be 16-byte aligned");
    }
    uint64_t dims[3] = {(uint64_t)d->w_k, (uint64_t)d->w_rows, (uint64_t)d->w_batches};
    uint64_t bstride = d->w_batches > 1 ? (uint64_t)d->w_batch_stride * 2 : (uint64_t)d->w_row_stride * 2 * d->w_rows;
    if (bstride % 16) bstride = ((bstride + 15) / 16) * 16;
    uint64_t strides[2] = {(uint64_t)d->w_row_stride * 2, bstride};
    uint32_t box[3] = {64, (uint32_t)(d->bn / cg), 1};
    int rc = encode_map(&p.b_map[0], d->w, 3, dims, strides, box);
    if (rc == EVC_OK && split) {
      if (reinterpret_cast<uintptr_t>(d->w_lo) & 15) rc = evc_set_error(EVC_ERR_INVALID, "w_lo must be 16-byte aligned");
      else rc = encode_map(&p.b_map[1], d->w_lo, 3, dims, strides, box);
    }
    if (rc != EVC_OK) {
      delete pl;
      return rc;
    }
  }
  p.out = d->out;
  p.out_lo = d->out_lo;
  pl->split = split;
  p.resid_lo = reinterpret_cast<const __nv_bfloat16*>(d->resid_lo);
  p.out_mode = d->out_mode;
  p.out_ld = d->out_ld;
  p.out_bs = d->out_bs;
  p.bias = d->bias;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(d->resid);
  p.resid_ld = d->resid_ld;
  p.alpha = d->alpha;
  const int a_box_bytes = 64 * 2 * p.rows_valid;
  p.tx_bytes = (unsigned)(a_box_bytes + (d->bn / cg) * 128);  // per CTA

  const int stage_bytes = kABytes + (d->bn / cg) * 128;
  // shared memory: [stages][barriers 256 B][bias 1 KB][residual rows 128 x (2*BN + 16) B, only with a residual]
  p.resid_smem = (!split && d->resid != nullptr && (d->resid_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(d->resid) & 15) == 0 &&
                  (d->w_rows % 8) == 0) ? 1 : 0;
  p.stats = reinterpret_cast<long long*>(d->stats);
  p.sample_rows = TW * TH;
  p.stats_combine = (TW * TH == 128) ? 1 : 0;
  if (d->stats != nullptr && (d->out_mode != EVC_OUT_BF16_ROWS || ((TW * TH) % 32) != 0)) {
    delete pl;
    return evc_set_error(EVC_ERR_INVALID, "fused GroupNorm statistics need bf16 row output and H*W % 32 == 0");
  }
  // TMA-store epilogue (see the kernel): bf16 rows, whole 128-row tiles, 32-column chunks
  static int tma_env = -1, rtma_env = -1;
  if (tma_env < 0) {
    const char* e = getenv("EVC_GEMM_TMA_STORE");
    tma_env = e ? atoi(e) : 2;
    e = getenv("EVC_GEMM_TMA_RESID");
    rtma_env = e ? atoi(e) : 1;
  }
  const long long m_total = (long long)d->B * d->H * d->W;
  const bool can_tma_out = tma_env > 0 && !split && d->out_mode == EVC_OUT_BF16_ROWS && p.rows_valid == 128 &&
                           (d->bn % 32) == 0 && (d->out_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
                           m_total < (1ll << 31) - 256;  // pixel coordinates of the TMA maps are 32-bit, incl. the tile of an odd last pair
  p.resid_tma = (can_tma_out && rtma_env > 0 && d->resid != nullptr && (d->resid_ld % 8) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->resid) & 15) == 0) ? 1 : 0;
  if (p.resid_tma) p.resid_smem = 0;
  // scratch behind the barrier block: [barriers 256 B][bias 1 KB][residual][statistics][store staging]
  auto layout = [&](int nbuf) {
    int off = 256 + ((p.tiles_n * d->bn * 4 + 255) & ~255);  // barriers, then the whole bias vector
    if (p.resid_tma) {
      off = (off + 1023) & ~1023;
      p.off_resid = off;
      off += ((d->bn + 63) / 64) * 16384;
    } else {
      p.off_resid = off;
      if (p.resid_smem) off += 128 * (d->bn * 2 + 16);
    }
    p.off_stat = off;
    if (d->stats) off += 32 * d->bn;
    p.off_coef = off;
    // two coefficient buffers, raw sums of the tile's groups (cpg <= 48 checked below), the AdaGN row of the launch
    if (d->gn_ss != nullptr) off += 2 * d->bn * 8 + (d->bn + 96) * 8 + 2 * d->w_rows * 4;
    if (nbuf > 0) {
      off = (off + 1023) & ~1023;
      p.off_stage = off;
      // fused GroupNorm apply: two whole-tile slots (bn x 256 B each) instead of the per-warp staging buffers
      off += (d->gn_ss != nullptr) ? 2 * d->bn * 256 : (kEpiThreads / 32) * nbuf * 2048;
    }
    return off;
  };
  p.tma_out = 0;
#ifdef EVC_GEMM_PROF
  p.exp_alt = getenv("EVC_EXP_ALT") ? atoi(getenv("EVC_EXP_ALT")) : 0;
  p.exp_skip = getenv("EVC_EXP_SKIP") ? atoi(getenv("EVC_EXP_SKIP")) : 0;
#endif
  if (can_tma_out) {
    uint64_t dims[2] = {(uint64_t)d->w_rows, (uint64_t)m_total};
    uint64_t strides[1] = {(uint64_t)d->out_ld * 2};
    uint32_t box[2] = {32, 32};
    int rc = encode_map(&p.out_map, d->out, 2, dims, strides, box, 1, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc == EVC_OK && p.resid_tma) {
      uint64_t rstrides[1] = {(uint64_t)d->resid_ld * 2};
      uint32_t rbox[2] = {64, 128};
      rc = encode_map(&p.resid_map, d->resid, 2, dims, rstrides, rbox);
    }
    if (rc != EVC_OK) {
      delete pl;
      return rc;
    }
    // two staging buffers per epilogue warp unless that would leave fewer than four pipeline stages
    p.tma_out = (tma_env >= 2 && (227 * 1024 - layout(2)) / stage_bytes >= 4) ? 2 : 1;
  }
  p.split_k = 1;
  if (split_k > 1) {
    const long long rows = (long long)((p.m_tiles + cg - 1) / cg) * cg * 128;
    const long long ld = (long long)p.tiles_n * d->bn;
    const long long need = (long long)split_k * rows * ld * 4;
    if ((d->bn % 32) != 0 || split || d->gn_ss != nullptr || d->sk_ws == nullptr || d->sk_ticket == nullptr ||
        split_k > total_kb || d->sk_ws_bytes < need || (reinterpret_cast<uintptr_t>(d->sk_ws) & 15)) {
      delete pl;
      return evc_set_error(EVC_ERR_INVALID,
                           "split_k needs bn % 32 == 0, no gn_ss / split-precision planes, split_k <= K blocks, and a "
                           "16-byte aligned sk_ws of split_k * roundup(m_tiles) * 128 * tiles_n * bn floats + sk_ticket");
    }
    p.split_k = split_k;
    p.sk_ws = reinterpret_cast<float*>(d->sk_ws);
    p.sk_ticket = d->sk_ticket;
    p.sk_plane = rows * ld;
    p.sk_ld = (int)ld;
  }
  p.gn_fuse = 0;
  if (d->gn_ss != nullptr) {
    const bool ok = p.tma_out != 0 && d->stats != nullptr && d->gn_ticket != nullptr && d->bias != nullptr &&
                    d->resid == nullptr && d->alpha == 1.0f && TW * TH == 128 && d->gn_groups > 0 &&
                    (d->w_rows % d->gn_groups) == 0 && d->w_rows / d->gn_groups <= 48 && d->max_ctas <= 0;
    if (!ok) {
      delete pl;
      return evc_set_error(EVC_ERR_INVALID,
                           "fused GroupNorm apply needs stats + ticket + bias, no residual, alpha 1, bf16 rows through the "
                           "TMA-store epilogue (H*W % 128 == 0, W <= 128, bn % 32 == 0) and N % groups == 0");
    }
    // (a) a CTA normalises tile k only after accumulating its tile k+1, and waits there for every tile of the sample:
    //     it must not own a third tile of the same sample (units are handed out round robin, `num_units` apart);
    // (b) all CTAs must be resident together (they wait for each other): one CTA per SM / one pair per SM pair.
    {
      const int cap = (d->max_ctas > 0 ? d->max_ctas : evc_num_sms()) / cg;
      const long long units_total = (long long)((p.m_tiles + cg - 1) / cg) * p.tiles_n;
      const long long num_units = units_total < cap ? units_total : cap;
      const long long sample_units = ((long long)p.tiles_x * p.tiles_y + cg - 1) / cg * p.tiles_n + (cg > 1 ? p.tiles_n : 0);
      int resident = 0;
      const int smem_need = 227 * 1024;
      cudaError_t oe = ensure_gemm_attrs();
      if (oe == cudaSuccess)
        oe = cg == 2
          ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, evc_gemm_kernel<2, false>, kThreads, smem_need)
          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, evc_gemm_kernel<1, false>, kThreads, smem_need);
      if (oe != cudaSuccess) {
        (void)cudaGetLastError();
        resident = 1;  // no device here (plan built for inspection only): the launch itself would fail first
      }
      if (sample_units > 2 * num_units || resident < 1 || num_units * cg > (long long)resident * evc_num_sms()) {
        delete pl;
        return evc_set_error(EVC_ERR_UNSUPPORTED,
                             "fused GroupNorm apply: a sample has more tiles than two rounds of the persistent grid (or the "
                             "grid cannot be co-resident); use the unfused gn_apply path for this shape");
      }
    }
    p.gn_fuse = 1;
    p.gn_ss = d->gn_ss;
    p.gn_ticket = d->gn_ticket;
    p.gn_eps = d->gn_eps;
    p.gn_adagn = d->gn_adagn;
    p.gn_cpg = d->w_rows / d->gn_groups;
    p.gn_inv_n = 1.0f / ((float)p.gn_cpg * (float)d->H * (float)d->W);
    p.tiles_per_sample = p.tiles_x * p.tiles_y * p.tiles_n;
  }
  const int tail_bytes = layout(p.tma_out);
  const int budget = 227 * 1024 - tail_bytes;  // the kernel's dynamic shared memory is declared 1024-byte aligned
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    delete pl;
    return evc_set_error(EVC_ERR_INVALID, "tile does not fit in shared memory");
  }
#ifdef EVC_GEMM_PROF
  if (getenv("EVC_EXP_STAGES") && atoi(getenv("EVC_EXP_STAGES")) >= 2 && atoi(getenv("EVC_EXP_STAGES")) < stages)
    stages = atoi(getenv("EVC_EXP_STAGES"));
#endif
  p.num_stages = stages;
  pl->smem_bytes = stages * stage_bytes + tail_bytes;

  const long long units = (long long)((p.m_tiles + cg - 1) / cg) * p.tiles_n * p.split_k;  // (tile | tile pair) x K slice
  int sms = evc_num_sms();
  int cap = (d->max_ctas > 0 ? d->max_ctas : sms) / cg;
  if (cap < 1) cap = 1;
  pl->grid = cg * (int)(units < cap ? units : cap);
  p.sk_coop = (p.split_k > 1 && units <= cap) ? 1 : 0;  // every K slice of every tile on its own resident CTA (pair)
  {
    // exact for n * d < 2^32: n <= 2 * units + 1 (unit, tile and M-tile indices), d <= 2^12 here
    auto mul = [&](int dd) -> unsigned {
      if (dd <= 1 || (unsigned long long)(2 * units + 4) * (unsigned long long)dd >= (1ull << 32)) return 0u;
      return (unsigned)(((1ull << 32) + (unsigned long long)dd - 1) / (unsigned long long)dd);
    };
    p.mul_tiles_x = mul(p.tiles_x);
    p.mul_tiles_y = mul(p.tiles_y);
    p.mul_tiles_n = mul(p.tiles_n);
    p.mul_split_k = mul(p.split_k);
  }
  pl->flops = 2.0 * (double)d->B * d->H * d->W * (double)d->w_rows * (double)d->w_k;
  *out_plan = pl;
  return EVC_OK;
}

extern "C" int evc_gemm_plan_launch(const evc_gemm_plan* pl, const float* bias_override, evc_stream_t stream) {
  return evc_gemm_plan_launch_ex(pl, bias_override, nullptr, nullptr, stream);
}

extern "C" int evc_gemm_plan_launch_gn(const evc_gemm_plan* pl, const float* bias_override, const float* gn_ss_override,
                                       evc_stream_t stream) {
  return evc_gemm_plan_launch_ex(pl, bias_override, gn_ss_override, nullptr, stream);
}

extern "C" int evc_gemm_plan_launch_ex(const evc_gemm_plan* pl, const float* bias_override, const float* gn_ss_override,
                                       void* out_override, evc_stream_t stream) {
  if (pl == nullptr) return evc_set_error(EVC_ERR_INVALID, "null plan");
  if (gn_ss_override != nullptr && !pl->p.gn_fuse)
    return evc_set_error(EVC_ERR_INVALID, "gn_ss_override on a plan without fused GroupNorm apply");
  if (out_override != nullptr && (pl->p.tma_out != 0 || pl->split))
    return evc_set_error(EVC_ERR_INVALID, "out_override needs a per-thread-store output (fp32 / transposed modes)");
  {
    cudaError_t ae = ensure_gemm_attrs();
    if (ae != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(ae));
  }
  GemmParams p = pl->p;
  if (bias_override != nullptr) p.bias = bias_override;
  if (gn_ss_override != nullptr) p.gn_ss = gn_ss_override;
  if (out_override != nullptr) p.out = out_override;
  cudaError_t e;
  const bool split = pl->split;
  void (*kernel)(GemmParams) = pl->cg == 2 ? (split ? evc_gemm_kernel<2, true> : evc_gemm_kernel<2, false>)
                                           : (split ? evc_gemm_kernel<1, true> : evc_gemm_kernel<1, false>);
  e = evc_launch(kernel, dim3(pl->grid), dim3(kThreads), pl->smem_bytes, (cudaStream_t)stream, pl->cg, p);
  if (e != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(e));
  return evc_check_launch(pl->cg == 2 ? "evc_gemm_kernel<2>" : "evc_gemm_kernel<1>");
}

// Synchronising debug query: how many fused-GroupNorm tile waits gave up since the last call (0 unless a caller broke
// the ticket contract).  Not for use inside stream capture.
extern "C" int64_t evc_gemm_fault_count(void) {
  unsigned int v = 0, z = 0;
  if (cudaMemcpyFromSymbol(&v, evc::g_gn_wait_faults, sizeof(v)) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  if (v != 0) (void)cudaMemcpyToSymbol(evc::g_gn_wait_faults, &z, sizeof(z));
  return (int64_t)v;
}

extern "C" int evc_gemm_plan_cta_group(const evc_gemm_plan* pl) { return pl ? pl->cg : 0; }

extern "C" void evc_gemm_plan_destroy(evc_gemm_plan* pl) { delete pl; }
extern "C" double evc_gemm_plan_flops(const evc_gemm_plan* pl) { return pl ? pl->flops : 0.0; }

#ifdef EVC_GEMM_PROF
// debug build only: copies the wait-time counters to the host and clears them
extern "C" int evc_gemm_prof_read(unsigned long long* host8) {
  cudaError_t e = cudaMemcpyFromSymbol(host8, evc::g_prof, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(e));
  unsigned long long z[16] = {0};
  e = cudaMemcpyToSymbol(evc::g_prof, z, sizeof(z));
  return e == cudaSuccess ? EVC_OK : evc_set_error(EVC_ERR_CUDA, cudaGetErrorString(e));
}
#endif
