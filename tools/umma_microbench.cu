// Isolated tcgen05.mma issue-rate microbenchmark ("settle the UMMA-floor question").
//
// Every CTA (or CTA pair) owns a few shared-memory operand stages (SWIZZLE_128B K-major layout, contents irrelevant),
// and one warp issues `kblocks` x 4 MMAs of shape (128*CG) x N x 16 (bf16 -> fp32 in TMEM), one commit per K block --
// exactly the inner loop of evc_gemm_kernel without TMA, barriers on operands or an epilogue.  Reported: SM cycles per
// MMA (clock64 of CTA 0 around the whole sequence, commit-to-completion included) and chip TFLOP/s from CUDA events.
//
//   STYLE 0  `if (threadIdx.x == 32)` single-thread loop: what round 1 shipped.  ptxas wraps every UTCHMMA / UTCBAR in an
//            ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop because the descriptors are computed in divergent code.
//   STYLE 1  whole warp walks the loop, descriptors are warp-uniform, one elected lane issues (what ships now).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/umma_microbench tools/umma_microbench.cu
// (tools/gpu_microbench.sh builds and runs it; never part of the product library).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200/csrc/evc_ptx.cuh"

using namespace evc;

constexpr int kStages = 4;
constexpr int kStageBytes = 16384 + 32768;  // A 128x64 bf16 + B up to 256x64 bf16

template <int CG, int STYLE>
__global__ void __launch_bounds__(128, 1) umma_bench(int N, int kblocks, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + kStages * kStageBytes;
  const uint32_t tmem_slot = bar + 64;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  // deterministic, finite operand contents (bf16 1.0 / 0.5 patterns)
  for (uint32_t i = threadIdx.x; i < kStages * kStageBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3f803f00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 2) tmem_alloc_2sm(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  const uint32_t idesc = umma_idesc_bf16(128u * CG, static_cast<uint32_t>(N));
  long long t0 = 0, t1 = 0;
  if (rank == 0) {
    if (STYLE == 0) {
      if (threadIdx.x == 32) {
        t0 = clock64();
        int stage = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          const uint32_t sa = base + stage * kStageBytes;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (CG == 2) umma_bf16_2sm(tmem_base, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(tmem_base, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (kb == kblocks - 1) {
            if (CG == 2) umma_commit_2sm(bar, 1);
            else umma_commit(bar);
          }
          if (++stage == kStages) stage = 0;
        }
        mbar_wait(bar, 0);
        t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
      }
    } else {
      if (warp == 1) {
        t0 = clock64();
        int stage = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          const uint32_t sa = base + stage * kStageBytes;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + 16384);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CG == 2) umma_bf16_2sm(tmem_base, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(tmem_base, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (kb == kblocks - 1) {
              if (CG == 2) umma_commit_2sm(bar, 1);
              else umma_commit(bar);
            }
          }
          __syncwarp();
          if (++stage == kStages) stage = 0;
        }
        mbar_wait(bar, 0);
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles_out[blockIdx.x] = t1 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

template <int CG, int STYLE>
static void run(int N, int kblocks, int grid, long long* d_cycles) {
  auto kern = umma_bench<CG, STYLE>;
  const int smem = kStages * kStageBytes + 1024 + 256;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, N, kblocks, d_cycles);
    cudaEventRecord(e1);
    if (e != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  const double mmas = 4.0 * kblocks;
  const double flops = 2.0 * 128.0 * N * 16.0 * mmas * grid;  // per CTA: 128 rows of the (128*CG) x N x 16 MMA
  printf("{\"cta_group\": %d, \"style\": \"%s\", \"N\": %d, \"grid\": %d, \"mmas_per_cta\": %.0f, \"cycles_per_mma\": %.1f, "
         "\"ideal_cycles\": %.0f, \"ms\": %.4f, \"tflops\": %.1f}\n",
         CG, STYLE == 0 ? "lane0-divergent" : "elect-uniform", N, grid, mmas, (double)h[0] / mmas, N / 2.0, best,
         flops / (best * 1e-3) / 1e12);
}

int main(int argc, char** argv) {
  const int kblocks = argc > 1 ? atoi(argv[1]) : 4096;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(long long) * 1024);
  const int Ns[] = {64, 96, 128, 192, 256};
  for (int grid : {2, sms & ~1}) {
    for (int N : Ns) {
      run<1, 0>(N, kblocks, grid, d_cycles);
      run<1, 1>(N, kblocks, grid, d_cycles);
      run<2, 0>(N, kblocks, grid, d_cycles);
      run<2, 1>(N, kblocks, grid, d_cycles);
    }
  }
  return 0;
}
