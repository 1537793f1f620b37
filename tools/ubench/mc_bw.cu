// Micro-benchmark: L2 -> shared-memory delivery rate of TMA tile loads on B200, unicast against cluster multicast.
// Question it answers for csrc/conv_tc.cu: the forward convolutions deliver ~9.4 TB/s of TMA bytes into the SMs and sit there;
// does multicasting the operand that CTAs share (weights across pixel tiles, pixels across channel tiles) lift that bound?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mc_bw mc_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int kStages = 6;
constexpr int kTileRows = 128;              // 128 rows x 64 bf16 = 16 KB
constexpr uint32_t kTileBytes = kTileRows * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n"
               ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mode 0: unicast, every CTA loads the same tile sequence          (weights without multicast)
// mode 1: unicast, every CTA loads its own tile sequence            (pixels)
// mode 2: multicast, the cluster's CTAs each load 1/C of the tile and multicast it; all clusters the same sequence
// mode 3: multicast, each cluster its own sequence
// mode 5: multicast of whole tiles, issued by the cluster's CTAs in turn (what conv_tc.cu does)
// mode 4: half the stages like mode 1 (own tiles, unicast), half like mode 2 (shared tiles, multicast): the convolution's mix
__global__ void __launch_bounds__(64, 1) bw_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_slice,
                                                   int mode, int C, int iters, int n_tiles, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t rank = cluster_rank(), cid = cluster_id();
  const bool any_mc = mode >= 2;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], any_mc ? C : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const int slice_rows = kTileRows / C;
  const uint16_t mask = (uint16_t)((1u << C) - 1);
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_expect_tx(&full_bar[stage], kTileBytes);
      const bool mc = mode == 2 || mode == 3 || (mode == 4 && (it & 1));
      int tile;
      if (mode == 0 || mode == 2 || (mode == 4 && (it & 1))) tile = (it * 37) % n_tiles;
      else if (mode == 1 || mode == 4) tile = (int)(((long long)blockIdx.x * 53 + (long long)it * 37) % n_tiles);
      else tile = (int)(((long long)cid * 53 + (long long)it * 37) % n_tiles);
      uint8_t* st = smem + (size_t)stage * kTileBytes;
      if (mode == 5) { if ((uint32_t)(it % C) == rank) tma_2d_mc(st, &map_full, &full_bar[stage], 0, ((it * 37) % n_tiles) * kTileRows, mask); }
      else if (mc) tma_2d_mc(st + rank * slice_rows * 128, &map_slice, &full_bar[stage], 0, tile * kTileRows + rank * slice_rows, mask);
      else tma_2d(st, &map_full, &full_bar[stage], 0, tile * kTileRows);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x >= 32) {
    // consumer warp: every lane waits, lane c releases the stage in CTA c (the remote arrivals go out in parallel)
    const int lane = threadIdx.x - 32;
    int stage = 0; uint32_t phase = 0;
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + (size_t)stage * kTileBytes + 8 * (it & 63));
      __syncwarp();
      if (any_mc) { if (lane < C) mbar_arrive_remote(&empty_bar[stage], lane); }
      else if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[stage])) : "memory");
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x1234567ull) *sink = acc;
  }
  __syncthreads();
  cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int n_tiles = argc > 1 ? atoi(argv[1]) : 2048;     // 2048 x 16 KB = 32 MB: L2-resident
  const int iters = 4000;
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  __nv_bfloat16* buf; CK(cudaMalloc(&buf, (size_t)n_tiles * kTileBytes)); CK(cudaMemset(buf, 1, (size_t)n_tiles * kTileBytes));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTileBytes + 1024));
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int C : {1, 2, 4, 8}) {
    CUtensorMap mfull, mslice;
    cuuint64_t dim[2] = {64, (cuuint64_t)n_tiles * kTileRows}; cuuint64_t str[1] = {128};
    cuuint32_t es[2] = {1, 1};
    cuuint32_t boxf[2] = {64, (cuuint32_t)kTileRows}, boxs[2] = {64, (cuuint32_t)(kTileRows / C)};
    if (enc(&mfull, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dim, str, boxf, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    if (enc(&mslice, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dim, str, boxs, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    for (int mode = 0; mode < 6; ++mode) {
      if (C == 1 && mode >= 2) continue;
      const int grid = (148 / C) * C;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = kStages * kTileBytes + 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, bw_kernel, mfull, mslice, mode, C, iters, n_tiles, sink));
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const double bytes = (double)grid * iters * kTileBytes;
      printf("cluster %d mode %d grid %d: %.3f ms, delivered %.2f TB/s (%.1f B/clk/SM at 1.965 GHz)\n", C, mode, grid, best, bytes / best * 1e-9,
             bytes / best * 1e-9 * 1e12 / 148 / 1.965e9);
    }
  }
  return 0;
}
