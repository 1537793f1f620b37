// Micro-benchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, both operands from shared memory, K-major,
// 128-byte swizzle) as a function of N -- the cost model behind csrc/conv_tc.cu's tile shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: every MMA reads the same A / B K-slices; mode 1: walks 4 K-slices of one tile (the convolution's pattern);
// mode 2: like 1 but alternating between two accumulators; mode 3: like 1 with 3 distinct A/B tile pairs per round (hi/lo products)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (3 * 16384 + 3 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 3 * 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int pair = mode == 3 ? it % 3 : 0;
      const uint64_t ad = make_desc(a0 + pair * 16384, 1024), bd = make_desc(b0 + pair * 32768, 1024);
      const uint32_t d = tmem + ((mode == 2 && (it & 1)) ? 256u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t kk = mode == 0 ? 0 : 2 * k;
        asm volatile("{\n.reg .pred pe, pa;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %4, 0;\n@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n}\n"
                     ::"r"(d), "l"(ad + kk), "l"(bd + kk), "r"(idesc), "r"((it | k) ? 1u : 0u) : "memory");
      }
    }
    asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    while (!mbar_try_wait(&bar, 0)) {}
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
int main() {
  const int smem = 3 * 16384 + 3 * 32768 + 1024, iters = 2000;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d; CK(cudaMalloc(&d, 148 * 8));
  long long h[148];
  for (int grid : {1, 148}) for (int mode = 0; mode < 4; ++mode) for (int N : {16, 32, 64, 128, 256}) {
    for (int rep = 0; rep < 2; ++rep) { rate_kernel<<<grid, 128, smem>>>(N, iters, mode, d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / (4.0 * iters);
    printf("grid %3d mode %d N %3d: %.1f cycles per MMA (floor N/2 = %d) -> %.0f%% of the dense rate\n", grid, mode, N, per, N / 2, 100.0 * (N / 2.0) / per);
  }
  return 0;
}
