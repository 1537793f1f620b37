// Shared helpers for the fsnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fsnet_b200.h"

namespace fsnet {

void set_error(const char* fmt, ...);

#define FSNET_REQUIRE(cond, ...)                         \
  do {                                                   \
    if (!(cond)) {                                       \
      ::fsnet::set_error(__VA_ARGS__);                   \
      return FSNET_ERR_INVALID;                          \
    }                                                    \
  } while (0)

#define FSNET_CUDA_OK(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::fsnet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FSNET_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

#define FSNET_LAUNCH_OK()                                                                      \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      ::fsnet::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FSNET_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float ldg(const float* p) { return __ldg(p); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch ------------------------------------------------------------------------------------------
// A training step is ~330 short launches in one stream (one CUDA graph): the launch latency, CTA ramp and prologue (barrier
// initialisation, TMEM allocation, tensor-map prefetch) of kernel k+1 can overlap the tail of kernel k.  Kernels launched through
// FSNET_LAUNCH_PDL may start while their predecessor is still running; they call pdl_launch_dependents() early (their successor
// may be scheduled as SM resources free up) and pdl_wait() before their first access to global memory (returns once every
// preceding grid has completed and its writes are visible).  Launched normally both calls are no-ops.  FSNET_PDL=0 disables it.
#ifdef FSNET_HOST_PLAN_ONLY
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
#define FSNET_LAUNCH_PDL(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);      // errors surface through cudaGetLastError() (FSNET_LAUNCH_OK)
}
#define FSNET_LAUNCH_PDL(kern, grid, block, smem, stream, ...) ::fsnet::launch_pdl(kern, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)
#endif

}  // namespace fsnet
