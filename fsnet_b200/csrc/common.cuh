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

}  // namespace fsnet
