// Minimal host stand-in for <cuda_bf16.h>: the storage type and the round-to-nearest-even conversions fsnet_b200/csrc/act_tc.cu uses.
#pragma once
#include <cstdint>
#include <cstring>

struct __nv_bfloat16 {
  uint16_t bits;
};
struct __nv_bfloat162 {
  __nv_bfloat16 x, y;
};
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  __nv_bfloat16 r;
  if ((u & 0x7fffffffu) > 0x7f800000u) { r.bits = 0x7fff; return r; }            // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                               // round to nearest even
  r.bits = (uint16_t)(u >> 16);
  return r;
}
static inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = (uint32_t)h.bits << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return {__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
static inline float __low2float(__nv_bfloat162 v) { return __bfloat162float(v.x); }
static inline float __high2float(__nv_bfloat162 v) { return __bfloat162float(v.y); }
