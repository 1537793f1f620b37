// A small SIMT runtime for running CUDA kernels that use only threads, warp shuffles, shared memory, barriers and atomics on the
// CPU: every CUDA thread of a block is a fiber (ucontext), blocks run one after the other, warp / block collectives are barriers
// between fibers.  TEST INFRASTRUCTURE: lets the loss-side kernels of fsnet_b200/csrc be checked against the oracle in the CPU
// suite (tests/host_emulation/emulate.py rewrites `kernel<<<grid, block, smem, stream>>>(args)` into simt::launch calls).
// Not emulated (and not needed by those files): tensor cores, TMA, clusters, dynamic shared memory, textures, streams (all work
// is synchronous), fast-math approximations (the accurate libm functions stand in for __expf & co).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __cluster_dims__(...)
#define __noinline__
#define FSNET_HOST_PLAN_ONLY 1
#define __shared__ static

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static const int warpSize = 32;
// runtime calls of the host-side planning code (conv_tc.cu): a 148-SM device, attributes always accepted
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributeNonPortableClusterSizeAllowed = 9 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
static const unsigned long long cudaEnableDefault = 0;
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
#define __grid_constant__

namespace simt {

// Context switch between fibers.  glibc's swapcontext saves the signal mask with a system call on every switch (~1 us); on x86-64
// a few callee-saved registers and the stack pointer are all that is needed (tests/host_emulation/simt_switch.cpp, ~10 ns).
#if defined(__x86_64__)
#define SIMT_FAST_SWITCH 1
extern "C" void simt_switch(void** save_sp, void* new_sp);
#else
#define SIMT_FAST_SWITCH 0
#endif

struct Fiber {
  ucontext_t ctx;
  void* sp = nullptr;
  dim3 tid;
  int lane = 0, warp = 0;
  bool done = false;
};
struct Warp {
  uint64_t slot[32];
  int arrived = 0, live = 0;
  unsigned phase = 0;
};
struct Block {
  dim3 bid, bdim, gdim;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  int arrived = 0, live = 0;
  unsigned phase = 0;
  std::function<void()> body;
  ucontext_t main;
  void* main_sp = nullptr;
};
inline Block* blk = nullptr;
inline Fiber* cur = nullptr;
inline std::vector<std::vector<char>> stacks;
inline std::vector<char> dyn_smem_buf;                     // `extern __shared__` of the running launch (emulate.py rewrites the declaration)
inline void* dyn_smem() { return dyn_smem_buf.data(); }
constexpr size_t kStack = 512 * 1024;

inline void to_main() {
#if SIMT_FAST_SWITCH
  simt_switch(&cur->sp, blk->main_sp);
#else
  swapcontext(&cur->ctx, &blk->main);
#endif
}
inline void to_fiber(Fiber* f) {
  cur = f;
#if SIMT_FAST_SWITCH
  simt_switch(&blk->main_sp, f->sp);
#else
  swapcontext(&blk->main, &f->ctx);
#endif
}
inline void yield() { to_main(); }

inline void trampoline() {
  blk->body();
  cur->done = true;
  blk->live--;
  blk->warps[cur->warp].live--;
  to_main();
  abort();                                                 // a finished fiber is never resumed
}

// Barrier over the live threads of a warp / of the block.  Whoever sees the count complete (the last arriver, or a waiter after
// another thread has exited) opens it.
inline void warp_barrier() {
  Warp& w = blk->warps[cur->warp];
  const unsigned ph = w.phase;
  w.arrived++;
  while (w.phase == ph) {
    if (w.arrived >= w.live) { w.arrived = 0; w.phase++; break; }
    yield();
  }
}
inline void block_barrier() {
  const unsigned ph = blk->phase;
  blk->arrived++;
  while (blk->phase == ph) {
    if (blk->arrived >= blk->live) { blk->arrived = 0; blk->phase++; break; }
    yield();
  }
}

template <class T>
inline T exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
  Warp& w = blk->warps[cur->warp];
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  w.slot[cur->lane] = bits;
  warp_barrier();
  T r;
  std::memcpy(&r, &w.slot[src_lane & 31], sizeof(T));
  warp_barrier();
  return r;
}

template <class Body>
inline void run_block(dim3 gdim, dim3 bid, dim3 bdim, Body&& body) {
  Block b;
  b.bid = bid; b.bdim = bdim; b.gdim = gdim;
  const int n = (int)(bdim.x * bdim.y * bdim.z);
  b.fibers.resize(n);
  b.warps.resize((n + 31) / 32);
  b.live = n;
  b.body = body;
  if ((int)stacks.size() < n) stacks.resize(n);
  Block* outer_blk = blk;
  Fiber* outer_cur = cur;
  blk = &b;
  for (int i = 0; i < n; ++i) {
    Fiber& f = b.fibers[i];
    f.tid = dim3(i % bdim.x, (i / bdim.x) % bdim.y, i / (bdim.x * bdim.y));
    f.lane = i % 32;
    f.warp = i / 32;
    b.warps[f.warp].live++;
    if (stacks[i].empty()) stacks[i].resize(kStack);
#if SIMT_FAST_SWITCH
    // initial frame: six callee-saved registers (popped by simt_switch), then the entry address it returns into, then one
    // alignment slot so that the entry sees the stack as after a call (rsp = 16n + 8)
    uintptr_t top = ((uintptr_t)stacks[i].data() + kStack) & ~(uintptr_t)15;
    void** frame = (void**)(top - 64);
    for (int k = 0; k < 6; ++k) frame[k] = nullptr;
    frame[6] = (void*)trampoline;
    frame[7] = nullptr;
    f.sp = frame;
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks[i].data();
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &b.main;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
#endif
  }
  while (b.live > 0)
    for (int i = 0; i < n; ++i)
      if (!b.fibers[i].done) to_fiber(&b.fibers[i]);
  blk = outer_blk;
  cur = outer_cur;
}

template <class Kernel, class... Args>
inline void launch(Kernel kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  const char* plan_only = getenv("FSNET_EMULATE_PLAN_ONLY");     // read per launch: tests switch it on and off within one process
  if (plan_only && plan_only[0] == '1') return;                                   // walking a big configuration through the planners only: values are garbage
  if (dyn_smem_buf.size() < smem_bytes + 16) dyn_smem_buf.resize(smem_bytes + 16);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) run_block(grid, dim3(x, y, z), block, [&] { kernel(args...); });
}

template <class... A> inline void no_launch(A...) {}      // plan-only translation: the kernel is not run

}  // namespace simt

#define threadIdx (simt::cur->tid)
#define blockIdx (simt::blk->bid)
#define blockDim (simt::blk->bdim)
#define gridDim (simt::blk->gdim)

static inline void __syncthreads() { simt::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return simt::exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return simt::exchange(v, simt::cur->lane ^ m); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  const int lane = simt::cur->lane;
  return simt::exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  const int lane = simt::cur->lane;
  return simt::exchange(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned bits = 0;
  for (int l = 0; l < 32; ++l) bits |= (simt::exchange(pred ? 1u : 0u, l) & 1u) << l;      // slow and simple
  return bits;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
  const unsigned live = __ballot_sync(m, 1);
  return __ballot_sync(m, pred) == live;
}

template <class T> static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
static inline float atomicAdd(float* p, double v) { float old = *p; *p = old + (float)v; return old; }
template <class T> static inline T atomicMax(T* p, T v) { T old = *p; *p = std::max(old, v); return old; }
template <class T> static inline T atomicMin(T* p, T v) { T old = *p; *p = std::min(old, v); return old; }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __frcp_rn(float x) { return 1.f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __saturatef(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
// __expf / __logf / __powf clash with glibc's internal declarations: emulate.py rewrites them to expf / logf / powf
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline double __dsqrt_rn(double x) { return sqrt(x); }
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline long long __double2ll_rn(double a) { return llrint(a); }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }

// CUDA's overload set of min / max (std::min would reject mixed integer types)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
