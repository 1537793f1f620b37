// Element-wise / reduction kernels around the tcgen05 convolutions: image -> bf16 planes, BatchNorm
// finalisation, BN-apply + residual + ReLU + hi/lo split (+ nearest x2 up-sampling into a concat
// buffer), max-pool, BatchNorm backward (reduce + apply), ring folding, weight re-layout.
//
// "planes": an activation x is stored as two NHWC bf16 tensors hi = bf16(x), lo = bf16(x - hi) inside a
// buffer with a one-pixel ring (replicate padding of the interior), so that tcgen05 convolutions read
// it directly with TMA.  All views are described by fsnet_view (include/fsnet_b200.h).
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace fsnet {
namespace {

typedef fsnet_view V;

__device__ __forceinline__ size_t vidx(const V& v, int n, int y, int x, int c) {
  const int pw = v.w + 2 * v.ring, ph = v.h + 2 * v.ring;
  return (((size_t)n * ph + (y + v.ring)) * pw + (x + v.ring)) * v.c_total + v.c_off + c;
}
__device__ __forceinline__ size_t plane_stride(const V& v) {
  return (size_t)v.n * (v.h + 2 * v.ring) * (v.w + 2 * v.ring) * v.c_total;
}
__device__ __forceinline__ void split_store(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t i, float x) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}
__device__ __forceinline__ float plane_load(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t i) {
  return __bfloat162float(hi[i]) + (lo ? __bfloat162float(lo[i]) : 0.f);
}

// writes value x of interior pixel (y,x) and its replicate copies in the ring
__device__ __forceinline__ void store_with_ring(const V& d, __nv_bfloat16* hi, __nv_bfloat16* lo, int n, int y, int x, int c, float val) {
  split_store(hi, lo, vidx(d, n, y, x, c), val);
  if (d.ring == 0) return;
  const bool l = x == 0, r = x == d.w - 1, t = y == 0, b = y == d.h - 1;
  if (l) split_store(hi, lo, vidx(d, n, y, -1, c), val);
  if (r) split_store(hi, lo, vidx(d, n, y, d.w, c), val);
  if (t) split_store(hi, lo, vidx(d, n, -1, x, c), val);
  if (b) split_store(hi, lo, vidx(d, n, d.h, x, c), val);
  if (l && t) split_store(hi, lo, vidx(d, n, -1, -1, c), val);
  if (r && t) split_store(hi, lo, vidx(d, n, -1, d.w, c), val);
  if (l && b) split_store(hi, lo, vidx(d, n, d.h, -1, c), val);
  if (r && b) split_store(hi, lo, vidx(d, n, d.h, d.w, c), val);
}


// ---- 8-channel vector helpers (all channel counts on this path are multiples of 16) ------------------------
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld_f8(const float* p) {
  F8 r; float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w; return r;
}
__device__ __forceinline__ void st_f8(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ F8 ld_bf8(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r; const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = __uint_as_float(w[i] << 16); r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  return r;
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// splits 8 floats into hi / lo bf16 and stores them (lo may be null)
__device__ __forceinline__ void st_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t i, const F8& x) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    h[k] = pack_bf2(x.v[2 * k], x.v[2 * k + 1]);
    float h0 = __uint_as_float(h[k] << 16), h1 = __uint_as_float(h[k] & 0xffff0000u);
    l[k] = pack_bf2(x.v[2 * k] - h0, x.v[2 * k + 1] - h1);
  }
  *reinterpret_cast<uint4*>(hi + i) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + i) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void store_with_ring8(const V& d, __nv_bfloat16* hi, __nv_bfloat16* lo, int n, int y, int x, int c, const F8& val) {
  st_split8(hi, lo, vidx(d, n, y, x, c), val);
  if (d.ring == 0) return;
  const bool l = x == 0, r = x == d.w - 1, t = y == 0, b = y == d.h - 1;
  if (!(l | r | t | b)) return;
  if (l) st_split8(hi, lo, vidx(d, n, y, -1, c), val);
  if (r) st_split8(hi, lo, vidx(d, n, y, d.w, c), val);
  if (t) st_split8(hi, lo, vidx(d, n, -1, x, c), val);
  if (b) st_split8(hi, lo, vidx(d, n, d.h, x, c), val);
  if (l && t) st_split8(hi, lo, vidx(d, n, -1, -1, c), val);
  if (r && t) st_split8(hi, lo, vidx(d, n, -1, d.w, c), val);
  if (l && b) st_split8(hi, lo, vidx(d, n, d.h, -1, c), val);
  if (r && b) st_split8(hi, lo, vidx(d, n, d.h, d.w, c), val);
}
// 32-bit decode of a flat (pixel, 8-channel group) index
struct Px { int n, y, x, c; };
__device__ __forceinline__ Px decode8(unsigned i, int H, int W, int C) {
  const unsigned g = (unsigned)C >> 3;
  Px p; unsigned pix = i / g; p.c = (int)(i - pix * g) << 3;
  unsigned t = pix / (unsigned)W; p.x = (int)(pix - t * (unsigned)W);
  unsigned n = t / (unsigned)H; p.y = (int)(t - n * (unsigned)H); p.n = (int)n;
  return p;
}

// ---- image (fp32 NCHW) -> planes with zero-filled extra channels ----------------------------------
// One thread per (PADDED pixel, 8-channel group): coalesced reads along x, one 128-bit store per plane.  The ring
// (any width) is either the replicate copy of the border or zeros (zero_ring: the zero padding of the 7x7 stem,
// materialised so that the stem can run on the folded-tap convolution path).
__global__ void image_to_planes_kernel(const float* __restrict__ img, int C, V d, int zero_ring) {
  const int ph = d.h + 2 * d.ring, pw = d.w + 2 * d.ring;
  const unsigned total = (unsigned)d.n * ph * pw * (d.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  Px q = decode8(i, ph, pw, d.c);
  int y = q.y - d.ring, x = q.x - d.ring;
  const bool outside = y < 0 || y >= d.h || x < 0 || x >= d.w;
  y = min(max(y, 0), d.h - 1); x = min(max(x, 0), d.w - 1);
  F8 v;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = q.c + k;
    v.v[k] = (c < C && !(outside && zero_ring)) ? __ldg(img + (((size_t)q.n * C + c) * d.h + y) * d.w + x) : 0.f;
  }
  __nv_bfloat16* hi = (__nv_bfloat16*)d.ptr;
  st_split8(hi, hi + plane_stride(d), vidx(d, q.n, q.y - d.ring, q.x - d.ring, q.c), v);
}

// ---- BatchNorm finalisation -------------------------------------------------------------------------
// training: batch statistics from the conv epilogue's fp64 sums; updates running stats (momentum, unbiased
// variance) exactly like nn.BatchNorm2d; the conv bias (if any) shifts the mean only.
__global__ void bn_finalize_kernel(double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches, float momentum, float eps, int training, int C,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
  pdl_launch_dependents();
  pdl_wait();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      float bias = conv_bias ? conv_bias[c] : 0.f;
      double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + bias);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    stats[c] = 0.0; stats[C + c] = 0.0;
    if (c == 0 && num_batches) *num_batches += 1;
  } else {
    float bias = conv_bias ? conv_bias[c] : 0.f;
    mean = running_mean[c] - bias;            // raw conv output excludes the bias
    invstd = rsqrtf(running_var[c] + eps);
  }
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[c] = g * invstd;
  scale_shift[C + c] = b - mean * g * invstd;
  if (mean_invstd) { mean_invstd[c] = mean; mean_invstd[C + c] = invstd; }
}

// ---- y = relu?(raw*scale + shift + residual) -> planes (optionally nearest x2 into a channel slice) -----
struct ActParams {
  V raw;                        // fp32, ring 0
  const float* ss;              // scale[C], shift[C] (or null: identity)
  int res_mode;                 // 0 none, 1 planes, 2 raw fp32 with its own scale/shift
  V res; const float* res_ss;
  int relu, up;
  V dst;                        // planes
};
__global__ void __launch_bounds__(256) act_planes_kernel(ActParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const V& s = p.raw;
  const unsigned total = (unsigned)s.n * s.h * s.w * (s.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const Px q = decode8(i, s.h, s.w, s.c);
  F8 v = ld_f8((const float*)s.ptr + vidx(s, q.n, q.y, q.x, q.c));
  if (p.ss) {
    F8 sc = ld_f8(p.ss + q.c), sh = ld_f8(p.ss + s.c + q.c);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = fmaf(v.v[k], sc.v[k], sh.v[k]);
  }
  if (p.res_mode == 1) {
    const __nv_bfloat16* rh = (const __nv_bfloat16*)p.res.ptr;
    const size_t ri = vidx(p.res, q.n, q.y, q.x, q.c);
    F8 a = ld_bf8(rh + ri), b = ld_bf8(rh + ri + plane_stride(p.res));
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] += a.v[k] + b.v[k];
  } else if (p.res_mode == 2) {
    F8 rv = ld_f8((const float*)p.res.ptr + vidx(p.res, q.n, q.y, q.x, q.c));
    F8 sc = ld_f8(p.res_ss + q.c), sh = ld_f8(p.res_ss + s.c + q.c);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] += fmaf(rv.v[k], sc.v[k], sh.v[k]);
  }
  if (p.relu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = fmaxf(v.v[k], 0.f);
  }
  __nv_bfloat16* hi = (__nv_bfloat16*)p.dst.ptr;
  __nv_bfloat16* lo = hi + plane_stride(p.dst);
  if (p.up == 1) {
    store_with_ring8(p.dst, hi, lo, q.n, q.y, q.x, q.c, v);
  } else {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) store_with_ring8(p.dst, hi, lo, q.n, 2 * q.y + dy, 2 * q.x + dx, q.c, v);
  }
}

// ---- planes -> planes channel-slice copy (skip connection into the concat buffer), ring included --------
__global__ void __launch_bounds__(256) copy_planes_kernel(V s, V d) {
  const int ph = s.h + 2 * s.ring, pw = s.w + 2 * s.ring;
  const unsigned total = (unsigned)s.n * ph * pw * (s.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  Px q = decode8(i, ph, pw, s.c);
  q.y -= s.ring; q.x -= s.ring;
  const __nv_bfloat16* sh = (const __nv_bfloat16*)s.ptr;
  __nv_bfloat16* dh = (__nv_bfloat16*)d.ptr;
  const size_t si = vidx(s, q.n, q.y, q.x, q.c), di = vidx(d, q.n, q.y, q.x, q.c);
  *reinterpret_cast<uint4*>(dh + di) = *reinterpret_cast<const uint4*>(sh + si);
  *reinterpret_cast<uint4*>(dh + di + plane_stride(d)) = *reinterpret_cast<const uint4*>(sh + si + plane_stride(s));
}

// ---- 3x3 / stride 2 / pad 1 max-pool on planes ---------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_planes_kernel(V s, V d, uint8_t* __restrict__ argmax) {
  const unsigned total = (unsigned)d.n * d.h * d.w * (d.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const Px q = decode8(i, d.h, d.w, d.c);
  const __nv_bfloat16* sh = (const __nv_bfloat16*)s.ptr;
  const __nv_bfloat16* sl = sh + plane_stride(s);
  F8 m; int am[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { m.v[k] = -INFINITY; am[k] = 0; }
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = 2 * q.y + dy, xx = 2 * q.x + dx;
      if (yy < 0 || yy >= s.h || xx < 0 || xx >= s.w) continue;
      const size_t si = vidx(s, q.n, yy, xx, q.c);
      F8 a = ld_bf8(sh + si), b = ld_bf8(sl + si);
      const int pos = (dy + 1) * 3 + dx + 1;
#pragma unroll
      for (int k = 0; k < 8; ++k) { float v = a.v[k] + b.v[k]; if (v > m.v[k]) { m.v[k] = v; am[k] = pos; } }   // first maximum wins (PyTorch)
    }
  __nv_bfloat16* dh = (__nv_bfloat16*)d.ptr;
  store_with_ring8(d, dh, dh + plane_stride(d), q.n, q.y, q.x, q.c, m);
  if (argmax) {
    uint2 pk = make_uint2(am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24), am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24));
    *reinterpret_cast<uint2*>(argmax + (((size_t)q.n * d.h + q.y) * d.w + q.x) * d.c + q.c) = pk;
  }
}
// backward: every source pixel looks at the (up to 4) windows that contain it and takes their gradient where the
// stored arg-max points at it.  g_src (+)= ...
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(V s, const uint8_t* __restrict__ argmax, V gd, V gs, int accumulate) {
  const unsigned total = (unsigned)s.n * s.h * s.w * (s.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const Px q = decode8(i, s.h, s.w, s.c);
  F8 g;
#pragma unroll
  for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
  for (int oy = q.y / 2; oy <= (q.y + 1) / 2; ++oy) {                 // windows with 2*oy-1 <= y <= 2*oy+1
    if (oy >= gd.h) continue;
    for (int ox = q.x / 2; ox <= (q.x + 1) / 2; ++ox) {
      if (ox >= gd.w) continue;
      const int pos = (q.y - 2 * oy + 1) * 3 + (q.x - 2 * ox + 1);
      const uint2 pk = *reinterpret_cast<const uint2*>(argmax + (((size_t)q.n * gd.h + oy) * gd.w + ox) * gd.c + q.c);
      const F8 gv = ld_f8((const float*)gd.ptr + vidx(gd, q.n, oy, ox, q.c));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int a = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xff;
        if (a == pos) g.v[k] += gv.v[k];
      }
    }
  }
  float* o = (float*)gs.ptr + vidx(gs, q.n, q.y, q.x, q.c);
  if (accumulate) { F8 old = ld_f8(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] += old.v[k]; }
  st_f8(o, g);
}

// ---- BatchNorm (+ReLU mask) backward ------------------------------------------------------------------------------
struct BnBwdParams {
  V g;                          // incoming gradient wrt the activation (fp32); `up`=2: adjoint of nearest x2
  int up;
  V mask; int has_mask;         // activation planes (ReLU output): gradient passes where value > 0
  const float* mask_ss;         // alternative mask: raw*scale+shift > 0 (activation only stored up-sampled)
  V raw;                        // conv raw output (fp32)
  const float* mean_invstd;     // [2C] or null (no BN: plain masked gradient)
  const float* gamma;           // [C] or null
  double* sums;                 // [2C]: sum g, sum g*xhat
  double count;
  V dy;                         // out: bf16 plane (hi only), ring zeroed
  V res; int res_mode;          // 0 none, 1 write masked g, 2 accumulate masked g   (identity residual / downsample input)
  float* dgamma; float* dbeta; int c_real;   // optional fp32 parameter gradients (= sums[C..], sums[..C]) written by block 0
};
__device__ __forceinline__ F8 masked_g8(const BnBwdParams& p, const Px& q, const F8& rawv) {
  const float* gp = (const float*)p.g.ptr;
  F8 g;
  if (p.up == 1) {
    g = ld_f8(gp + vidx(p.g, q.n, q.y, q.x, q.c));
  } else {
    g = ld_f8(gp + vidx(p.g, q.n, 2 * q.y, 2 * q.x, q.c));
    F8 b = ld_f8(gp + vidx(p.g, q.n, 2 * q.y, 2 * q.x + 1, q.c)), c = ld_f8(gp + vidx(p.g, q.n, 2 * q.y + 1, 2 * q.x, q.c)),
       d = ld_f8(gp + vidx(p.g, q.n, 2 * q.y + 1, 2 * q.x + 1, q.c));
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] += b.v[k] + c.v[k] + d.v[k];
  }
  if (p.has_mask) {
    const __nv_bfloat16* mh = (const __nv_bfloat16*)p.mask.ptr;
    const size_t mi = vidx(p.mask, q.n, q.y, q.x, q.c);
    F8 a = ld_bf8(mh + mi), b = ld_bf8(mh + mi + plane_stride(p.mask));
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = (a.v[k] + b.v[k]) > 0.f ? g.v[k] : 0.f;
  } else if (p.mask_ss) {
    F8 sc = ld_f8(p.mask_ss + q.c), sh = ld_f8(p.mask_ss + p.raw.c + q.c);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = fmaf(rawv.v[k], sc.v[k], sh.v[k]) > 0.f ? g.v[k] : 0.f;
  }
  return g;
}
// Each thread owns one 8-channel group and walks pixels; per-block partial sums go through shared-memory
// float atomics, then one fp64 atomic per channel and block.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(BnBwdParams p, int pix_per_iter, int threads_used) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_acc[];           // [2*C]
  const V& r = p.raw;
  const int C = r.c, groups = C >> 3;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  if ((int)threadIdx.x < threads_used) {
    const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
    const int c = cg << 3;
    const unsigned npix = (unsigned)r.n * r.h * r.w;
    F8 mean, inv;
#pragma unroll
    for (int k = 0; k < 8; ++k) { mean.v[k] = 0.f; inv.v[k] = 0.f; }
    if (p.mean_invstd) { mean = ld_f8(p.mean_invstd + c); inv = ld_f8(p.mean_invstd + C + c); }
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    // two pixels per iteration: their loads are independent, which doubles the bytes in flight per thread
    const unsigned step = gridDim.x * pix_per_iter;
    for (unsigned pix = blockIdx.x * pix_per_iter + pl; pix < npix; pix += 2 * step) {
      const unsigned pix2 = pix + step;
      const bool has2 = pix2 < npix;
      Px q; unsigned t = pix / (unsigned)r.w; q.x = (int)(pix - t * r.w);
      unsigned n = t / (unsigned)r.h; q.y = (int)(t - n * r.h); q.n = (int)n; q.c = c;
      Px q2 = q;
      if (has2) { unsigned t2 = pix2 / (unsigned)r.w; q2.x = (int)(pix2 - t2 * r.w);
                  unsigned n2 = t2 / (unsigned)r.h; q2.y = (int)(t2 - n2 * r.h); q2.n = (int)n2; }
      const F8 rv = ld_f8((const float*)r.ptr + vidx(r, q.n, q.y, q.x, c));
      const F8 rv2 = ld_f8((const float*)r.ptr + vidx(r, q2.n, q2.y, q2.x, c));
      const F8 g = masked_g8(p, q, rv);
      const F8 g2 = masked_g8(p, q2, rv2);
      const float w2 = has2 ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s1[k] += g.v[k]; s2[k] = fmaf(g.v[k], (rv.v[k] - mean.v[k]) * inv.v[k], s2[k]);
        const float gg = g2.v[k] * w2;
        s1[k] += gg; s2[k] = fmaf(gg, (rv2.v[k] - mean.v[k]) * inv.v[k], s2[k]);
      }
    }
    // lanes of a warp that own the same channel group (lane % groups, when groups divides 32) combine by shuffles first:
    // for the 16..64-channel full-resolution layers the shared-memory atomics were 128-way contended
    if (groups < 32 && (groups & (groups - 1)) == 0 && threads_used == 256) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        for (int o = 16; o >= groups; o >>= 1) {
          s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
          s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
        }
      }
      if ((int)(threadIdx.x & 31) < groups) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { atomicAdd(&s_acc[c + k], s1[k]); atomicAdd(&s_acc[C + c + k], s2[k]); }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) { atomicAdd(&s_acc[c + k], s1[k]); atomicAdd(&s_acc[C + c + k], s2[k]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float v = s_acc[i];
    if (v != 0.f) atomicAdd(p.sums + i, (double)v);
  }
}
// Ringed dy (thin replicate layers: the data gradient reads the zero padding of k-1 pixels as data): the ring -- r rows above and
// below and r columns left and right of every image -- is written by the apply kernel instead of a torch fill of the whole
// two-plane buffer before the launch (97 MB per full-resolution 16-channel layer, ~70 us per step).  Not inlined: inlined, its
// index arithmetic took bn_bwd_apply_kernel from 48 to 100 registers and every layer's launch got 20-40 % slower.
__device__ __noinline__ void zero_dy_ring(const V dy, int N, int H, int W, int C) {
  const int rg = dy.ring, pw = W + 2 * rg;
  const unsigned top = (unsigned)(2 * rg * pw), side = (unsigned)(2 * rg * H), per_img = top + side, groups = (unsigned)(C >> 3);
  const unsigned ring_total = (unsigned)N * per_img * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < ring_total; i += gridDim.x * blockDim.x) {
    const unsigned cg = i % groups; unsigned px = i / groups;
    const int n = (int)(px / per_img); px -= (unsigned)n * per_img;
    int y, x;
    if (px < top) {
      const int row = (int)(px / pw);
      x = (int)(px - row * pw) - rg;
      y = row < rg ? row - rg : H + (row - rg);
    } else {
      px -= top;
      y = (int)(px / (2 * rg));
      const int xx = (int)(px - y * 2 * rg);
      x = xx < rg ? xx - rg : W + (xx - rg);
    }
    *reinterpret_cast<uint4*>((__nv_bfloat16*)dy.ptr + vidx(dy, n, y, x, (int)cg * 8)) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// dy = A * g + B * raw + K per channel: the BatchNorm input gradient gamma*inv*(g - mean(g) - xhat*mean(g*xhat)) with
// xhat = (raw - mean)*inv, rearranged so that an element costs two FMAs.  The three coefficients per channel are built once
// per block from the fp64 sums into shared memory (round 1 read 16 doubles and converted them per 8 elements: the kernel sat
// on the LSU queue, lg_throttle up to 3.9 per issue in ncu r2c8); persistent blocks walk the tensor.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(BnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_coef[];          // A[C], B[C], K[C]
  const V& r = p.raw;
  const int C = r.c;
  const unsigned total = (unsigned)r.n * r.h * r.w * (C >> 3);
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < p.c_real; c += blockDim.x) {
      if (p.dbeta) p.dbeta[c] = (float)p.sums[c];
      if (p.dgamma) p.dgamma[c] = (float)p.sums[C + c];
    }
  }
  const float rc = (float)(1.0 / p.count);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float A = 1.f, B = 0.f, K = 0.f;
    if (p.mean_invstd) {
      const float mean = p.mean_invstd[c], inv = p.mean_invstd[C + c];
      const float gam = p.gamma ? p.gamma[c] : 1.f;
      const float sg = (float)p.sums[c] * rc, sgx = (float)p.sums[C + c] * rc;
      A = gam * inv;
      B = -A * inv * sgx;
      K = -A * sg - B * mean;
    }
    s_coef[c] = A; s_coef[C + c] = B; s_coef[2 * C + c] = K;
  }
  __syncthreads();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const Px q = decode8(i, r.h, r.w, C);
    const F8 rv = ld_f8((const float*)r.ptr + vidx(r, q.n, q.y, q.x, q.c));
    const F8 g = masked_g8(p, q, rv);
    if (p.res_mode) {
      float* o = (float*)p.res.ptr + vidx(p.res, q.n, q.y, q.x, q.c);
      F8 w = g;
      if (p.res_mode == 2) { F8 old = ld_f8(o);
#pragma unroll
        for (int k = 0; k < 8; ++k) w.v[k] += old.v[k]; }
      st_f8(o, w);
    }
    const F8 A = ld_f8(s_coef + q.c), B = ld_f8(s_coef + C + q.c), K = ld_f8(s_coef + 2 * C + q.c);
    F8 d;
#pragma unroll
    for (int k = 0; k < 8; ++k) d.v[k] = fmaf(A.v[k], g.v[k], fmaf(B.v[k], rv.v[k], K.v[k]));
    st_split8((__nv_bfloat16*)p.dy.ptr, nullptr, vidx(p.dy, q.n, q.y, q.x, q.c), d);
  }
  if (p.dy.ring > 0) zero_dy_ring(p.dy, r.n, r.h, r.w, C);
}

// ---- ring folding: adjoint of replicate padding on a ringed fp32 gradient ----------------------------------------------
__global__ void fold_ring_kernel(V g) {
  // one thread per (n, border pixel, c); border pixels enumerated as top row, bottom row, then left/right columns
  const int per = 2 * g.w + 2 * (g.h - 2 > 0 ? g.h - 2 : 0);
  size_t total = (size_t)g.n * per * g.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % g.c); size_t t = i / g.c;
  int b = (int)(t % per); int n = (int)(t / per);
  int y, x;
  if (b < g.w) { y = 0; x = b; }
  else if (b < 2 * g.w) { y = g.h - 1; x = b - g.w; }
  else { int k = b - 2 * g.w; y = 1 + k / 2; x = (k & 1) ? g.w - 1 : 0; }
  if (g.h == 1 && b >= g.w) return;
  float* p = (float*)g.ptr;
  float acc = 0.f;
  const bool l = x == 0, r = x == g.w - 1, tp = y == 0, bt = y == g.h - 1;
  if (l) acc += p[vidx(g, n, y, -1, c)];
  if (r) acc += p[vidx(g, n, y, g.w, c)];
  if (tp) acc += p[vidx(g, n, -1, x, c)];
  if (bt) acc += p[vidx(g, n, g.h, x, c)];
  if (l && tp) acc += p[vidx(g, n, -1, -1, c)];
  if (r && tp) acc += p[vidx(g, n, -1, g.w, c)];
  if (l && bt) acc += p[vidx(g, n, g.h, -1, c)];
  if (r && bt) acc += p[vidx(g, n, g.h, g.w, c)];
  p[vidx(g, n, y, x, c)] += acc;
}

// dst (+)= src channel slice (both fp32 views, same N,H,W,C)
__global__ void __launch_bounds__(256) add_slice_kernel(V d, V s, int accumulate) {
  const unsigned total = (unsigned)d.n * d.h * d.w * (d.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const Px q = decode8(i, d.h, d.w, d.c);
  F8 v = ld_f8((const float*)s.ptr + vidx(s, q.n, q.y, q.x, q.c));
  float* o = (float*)d.ptr + vidx(d, q.n, q.y, q.x, q.c);
  if (accumulate) { F8 old = ld_f8(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] += old.v[k]; }
  st_f8(o, v);
}

// zero-insertion x2 of a bf16 plane (transposed stride-2 convolution as a stride-1 convolution)
__global__ void __launch_bounds__(256) zero_insert_kernel(V s, V d) {
  const unsigned total = (unsigned)d.n * d.h * d.w * (d.c >> 3);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const Px q = decode8(i, d.h, d.w, d.c);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (!(q.y & 1) && !(q.x & 1) && (q.y >> 1) < s.h && (q.x >> 1) < s.w)
    v = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)s.ptr + vidx(s, q.n, q.y >> 1, q.x >> 1, q.c));
  *reinterpret_cast<uint4*>((__nv_bfloat16*)d.ptr + vidx(d, q.n, q.y, q.x, q.c)) = v;
}

// weights fp32 [Cout,Cin,KH,KW] -> forward planes [Cout_pad,KH,KW,Cin_pad] (hi, lo) and, optionally, the
// dgrad operand [Cin_pad,KH,KW (both flipped),Cout_pad] (hi only)
__global__ void weight_planes_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad,
                                     __nv_bfloat16* __restrict__ fwd_hi, __nv_bfloat16* __restrict__ fwd_lo,
                                     __nv_bfloat16* __restrict__ dg_hi) {
  size_t total = (size_t)Cout_pad * KH * KW * Cin_pad;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ci = (int)(i % Cin_pad); size_t t = i / Cin_pad;
  int s = (int)(t % KW); t /= KW;
  int r = (int)(t % KH); int co = (int)(t / KH);
  float v = (co < Cout && ci < Cin) ? __ldg(w + (((size_t)co * Cin + ci) * KH + r) * KW + s) : 0.f;
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  fwd_hi[i] = h;
  if (fwd_lo) fwd_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  if (dg_hi) dg_hi[(((size_t)ci * KH + (KH - 1 - r)) * KW + (KW - 1 - s)) * Cout_pad + co] = h;
}

// all layers of a network in one launch: blockIdx.y selects the layer descriptor
// All layers in one launch (blockIdx.y = layer).  A block stages a [co tile][ci tile][taps] brick of the fp32 parameter
// (contiguous in memory) in shared memory and writes it out in both operand layouts with the destination's fastest
// index across threads: [co][tap][ci] for the forward planes (hi, lo) and [ci][flipped tap][co] for the data gradient.
// (The scalar version wrote the transposed layout with 2-byte scattered stores: 288 us per step at cfg2.)
constexpr int kWpTile = 4608;                // floats of shared memory per brick
// (TT = taps as a compile-time constant for the 3x3 and 1x1 layers, 0 = run-time: with run-time divisors the index arithmetic --
// six integer divisions per element -- made both batched re-layout kernels compute bound: 128 + 87 us per step for 117 + 102 MB)
template <int TT>
__device__ __forceinline__ void weight_planes_tiles(const fsnet_weight_desc& d, float* s_w) {
  const int T = TT ? TT : d.kh * d.kw;
  const int CI_T = d.cin_pad < 32 ? d.cin_pad : 32;          // 8, 16 or 32: a power of two
  int cs = 0;
  while ((1 << cs) < CI_T) ++cs;
  int CO_T = kWpTile / (CI_T * T);
  CO_T = CO_T > 16 ? 16 : (CO_T < 1 ? 1 : CO_T);
  if (TT) CO_T = 16;                                         // 4608 / (32 * 9) = 16 and more for thinner layers
  if (CI_T * T > kWpTile) return;            // host guarantees this never happens (taps <= 49, CI_T <= 32)
  const int ci_tiles = (d.cin_pad + CI_T - 1) / CI_T, co_tiles = (d.cout_pad + CO_T - 1) / CO_T;
  const int brick = CO_T * CI_T * T;
  __nv_bfloat16* fh = (__nv_bfloat16*)d.fwd_hi;
  __nv_bfloat16* fl = (__nv_bfloat16*)d.fwd_lo;
  __nv_bfloat16* dg = (__nv_bfloat16*)d.dgrad_hi;
  for (int tile = blockIdx.x; tile < ci_tiles * co_tiles; tile += gridDim.x) {
    const int co0 = (tile / ci_tiles) * CO_T, ci0 = (tile % ci_tiles) * CI_T;
    __syncthreads();
    for (int i = threadIdx.x; i < brick; i += 256) {
      const int q = i / T, co_l = q >> cs, ci_l = q & (CI_T - 1), rem = i - co_l * (CI_T * T);      // i = (co_l * CI_T + ci_l) * T + tap
      const int co = co0 + co_l, ci = ci0 + ci_l;
      s_w[i] = (co < d.cout && ci < d.cin) ? __ldg(d.w + ((size_t)co * d.cin + ci0) * T + rem) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < brick; i += 256) {         // forward layout: ci fastest
      const int ci_l = i & (CI_T - 1), t2 = i >> cs, co_l = t2 / T, tap = t2 - co_l * T;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (co < d.cout_pad && ci < d.cin_pad) {
        const float v = s_w[(co_l * CI_T + ci_l) * T + tap];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const size_t o = ((size_t)co * T + tap) * d.cin_pad + ci;
        fh[o] = h;
        if (fl) fl[o] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
    if (dg) {
      for (int i = threadIdx.x; i < brick; i += 256) {       // data-gradient layout: co fastest, taps flipped
        const int co_l = i % CO_T, t2 = i / CO_T, ci_l = t2 / T, tap = t2 - ci_l * T;
        const int co = co0 + co_l, ci = ci0 + ci_l;
        if (co < d.cout_pad && ci < d.cin_pad)
          dg[((size_t)ci * T + (T - 1 - tap)) * d.cout_pad + co] = __float2bfloat16_rn(s_w[(co_l * CI_T + ci_l) * T + tap]);
      }
    }
  }
}
__global__ void __launch_bounds__(256) weight_planes_batched_kernel(const fsnet_weight_desc* __restrict__ table) {
  __shared__ float s_w[kWpTile];
  const fsnet_weight_desc d = table[blockIdx.y];
  const int T = d.kh * d.kw;
  if (T == 9) weight_planes_tiles<9>(d, s_w);
  else if (T == 1) weight_planes_tiles<1>(d, s_w);
  else weight_planes_tiles<0>(d, s_w);
}
// All layers' accumulators -> parameter-gradient layout in one launch (blockIdx.y = layer): a [co tile][tap][ci tile]
// brick of the padded fp32 accumulator [Cout_pad, KH, KW, Cin_pad] goes through shared memory and comes out as the
// contiguous [co][ci][tap] run of the nn.Conv2d weight gradient.  Offsets are relative to the two base pointers, so the
// table is static across steps.
template <int TT>
__device__ __forceinline__ void wgrad_to_param_tiles(const fsnet_wgrad_desc& d, const float* __restrict__ acc_base, float* __restrict__ grad_base,
                                                     float* s_w) {
  const int T = TT ? TT : d.kh * d.kw;
  const int CI_T = d.cin_pad < 32 ? d.cin_pad : 32;
  int cs = 0;
  while ((1 << cs) < CI_T) ++cs;
  int CO_T = kWpTile / (CI_T * T);
  CO_T = CO_T > 16 ? 16 : (CO_T < 1 ? 1 : CO_T);
  if (TT) CO_T = 16;
  if (CI_T * T > kWpTile) return;
  const int ci_tiles = (d.cin + CI_T - 1) / CI_T, co_tiles = (d.cout + CO_T - 1) / CO_T;
  const int brick = CO_T * CI_T * T;
  const float* acc = acc_base + d.acc_off;
  float* grad = grad_base + d.grad_off;
  for (int tile = blockIdx.x; tile < ci_tiles * co_tiles; tile += gridDim.x) {
    const int co0 = (tile / ci_tiles) * CO_T, ci0 = (tile % ci_tiles) * CI_T;
    __syncthreads();
    for (int i = threadIdx.x; i < brick; i += 256) {         // accumulator layout: ci fastest
      const int ci_l = i & (CI_T - 1), t2 = i >> cs, co_l = t2 / T, tap = t2 - co_l * T;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      s_w[(co_l * CI_T + ci_l) * T + tap] = (co < d.cout && ci < d.cin) ? acc[((size_t)co * T + tap) * d.cin_pad + ci] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < brick; i += 256) {         // parameter layout: (ci, tap) contiguous
      const int q = i / T, co_l = q >> cs, ci_l = q & (CI_T - 1), rem = i - co_l * (CI_T * T);
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (co < d.cout && ci < d.cin) grad[((size_t)co * d.cin + ci0) * T + rem] = s_w[i];
    }
  }
}
__global__ void __launch_bounds__(256) wgrad_to_param_batched_kernel(const fsnet_wgrad_desc* __restrict__ table,
                                                                     const float* __restrict__ acc_base, float* __restrict__ grad_base) {
  __shared__ float s_w[kWpTile];
  const fsnet_wgrad_desc d = table[blockIdx.y];
  const int T = d.kh * d.kw;
  if (T == 9) wgrad_to_param_tiles<9>(d, acc_base, grad_base, s_w);
  else if (T == 1) wgrad_to_param_tiles<1>(d, acc_base, grad_base, s_w);
  else wgrad_to_param_tiles<0>(d, acc_base, grad_base, s_w);
}
__global__ void wgrad_to_param_kernel(const float* __restrict__ acc, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad,
                                      float* __restrict__ grad, int accumulate) {
  size_t total = (size_t)Cout * Cin * KH * KW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int s = (int)(i % KW); size_t t = i / KW;
  int r = (int)(t % KH); t /= KH;
  int ci = (int)(t % Cin); int co = (int)(t / Cin);
  float v = acc[(((size_t)co * KH + r) * KW + s) * Cin_pad + ci];
  grad[i] = accumulate ? grad[i] + v : v;
}

inline unsigned blocks_for(size_t total, int threads = 256) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace
}  // namespace fsnet

using namespace fsnet;

// blocks per layer of the two batched re-layout kernels (blockIdx.y = layer; blocks of small layers exit at once)
static unsigned relayout_grid_x() {
  static int g = 0;
  if (!g) { const char* e = getenv("FSNET_RELAYOUT_GRID"); g = e ? atoi(e) : 148; if (g < 1) g = 148; }
  return (unsigned)g;
}

extern "C" int fsnet_image_to_planes_ring(const float* img, int C, const fsnet_view* dst, int zero_ring, void* stream) {
  FSNET_REQUIRE(img && dst && dst->ptr && C <= dst->c, "fsnet_image_to_planes: bad arguments");
  FSNET_REQUIRE(dst->c % 8 == 0 && dst->c_off == 0 && dst->c == dst->c_total, "fsnet_image_to_planes: destination must be a whole buffer with channels % 8 == 0");
  size_t total = (size_t)dst->n * (dst->h + 2 * dst->ring) * (dst->w + 2 * dst->ring) * (dst->c / 8);
  FSNET_REQUIRE(total < (1ull << 32), "fsnet_image_to_planes: tensor too large for 32-bit indexing");
  image_to_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(img, C, *dst, zero_ring);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_image_to_planes(const float* img, int C, const fsnet_view* dst, void* stream) {
  return fsnet_image_to_planes_ring(img, C, dst, 0, stream);
}

extern "C" int fsnet_bn_finalize(double* stats, double count, const float* gamma, const float* beta, const float* conv_bias,
                                 float* running_mean, float* running_var, long long* num_batches, float momentum, float eps,
                                 int training, int C, float* scale_shift, float* mean_invstd, void* stream) {
  FSNET_REQUIRE(scale_shift && C > 0 && (training ? stats != nullptr : (running_mean && running_var)), "fsnet_bn_finalize: bad arguments");
  FSNET_LAUNCH_PDL(bn_finalize_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)stream, stats, count, gamma, beta, conv_bias, running_mean, running_var,
                                                                         num_batches, momentum, eps, training, C, scale_shift, mean_invstd);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_act_planes(const fsnet_view* raw, const float* scale_shift, int res_mode, const fsnet_view* res,
                                const float* res_scale_shift, int relu, int up, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(raw && dst && raw->ptr && dst->ptr && (up == 1 || up == 2), "fsnet_act_planes: bad arguments");
  FSNET_REQUIRE(dst->h == raw->h * up && dst->w == raw->w * up && dst->c == raw->c && dst->n == raw->n, "fsnet_act_planes: shape mismatch");
  FSNET_REQUIRE(res_mode == 0 || (res && res->ptr && (res_mode == 1 || res_scale_shift)), "fsnet_act_planes: residual arguments");
  ActParams p = {};
  p.raw = *raw; p.ss = scale_shift; p.res_mode = res_mode; if (res) p.res = *res; p.res_ss = res_scale_shift;
  p.relu = relu; p.up = up; p.dst = *dst;
  FSNET_REQUIRE(raw->c % 8 == 0 && raw->c_off % 8 == 0 && dst->c_off % 8 == 0 && dst->c_total % 8 == 0, "fsnet_act_planes: channels must be multiples of 8");
  size_t total = (size_t)raw->n * raw->h * raw->w * (raw->c / 8);
  FSNET_LAUNCH_PDL(act_planes_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_copy_planes(const fsnet_view* src, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(src && dst && src->ptr && dst->ptr && src->ring == dst->ring && src->h == dst->h && src->w == dst->w && src->c == dst->c,
                "fsnet_copy_planes: bad arguments");
  FSNET_REQUIRE(src->c % 8 == 0 && src->c_off % 8 == 0 && dst->c_off % 8 == 0 && dst->c_total % 8 == 0 && src->c_total % 8 == 0, "fsnet_copy_planes: channels must be multiples of 8");
  size_t total = (size_t)src->n * (src->h + 2 * src->ring) * (src->w + 2 * src->ring) * (src->c / 8);
  copy_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, *dst);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_maxpool_planes(const fsnet_view* src, const fsnet_view* dst, uint8_t* argmax, void* stream) {
  FSNET_REQUIRE(src && dst && src->ptr && dst->ptr && dst->h == (src->h + 1) / 2 && dst->w == (src->w + 1) / 2 && dst->c == src->c && src->c % 8 == 0,
                "fsnet_maxpool_planes: bad arguments");
  size_t total = (size_t)dst->n * dst->h * dst->w * (dst->c / 8);
  maxpool_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, *dst, argmax);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_maxpool_bwd(const fsnet_view* src, const uint8_t* argmax, const fsnet_view* grad_dst, const fsnet_view* grad_src,
                                 int accumulate, void* stream) {
  FSNET_REQUIRE(src && argmax && grad_dst && grad_src && grad_dst->ptr && grad_src->ptr, "fsnet_maxpool_bwd: bad arguments");
  size_t total = (size_t)src->n * src->h * src->w * (src->c / 8);
  maxpool_bwd_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, argmax, *grad_dst, *grad_src, accumulate);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

static int fill_bn_bwd(BnBwdParams& p, const fsnet_view* g, int up, const fsnet_view* mask, const float* mask_ss, const fsnet_view* raw,
                       const float* mean_invstd, const float* gamma, double* sums, double count) {
  FSNET_REQUIRE(g && raw && g->ptr && raw->ptr && sums && (up == 1 || up == 2), "fsnet_bn_bwd: bad arguments");
  FSNET_REQUIRE(g->h == raw->h * up && g->w == raw->w * up && g->c == raw->c, "fsnet_bn_bwd: gradient / raw shape mismatch");
  p.g = *g; p.up = up; p.has_mask = mask != nullptr && mask->ptr != nullptr; if (p.has_mask) p.mask = *mask;
  p.raw = *raw; p.mean_invstd = mean_invstd; p.gamma = gamma; p.sums = sums; p.count = count; p.mask_ss = mask_ss;
  return FSNET_OK;
}

extern "C" int fsnet_bn_bwd_reduce(const fsnet_view* g, int up, const fsnet_view* mask, const float* mask_ss, const fsnet_view* raw,
                                   const float* mean_invstd, double* sums, void* stream) {
  BnBwdParams p = {};
  int rc = fill_bn_bwd(p, g, up, mask, mask_ss, raw, mean_invstd, nullptr, sums, 1.0);
  if (rc) return rc;
  FSNET_REQUIRE(raw->c % 8 == 0 && raw->c <= 2048, "fsnet_bn_bwd_reduce: channels must be a multiple of 8 and <= 2048");
  const int groups = raw->c / 8;
  const int pix_per_iter = 256 / groups > 0 ? 256 / groups : 1;
  const int threads_used = groups * pix_per_iter;
  size_t npix = (size_t)raw->n * raw->h * raw->w;
  // at most 16 pixels per thread, but never fewer than ~2 blocks per SM: the small deep layers were latency bound
  // with a few dozen blocks walking their pixels serially
  size_t per_thread = npix / ((size_t)pix_per_iter * 296);
  per_thread = per_thread < 4 ? 4 : (per_thread > 16 ? 16 : per_thread);
  unsigned grid = (unsigned)((npix + (size_t)pix_per_iter * per_thread - 1) / ((size_t)pix_per_iter * per_thread));
  if (grid > 148 * 4) grid = 148 * 4;
  // every block ends with 2C fp64 atomics on the same 2C addresses: wide layers get fewer blocks
  const unsigned cap = (unsigned)(49152 / raw->c) > 24u ? (unsigned)(49152 / raw->c) : 24u;
  if (grid > cap) grid = cap;
  // whole waves: the kernel keeps two blocks per SM resident (126 registers), so 320 or 360 blocks ran as one full wave plus a
  // second, nearly empty one of the same length (ncu r2c8: 5 loop iterations per warp, the launch twice as long as a block)
  if (grid > 296u) grid = grid >= 592u ? 592u : 296u;
  if (grid == 0) grid = 1;
  FSNET_LAUNCH_PDL(bn_bwd_reduce_kernel, grid, 256, 2 * raw->c * sizeof(float), (cudaStream_t)stream, p, pix_per_iter, threads_used);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_bn_bwd_apply(const fsnet_view* g, int up, const fsnet_view* mask, const float* mask_ss, const fsnet_view* raw,
                                  const float* mean_invstd, const float* gamma, double* sums, double count,
                                  const fsnet_view* dy, int res_mode, const fsnet_view* res, float* dgamma, float* dbeta, int c_real,
                                  void* stream) {
  BnBwdParams p = {};
  int rc = fill_bn_bwd(p, g, up, mask, mask_ss, raw, mean_invstd, gamma, sums, count);
  if (rc) return rc;
  FSNET_REQUIRE(dy && dy->ptr && (res_mode == 0 || (res && res->ptr)), "fsnet_bn_bwd_apply: bad arguments");
  p.dy = *dy; p.res_mode = res_mode; if (res) p.res = *res;
  p.dgamma = dgamma; p.dbeta = dbeta; p.c_real = c_real;
  size_t total = (size_t)raw->n * raw->h * raw->w * (raw->c / 8);
  FSNET_REQUIRE(raw->c <= 2048 && total < (1ull << 32), "fsnet_bn_bwd_apply: channels <= 2048, 32-bit indexing");
  unsigned grid = blocks_for(total);
  if (grid > 148u * 8u) grid = 148u * 8u;          // persistent blocks: the per-channel coefficients are built once per block
  FSNET_LAUNCH_PDL(bn_bwd_apply_kernel, grid, 256, 3 * raw->c * sizeof(float), (cudaStream_t)stream, p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_fold_ring(const fsnet_view* g, void* stream) {
  FSNET_REQUIRE(g && g->ptr && g->ring == 1, "fsnet_fold_ring: needs a ringed fp32 view");
  const int per = 2 * g->w + 2 * (g->h - 2 > 0 ? g->h - 2 : 0);
  size_t total = (size_t)g->n * per * g->c;
  fold_ring_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*g);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_add_slice(const fsnet_view* dst, const fsnet_view* src, int accumulate, void* stream) {
  FSNET_REQUIRE(dst && src && dst->ptr && src->ptr && dst->h == src->h && dst->w == src->w && dst->c == src->c, "fsnet_add_slice: bad arguments");
  size_t total = (size_t)dst->n * dst->h * dst->w * (dst->c / 8);
  add_slice_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*dst, *src, accumulate);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_zero_insert(const fsnet_view* src, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(dst && src && dst->ptr && src->ptr && dst->c == src->c, "fsnet_zero_insert: bad arguments");
  size_t total = (size_t)dst->n * dst->h * dst->w * (dst->c / 8);
  zero_insert_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, *dst);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_weight_planes(const float* w, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad,
                                   void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* stream) {
  FSNET_REQUIRE(w && fwd_hi && Cout_pad >= Cout && Cin_pad >= Cin, "fsnet_weight_planes: bad arguments");
  size_t total = (size_t)Cout_pad * KH * KW * Cin_pad;
  weight_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, KH, KW, Cout_pad, Cin_pad, (__nv_bfloat16*)fwd_hi,
                                                                          (__nv_bfloat16*)fwd_lo, (__nv_bfloat16*)dgrad_hi);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_weight_planes_batched(const fsnet_weight_desc* table_device, int n_layers, void* stream) {
  FSNET_REQUIRE(table_device && n_layers > 0, "fsnet_weight_planes_batched: bad arguments");
  dim3 grid(relayout_grid_x(), n_layers);
  weight_planes_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table_device);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_wgrad_to_param_batched(const fsnet_wgrad_desc* table_device, int n_layers, const float* acc_base, float* grad_base,
                                            void* stream) {
  FSNET_REQUIRE(table_device && n_layers > 0 && acc_base && grad_base, "fsnet_wgrad_to_param_batched: bad arguments");
  dim3 grid(relayout_grid_x(), n_layers);
  wgrad_to_param_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table_device, acc_base, grad_base);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_wgrad_to_param(const float* acc, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad, float* grad,
                                    int accumulate, void* stream) {
  FSNET_REQUIRE(acc && grad, "fsnet_wgrad_to_param: bad arguments");
  size_t total = (size_t)Cout * Cin * KH * KW;
  wgrad_to_param_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(acc, Cout, Cin, KH, KW, Cout_pad, Cin_pad, grad, accumulate);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
