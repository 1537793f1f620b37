// Element-wise / reduction kernels around the tcgen05 convolutions: image -> bf16 planes, BatchNorm
// finalisation, BN-apply + residual + ReLU + hi/lo split (+ nearest x2 up-sampling into a concat
// buffer), max-pool, BatchNorm backward (reduce + apply), ring folding, weight re-layout.
//
// "planes": an activation x is stored as two NHWC bf16 tensors hi = bf16(x), lo = bf16(x - hi) inside a
// buffer with a one-pixel ring (replicate padding of the interior), so that tcgen05 convolutions read
// it directly with TMA.  All views are described by fsnet_view (include/fsnet_b200.h).
#include <cuda_bf16.h>
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

// ---- image (fp32 NCHW) -> planes with zero-filled extra channels ----------------------------------
__global__ void image_to_planes_kernel(const float* __restrict__ img, int C, V d) {
  size_t total = (size_t)d.n * d.h * d.w * d.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % d.c); size_t r = i / d.c;
  int x = (int)(r % d.w); r /= d.w;
  int y = (int)(r % d.h); int n = (int)(r / d.h);
  float v = c < C ? __ldg(img + (((size_t)n * C + c) * d.h + y) * d.w + x) : 0.f;
  __nv_bfloat16* hi = (__nv_bfloat16*)d.ptr;
  store_with_ring(d, hi, hi + plane_stride(d), n, y, x, c, v);
}

// ---- BatchNorm finalisation -------------------------------------------------------------------------
// training: batch statistics from the conv epilogue's fp64 sums; updates running stats (momentum, unbiased
// variance) exactly like nn.BatchNorm2d; the conv bias (if any) shifts the mean only.
__global__ void bn_finalize_kernel(double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches, float momentum, float eps, int training, int C,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
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
__global__ void act_planes_kernel(ActParams p) {
  const V& s = p.raw;
  size_t total = (size_t)s.n * s.h * s.w * s.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % s.c); size_t r = i / s.c;
  int x = (int)(r % s.w); r /= s.w;
  int y = (int)(r % s.h); int n = (int)(r / s.h);
  float v = ((const float*)s.ptr)[vidx(s, n, y, x, c)];
  if (p.ss) v = fmaf(v, __ldg(p.ss + c), __ldg(p.ss + s.c + c));
  if (p.res_mode == 1) {
    const __nv_bfloat16* rh = (const __nv_bfloat16*)p.res.ptr;
    v += plane_load(rh, rh + plane_stride(p.res), vidx(p.res, n, y, x, c));
  } else if (p.res_mode == 2) {
    float rv = ((const float*)p.res.ptr)[vidx(p.res, n, y, x, c)];
    v += fmaf(rv, __ldg(p.res_ss + c), __ldg(p.res_ss + s.c + c));
  }
  if (p.relu) v = fmaxf(v, 0.f);
  __nv_bfloat16* hi = (__nv_bfloat16*)p.dst.ptr;
  __nv_bfloat16* lo = hi + plane_stride(p.dst);
  if (p.up == 1) {
    store_with_ring(p.dst, hi, lo, n, y, x, c, v);
  } else {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) store_with_ring(p.dst, hi, lo, n, 2 * y + dy, 2 * x + dx, c, v);
  }
}

// ---- planes -> planes channel-slice copy (skip connection into the concat buffer), ring included --------
__global__ void copy_planes_kernel(V s, V d) {
  const int ph = s.h + 2 * s.ring, pw = s.w + 2 * s.ring;
  size_t total = (size_t)s.n * ph * pw * s.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % s.c); size_t r = i / s.c;
  int x = (int)(r % pw) - s.ring; r /= pw;
  int y = (int)(r % ph) - s.ring; int n = (int)(r / ph);
  const __nv_bfloat16* sh = (const __nv_bfloat16*)s.ptr;
  __nv_bfloat16* dh = (__nv_bfloat16*)d.ptr;
  size_t si = vidx(s, n, y, x, c), di = vidx(d, n, y, x, c);
  dh[di] = sh[si];
  dh[di + plane_stride(d)] = sh[si + plane_stride(s)];
}

// ---- 3x3 / stride 2 / pad 1 max-pool on planes ---------------------------------------------------------------
__global__ void maxpool_planes_kernel(V s, V d) {
  size_t total = (size_t)d.n * d.h * d.w * d.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % d.c); size_t r = i / d.c;
  int x = (int)(r % d.w); r /= d.w;
  int y = (int)(r % d.h); int n = (int)(r / d.h);
  const __nv_bfloat16* sh = (const __nv_bfloat16*)s.ptr;
  const __nv_bfloat16* sl = sh + plane_stride(s);
  float m = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      int yy = 2 * y + dy, xx = 2 * x + dx;
      if (yy < 0 || yy >= s.h || xx < 0 || xx >= s.w) continue;
      m = fmaxf(m, plane_load(sh, sl, vidx(s, n, yy, xx, c)));
    }
  __nv_bfloat16* dh = (__nv_bfloat16*)d.ptr;
  store_with_ring(d, dh, dh + plane_stride(d), n, y, x, c, m);
}
// backward: each source pixel gathers from the (up to 4) windows that contain it and whose first maximum
// (row-major scan order, as PyTorch) it is.  g_src (+)= ...
__global__ void maxpool_bwd_kernel(V s, V gd, V gs, int accumulate) {
  size_t total = (size_t)s.n * s.h * s.w * s.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % s.c); size_t r = i / s.c;
  int x = (int)(r % s.w); r /= s.w;
  int y = (int)(r % s.h); int n = (int)(r / s.h);
  const __nv_bfloat16* sh = (const __nv_bfloat16*)s.ptr;
  const __nv_bfloat16* sl = sh + plane_stride(s);
  const float me = plane_load(sh, sl, vidx(s, n, y, x, c));
  float g = 0.f;
  for (int oy = (y - 1 + 1) / 2; oy <= (y + 1) / 2; ++oy) {          // windows with 2*oy-1 <= y <= 2*oy+1
    if (oy < 0 || oy >= gd.h) continue;
    for (int ox = x / 2; ox <= (x + 1) / 2; ++ox) {
      if (ox < 0 || ox >= gd.w) continue;
      bool first = true;                                               // is (y,x) the first maximum of window (oy,ox)?
      for (int dy = -1; dy <= 1 && first; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          int yy = 2 * oy + dy, xx = 2 * ox + dx;
          if (yy < 0 || yy >= s.h || xx < 0 || xx >= s.w) continue;
          float v = plane_load(sh, sl, vidx(s, n, yy, xx, c));
          bool before = (yy < y) || (yy == y && xx < x);
          if (v > me || (before && v == me)) { first = false; break; }
        }
      if (first) g += ((const float*)gd.ptr)[vidx(gd, n, oy, ox, c)];
    }
  }
  float* o = (float*)gs.ptr + vidx(gs, n, y, x, c);
  *o = accumulate ? *o + g : g;
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
};
__device__ __forceinline__ float read_g(const BnBwdParams& p, int n, int y, int x, int c) {
  const float* g = (const float*)p.g.ptr;
  if (p.up == 1) return g[vidx(p.g, n, y, x, c)];
  return g[vidx(p.g, n, 2 * y, 2 * x, c)] + g[vidx(p.g, n, 2 * y, 2 * x + 1, c)] + g[vidx(p.g, n, 2 * y + 1, 2 * x, c)] +
         g[vidx(p.g, n, 2 * y + 1, 2 * x + 1, c)];
}
__device__ __forceinline__ float masked_g(const BnBwdParams& p, int n, int y, int x, int c) {
  float g = read_g(p, n, y, x, c);
  if (p.has_mask) {
    const __nv_bfloat16* mh = (const __nv_bfloat16*)p.mask.ptr;
    size_t mi = vidx(p.mask, n, y, x, c);
    float a = __bfloat162float(mh[mi]) + __bfloat162float(mh[mi + plane_stride(p.mask)]);
    g = a > 0.f ? g : 0.f;
  } else if (p.mask_ss) {
    float a = fmaf(((const float*)p.raw.ptr)[vidx(p.raw, n, y, x, c)], __ldg(p.mask_ss + c), __ldg(p.mask_ss + p.raw.c + c));
    g = a > 0.f ? g : 0.f;
  }
  return g;
}
// grid: (pixel blocks); block 256 threads = 8 pixel-rows x 32 channel lanes; channels looped in chunks of 32
__global__ void bn_bwd_reduce_kernel(BnBwdParams p) {
  const V& r = p.raw;
  const int C = r.c;
  const size_t npix = (size_t)r.n * r.h * r.w;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  __shared__ float red[2][8][32];
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float s1 = 0.f, s2 = 0.f;
    if (c < C) {
      const float mean = p.mean_invstd ? __ldg(p.mean_invstd + c) : 0.f, inv = p.mean_invstd ? __ldg(p.mean_invstd + C + c) : 0.f;
      for (size_t pix = (size_t)blockIdx.x * 8 + row; pix < npix; pix += (size_t)gridDim.x * 8) {
        int x = (int)(pix % r.w); size_t t = pix / r.w;
        int y = (int)(t % r.h); int n = (int)(t / r.h);
        float g = masked_g(p, n, y, x, c);
        float xh = (((const float*)r.ptr)[vidx(r, n, y, x, c)] - mean) * inv;
        s1 += g; s2 = fmaf(g, xh, s2);
      }
    }
    red[0][row][lane] = s1; red[1][row][lane] = s2;
    __syncthreads();
    if (row < 2 && c < C) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += red[row][k][lane];
      atomicAdd(p.sums + row * C + c, (double)t);
    }
    __syncthreads();
  }
}
__global__ void bn_bwd_apply_kernel(BnBwdParams p) {
  const V& r = p.raw;
  const int C = r.c;
  size_t total = (size_t)r.n * r.h * r.w * C;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C); size_t t = i / C;
  int x = (int)(t % r.w); t /= r.w;
  int y = (int)(t % r.h); int n = (int)(t / r.h);
  float g = masked_g(p, n, y, x, c);
  if (p.res_mode) {
    float* o = (float*)p.res.ptr + vidx(p.res, n, y, x, c);
    *o = p.res_mode == 2 ? *o + g : g;
  }
  float d = g;
  if (p.mean_invstd) {
    const float mean = __ldg(p.mean_invstd + c), inv = __ldg(p.mean_invstd + C + c);
    const float xh = (((const float*)r.ptr)[vidx(r, n, y, x, c)] - mean) * inv;
    const float sg = (float)(p.sums[c] / p.count), sgx = (float)(p.sums[C + c] / p.count);
    d = (p.gamma ? __ldg(p.gamma + c) : 1.f) * inv * (g - sg - xh * sgx);
  }
  __nv_bfloat16* dh = (__nv_bfloat16*)p.dy.ptr;
  dh[vidx(p.dy, n, y, x, c)] = __float2bfloat16_rn(d);
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
__global__ void add_slice_kernel(V d, V s, int accumulate) {
  size_t total = (size_t)d.n * d.h * d.w * d.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % d.c); size_t t = i / d.c;
  int x = (int)(t % d.w); t /= d.w;
  int y = (int)(t % d.h); int n = (int)(t / d.h);
  float v = ((const float*)s.ptr)[vidx(s, n, y, x, c)];
  float* o = (float*)d.ptr + vidx(d, n, y, x, c);
  *o = accumulate ? *o + v : v;
}

// zero-insertion x2 of a bf16 plane (transposed stride-2 convolution as a stride-1 convolution)
__global__ void zero_insert_kernel(V s, V d) {
  size_t total = (size_t)d.n * d.h * d.w * d.c;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % d.c); size_t t = i / d.c;
  int x = (int)(t % d.w); t /= d.w;
  int y = (int)(t % d.h); int n = (int)(t / d.h);
  __nv_bfloat16 v = __float2bfloat16_rn(0.f);
  if (!(y & 1) && !(x & 1) && (y >> 1) < s.h && (x >> 1) < s.w) v = ((const __nv_bfloat16*)s.ptr)[vidx(s, n, y >> 1, x >> 1, c)];
  ((__nv_bfloat16*)d.ptr)[vidx(d, n, y, x, c)] = v;
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

// wgrad accumulator fp32 [Cout_pad,KH,KW,Cin_pad] -> parameter gradient [Cout,Cin,KH,KW] (+=)
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

extern "C" int fsnet_image_to_planes(const float* img, int C, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(img && dst && dst->ptr && C <= dst->c, "fsnet_image_to_planes: bad arguments");
  size_t total = (size_t)dst->n * dst->h * dst->w * dst->c;
  image_to_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(img, C, *dst);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_bn_finalize(double* stats, double count, const float* gamma, const float* beta, const float* conv_bias,
                                 float* running_mean, float* running_var, long long* num_batches, float momentum, float eps,
                                 int training, int C, float* scale_shift, float* mean_invstd, void* stream) {
  FSNET_REQUIRE(scale_shift && C > 0 && (training ? stats != nullptr : (running_mean && running_var)), "fsnet_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, conv_bias, running_mean, running_var,
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
  size_t total = (size_t)raw->n * raw->h * raw->w * raw->c;
  act_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_copy_planes(const fsnet_view* src, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(src && dst && src->ptr && dst->ptr && src->ring == dst->ring && src->h == dst->h && src->w == dst->w && src->c == dst->c,
                "fsnet_copy_planes: bad arguments");
  size_t total = (size_t)src->n * (src->h + 2 * src->ring) * (src->w + 2 * src->ring) * src->c;
  copy_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, *dst);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_maxpool_planes(const fsnet_view* src, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(src && dst && src->ptr && dst->ptr && dst->h == (src->h + 1) / 2 && dst->w == (src->w + 1) / 2 && dst->c == src->c,
                "fsnet_maxpool_planes: bad arguments");
  size_t total = (size_t)dst->n * dst->h * dst->w * dst->c;
  maxpool_planes_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, *dst);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_maxpool_bwd(const fsnet_view* src, const fsnet_view* grad_dst, const fsnet_view* grad_src, int accumulate, void* stream) {
  FSNET_REQUIRE(src && grad_dst && grad_src && src->ptr && grad_dst->ptr && grad_src->ptr, "fsnet_maxpool_bwd: bad arguments");
  size_t total = (size_t)src->n * src->h * src->w * src->c;
  maxpool_bwd_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*src, *grad_dst, *grad_src, accumulate);
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
  size_t npix = (size_t)raw->n * raw->h * raw->w;
  unsigned grid = (unsigned)((npix + 8 * 16 - 1) / (8 * 16));
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid == 0) grid = 1;
  bn_bwd_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_bn_bwd_apply(const fsnet_view* g, int up, const fsnet_view* mask, const float* mask_ss, const fsnet_view* raw,
                                  const float* mean_invstd, const float* gamma, double* sums, double count,
                                  const fsnet_view* dy, int res_mode, const fsnet_view* res, void* stream) {
  BnBwdParams p = {};
  int rc = fill_bn_bwd(p, g, up, mask, mask_ss, raw, mean_invstd, gamma, sums, count);
  if (rc) return rc;
  FSNET_REQUIRE(dy && dy->ptr && (res_mode == 0 || (res && res->ptr)), "fsnet_bn_bwd_apply: bad arguments");
  p.dy = *dy; p.res_mode = res_mode; if (res) p.res = *res;
  size_t total = (size_t)raw->n * raw->h * raw->w * raw->c;
  bn_bwd_apply_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(p);
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
  size_t total = (size_t)dst->n * dst->h * dst->w * dst->c;
  add_slice_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(*dst, *src, accumulate);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_zero_insert(const fsnet_view* src, const fsnet_view* dst, void* stream) {
  FSNET_REQUIRE(dst && src && dst->ptr && src->ptr && dst->c == src->c, "fsnet_zero_insert: bad arguments");
  size_t total = (size_t)dst->n * dst->h * dst->w * dst->c;
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

extern "C" int fsnet_wgrad_to_param(const float* acc, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad, float* grad,
                                    int accumulate, void* stream) {
  FSNET_REQUIRE(acc && grad, "fsnet_wgrad_to_param: bad arguments");
  size_t total = (size_t)Cout * Cin * KH * KW;
  wgrad_to_param_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(acc, Cout, Cin, KH, KW, Cout_pad, Cin_pad, grad, accumulate);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
