// Edge-aware disparity smoothness, the softmax-over-bins / sigmoid depth head and the loss finaliser.
// Reference: monodepth_utils.py:168-181 + monodepth2_decoder.py:214-219,294-303 (smoothness),
// depth_encoder.py:76-88,104-109,115-121 + monodepth_utils.py:8-24 (heads).
#include "common.cuh"

namespace fsnet {
namespace {

constexpr int TX = 32, TY = 8;     // low-resolution tile of one block

// Loads the (TY+2) x (TX+2) halo tile of disparity and of the k x k box-averaged colour image
// (== adaptive_avg_pool2d for H = h*k) into shared memory; out-of-range positions are clamped and
// never used by the callers.
__device__ __forceinline__ void load_tile(const float* __restrict__ disp, const float* __restrict__ img,
                                          int b, int h, int w, int H, int W, int k, int ty0, int tx0,
                                          float (&sd)[TY + 2][TX + 2], float (&sc)[3][TY + 2][TX + 2]) {
  const float inv = 1.f / (float)(k * k);
  for (int i = threadIdx.x; i < (TY + 2) * (TX + 2); i += blockDim.x) {
    int ly = i / (TX + 2), lx = i % (TX + 2);
    int y = min(max(ty0 + ly - 1, 0), h - 1), x = min(max(tx0 + lx - 1, 0), w - 1);
    sd[ly][lx] = __ldg(disp + ((size_t)b * h + y) * w + x);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* base = img + (((size_t)b * 3 + c) * H + (size_t)y * k) * W + (size_t)x * k;
      float s = 0.f;
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) s += __ldg(base + (size_t)dy * W + dx);
      sc[c][ly][lx] = k == 1 ? s : s * inv;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float edge_weight(const float (&sc)[3][TY + 2][TX + 2], int y0, int x0, int y1, int x1) {
  float g = fabsf(sc[0][y0][x0] - sc[0][y1][x1]) + fabsf(sc[1][y0][x0] - sc[1][y1][x1]) + fabsf(sc[2][y0][x0] - sc[2][y1][x1]);
  return __expf(-g * (1.f / 3.f));
}

// sums[b] = { sum(disp), sum_x-edges w|d_i-d_j|, sum_y-edges w|d_i-d_j| }  (un-normalised disparity)
__global__ void __launch_bounds__(TX * TY) smooth_fwd_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                                             int h, int w, int H, int W, int k, double* __restrict__ sums) {
  __shared__ float sd[TY + 2][TX + 2];
  __shared__ float sc[3][TY + 2][TX + 2];
  __shared__ float red[3][TX * TY / 32];
  const int b = blockIdx.z, ty0 = blockIdx.y * TY, tx0 = blockIdx.x * TX;
  load_tile(disp, img, b, h, w, H, W, k, ty0, tx0, sd, sc);
  const int lx = threadIdx.x % TX + 1, ly = threadIdx.x / TX + 1;
  const int x = tx0 + lx - 1, y = ty0 + ly - 1;
  float s0 = 0.f, sx = 0.f, sy = 0.f;
  if (x < w && y < h) {
    float d = sd[ly][lx];
    s0 = d;
    if (x + 1 < w) sx = fabsf(d - sd[ly][lx + 1]) * edge_weight(sc, ly, lx, ly, lx + 1);
    if (y + 1 < h) sy = fabsf(d - sd[ly + 1][lx]) * edge_weight(sc, ly, lx, ly + 1, lx);
  }
  s0 = warp_sum(s0); sx = warp_sum(sx); sy = warp_sum(sy);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = sx; red[2][warp] = sy; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < TX * TY / 32; ++i) t += (double)red[threadIdx.x][i];
    atomicAdd(sums + (size_t)b * 3 + threadIdx.x, t);
  }
}

__global__ void smooth_finalize_kernel(const double* __restrict__ sums, int B, int h, int w, float weight, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nx = (double)B * h * (w - 1), ny = (double)B * (h - 1) * w;
  double acc = 0.0;
  for (int b = 0; b < B; ++b) {
    float mean = (float)(sums[b * 3] / ((double)h * w));
    double m = (double)(mean + 1e-7f);
    acc += (sums[b * 3 + 1] / nx + sums[b * 3 + 2] / ny) / m;
  }
  *out += acc * (double)weight;
}

__global__ void __launch_bounds__(TX * TY) smooth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                                             int B, int h, int w, int H, int W, int k, float weight,
                                                             const double* __restrict__ sums, const float* __restrict__ gout,
                                                             float* __restrict__ grad) {
  __shared__ float sd[TY + 2][TX + 2];
  __shared__ float sc[3][TY + 2][TX + 2];
  const int b = blockIdx.z, ty0 = blockIdx.y * TY, tx0 = blockIdx.x * TX;
  load_tile(disp, img, b, h, w, H, W, k, ty0, tx0, sd, sc);
  const int lx = threadIdx.x % TX + 1, ly = threadIdx.x / TX + 1;
  const int x = tx0 + lx - 1, y = ty0 + ly - 1;
  if (x >= w || y >= h) return;
  const double nx = (double)B * h * (w - 1), ny = (double)B * (h - 1) * w;
  const float g = __ldg(gout) * weight;
  const float cx = (float)((double)g / nx), cy = (float)((double)g / ny);
  const float mean = (float)(sums[b * 3] / ((double)h * w));
  const float m = mean + 1e-7f;
  const float d = sd[ly][lx];
  auto sgn = [](float v) { return (v > 0.f ? 1.f : 0.f) - (v < 0.f ? 1.f : 0.f); };
  float gn = 0.f;      // d L / d normalised disparity at this pixel (times m)
  if (x + 1 < w) gn += cx * sgn(d - sd[ly][lx + 1]) * edge_weight(sc, ly, lx, ly, lx + 1);
  if (x > 0) gn -= cx * sgn(sd[ly][lx - 1] - d) * edge_weight(sc, ly, lx - 1, ly, lx);
  if (y + 1 < h) gn += cy * sgn(d - sd[ly + 1][lx]) * edge_weight(sc, ly, lx, ly + 1, lx);
  if (y > 0) gn -= cy * sgn(sd[ly - 1][lx] - d) * edge_weight(sc, ly - 1, lx, ly, lx);
  // sum_i gnd_i d_i = g * (Sx/nx + Sy/ny) / m  -- each edge contributes coefficient * |d_i - d_j|
  const double dot = (double)g * (sums[b * 3 + 1] / nx + sums[b * 3 + 2] / ny) / (double)m;
  grad[((size_t)b * h + y) * w + x] = gn / m - (float)(dot / ((double)m * (double)h * w));
}

__global__ void depth_head_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ bins,
                                      const float* __restrict__ scale, int B, int n, int hw, int channels_last,
                                      int sigmoid_head, float min_depth, float max_depth,
                                      float* __restrict__ depth, float* __restrict__ disp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * hw) return;
  int b = (int)(i / hw);
  size_t pix = i % hw;
  const float sc = scale ? __ldg(scale + b) : 1.f;
  const size_t base = channels_last ? i * n : (size_t)b * n * hw + pix;
  const size_t cs = channels_last ? 1 : hw;
  if (sigmoid_head) {
    float l = __ldg(logits + base);
    float s = 1.f / (1.f + __expf(-l));
    disp[i] = s;
    depth[i] = sc / (1.f / max_depth + (1.f / min_depth - 1.f / max_depth) * s);
    return;
  }
  float se = 0.f, sb = 0.f;
  for (int c = 0; c < n; ++c) {
    float l = fminf(fmaxf(__ldg(logits + base + c * cs), -10.f), 10.f);
    float e = __expf(l);
    se += e;
    sb = fmaf(e, __ldg(bins + c), sb);
  }
  float d = sb / se * sc;
  const float mn = min_depth * sc, mx = max_depth * sc;
  depth[i] = d;
  disp[i] = (1.f / d - 1.f / mx) / (1.f / mn - 1.f / mx);
}

__global__ void depth_head_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ bins,
                                      const float* __restrict__ scale, int B, int n, int hw, int channels_last,
                                      int sigmoid_head, float min_depth, float max_depth,
                                      const float* __restrict__ gdepth, const float* __restrict__ gdisp,
                                      float* __restrict__ glogits) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * hw) return;
  int b = (int)(i / hw);
  size_t pix = i % hw;
  const float sc = scale ? __ldg(scale + b) : 1.f;
  const size_t base = channels_last ? i * n : (size_t)b * n * hw + pix;
  const size_t cs = channels_last ? 1 : hw;
  const float gd = gdepth ? __ldg(gdepth + i) : 0.f, gs = gdisp ? __ldg(gdisp + i) : 0.f;
  if (sigmoid_head) {
    float l = __ldg(logits + base);
    float s = 1.f / (1.f + __expf(-l));
    float span = 1.f / min_depth - 1.f / max_depth;
    float raw = 1.f / (1.f / max_depth + span * s);
    float g = gs + gd * (-raw * raw * span * sc);
    glogits[base] = g * s * (1.f - s);
    return;
  }
  float se = 0.f, sb = 0.f;
  for (int c = 0; c < n; ++c) {
    float l = fminf(fmaxf(__ldg(logits + base + c * cs), -10.f), 10.f);
    float e = __expf(l);
    se += e;
    sb = fmaf(e, __ldg(bins + c), sb);
  }
  const float raw = sb / se;
  const float d = raw * sc;
  const float mn = min_depth * sc, mx = max_depth * sc;
  const float g = (gd + gs * (-1.f / (d * d)) / (1.f / mn - 1.f / mx)) * sc;     // d L / d raw
  const float rse = 1.f / se;
  for (int c = 0; c < n; ++c) {
    float lr = __ldg(logits + base + c * cs);
    float l = fminf(fmaxf(lr, -10.f), 10.f);
    float pc = __expf(l) * rse;
    float live = (lr >= -10.f && lr <= 10.f) ? 1.f : 0.f;
    glogits[base + c * cs] = g * pc * (__ldg(bins + c) - raw) * live;
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, int S, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double total = 0.0;
  for (int s = 0; s < S; ++s) {
    double ls = acc[4 * s] / (acc[4 * s + 1] + 1e-6) + acc[4 * s + 2];
    out[s] = ls;
    out[S + s] = acc[4 * s + 2];
    total += ls;
  }
  out[2 * S] = total / (double)S;
  out[2 * S + 1] = 0.0;
}

int smooth_check(const float* disp, const float* img, int B, int h, int w, int H, int W, const double* sums) {
  FSNET_REQUIRE(disp && img && sums, "fsnet_smooth: null pointer");
  FSNET_REQUIRE(B > 0 && h >= 2 && w >= 2 && H >= h && W >= w, "fsnet_smooth: bad shape");
  FSNET_REQUIRE(H % h == 0 && W % w == 0 && H / h == W / w, "fsnet_smooth: image must be an integer multiple of the disparity (H=%d h=%d W=%d w=%d)", H, h, W, w);
  return FSNET_OK;
}


// k x k box average of an NCHW image (== F.adaptive_avg_pool2d for H = h*k, monodepth2_decoder.py:219): the colour image of the
// smoothness term at scale s, computed ONCE per step and scale instead of inside both smoothness kernels.
__global__ void __launch_bounds__(256) box_pool_kernel(const float* __restrict__ img, int planes, int H, int W, int k, float* __restrict__ out) {
  const int h = H / k, w = W / k;
  const unsigned total = (unsigned)planes * h * w;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned x = i % w, t = i / w, y = t % h, pl = t / h;
  const float* base = img + ((size_t)pl * H + (size_t)y * k) * W + (size_t)x * k;
  float s = 0.f;
  for (int dy = 0; dy < k; ++dy)
    for (int dx = 0; dx < k; ++dx) s += __ldg(base + (size_t)dy * W + dx);
  out[i] = s * (1.f / (float)(k * k));
}

// ---- axis-angle + translation -> 4x4 (PoseNet output to cam_T_cam) --------------------------------------------------------------
// rot_from_axisangle / get_translation_matrix / transformation_from_parameters (monodepth_utils.py:298-337, 31-44, 46-63): Rodrigues with
// the reference's axis = v / (|v| + 1e-7); M = T * R, or R^T * T(-t) when inverted.  One thread per sample; the backward contracts the
// incoming d loss / d M with forward-mode derivatives of the same formula (six dual-number evaluations) -- ~40 eager PyTorch launches per
// direction in the reference, one tiny launch here.
struct Dual { float v, d; };
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) { const float q = a.v / b.v; return {q, (a.d - q * b.d) / b.v}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual dsqrt(Dual a) { const float r = sqrtf(a.v); return {r, r > 0.f ? 0.5f * a.d / r : 0.f}; }
__device__ __forceinline__ Dual dsin(Dual a) { return {sinf(a.v), cosf(a.v) * a.d}; }
__device__ __forceinline__ Dual dcos(Dual a) { return {cosf(a.v), -sinf(a.v) * a.d}; }
__device__ __forceinline__ Dual cst(float c) { return {c, 0.f}; }

// M[3][4] (the last row is 0 0 0 1) from the six parameters; `seed` = index of the parameter whose derivative rides along (-1: none)
__device__ __forceinline__ void pose_matrix_dual(const float* aa, const float* tr, int invert, int seed, Dual (&M)[12]) {
  Dual v[3], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = {aa[i], seed == i ? 1.f : 0.f}; t[i] = {tr[i], seed == 3 + i ? 1.f : 0.f}; }
  const Dual angle = dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const Dual den = angle + cst(1e-7f);
  const Dual x = v[0] / den, y = v[1] / den, z = v[2] / den;
  const Dual ca = dcos(angle), sa = dsin(angle), C = cst(1.f) - ca;
  Dual R[9] = {x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa,
               x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa,
               z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca};
  if (!invert) {                              // T * R = [R | t]
#pragma unroll
    for (int i = 0; i < 3; ++i) { M[4 * i] = R[3 * i]; M[4 * i + 1] = R[3 * i + 1]; M[4 * i + 2] = R[3 * i + 2]; M[4 * i + 3] = t[i]; }
  } else {                                    // R^T * T(-t) = [R^T | -R^T t]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      M[4 * i] = R[i]; M[4 * i + 1] = R[3 + i]; M[4 * i + 2] = R[6 + i];
      M[4 * i + 3] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
    }
  }
}
__global__ void pose_matrix_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int B, int invert, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  Dual M[12];
  pose_matrix_dual(aa + 3 * b, tr + 3 * b, invert, -1, M);
  float* o = out + 16 * b;
#pragma unroll
  for (int i = 0; i < 12; ++i) o[i] = M[i].v;
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}
__global__ void pose_matrix_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, const float* __restrict__ gM, int B,
                                       int invert, float* __restrict__ g_aa, float* __restrict__ g_tr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 6) return;
  const int b = i / 6, k = i - 6 * b;
  Dual M[12];
  pose_matrix_dual(aa + 3 * b, tr + 3 * b, invert, k, M);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 12; ++j) s = fmaf(gM[16 * b + j], M[j].d, s);
  if (k < 3) g_aa[3 * b + k] = s; else g_tr[3 * b + k - 3] = s;
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

extern "C" int fsnet_smooth_fwd(const float* disp, const float* img, int B, int h, int w, int H, int W,
                                float weight, double* sums, double* out, void* stream) {
  int rc = smooth_check(disp, img, B, h, w, H, W, sums);
  if (rc) return rc;
  FSNET_REQUIRE(out, "fsnet_smooth_fwd: null output");
  dim3 grid(ceil_div(w, TX), ceil_div(h, TY), B);
  smooth_fwd_kernel<<<grid, TX * TY, 0, (cudaStream_t)stream>>>(disp, img, h, w, H, W, H / h, sums);
  FSNET_LAUNCH_OK();
  smooth_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, B, h, w, weight, out);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_smooth_bwd(const float* disp, const float* img, int B, int h, int w, int H, int W,
                                float weight, double* sums, const float* gout, float* grad_disp, void* stream) {
  int rc = smooth_check(disp, img, B, h, w, H, W, sums);
  if (rc) return rc;
  FSNET_REQUIRE(gout && grad_disp, "fsnet_smooth_bwd: null pointer");
  dim3 grid(ceil_div(w, TX), ceil_div(h, TY), B);
  smooth_bwd_kernel<<<grid, TX * TY, 0, (cudaStream_t)stream>>>(disp, img, B, h, w, H, W, H / h, weight, sums, gout, grad_disp);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

// ---- channels-last softmax head, n in {4, 8, 16, 32, 64, 128}: L = n/4 lanes per pixel, one float4 of logits per lane, the two
// sums over bins by xor-shuffles inside the lane group -> every global access is a fully coalesced 128-bit transaction
// (the thread-per-pixel kernel above strides 4-byte accesses by n*4 bytes across the warp).
template <int BWD>
__global__ void __launch_bounds__(256) depth_head_cl_kernel(const float* __restrict__ logits, const float* __restrict__ bins,
                                                            const float* __restrict__ scale, int B, int n, int hw, int L,
                                                            float min_depth, float max_depth, float* __restrict__ depth,
                                                            float* __restrict__ disp, const float* __restrict__ gdepth,
                                                            const float* __restrict__ gdisp, float* __restrict__ glogits) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i = t / L;                       // pixel
  const int sub = (int)(t % L);                 // float4 index inside the pixel's bin vector
  const bool live_px = i < (size_t)B * hw;
  const size_t ic = live_px ? i : 0;
  const float4 lr = __ldg(reinterpret_cast<const float4*>(logits + ic * n) + sub);
  const float4 bn = __ldg(reinterpret_cast<const float4*>(bins) + sub);
  const float l[4] = {lr.x, lr.y, lr.z, lr.w}, bv[4] = {bn.x, bn.y, bn.z, bn.w};
  float e[4], se = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e[k] = __expf(fminf(fmaxf(l[k], -10.f), 10.f));
    se += e[k];
    sb = fmaf(e[k], bv[k], sb);
  }
  for (int o = L >> 1; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  const float sc = scale ? __ldg(scale + ic / hw) : 1.f;
  const float raw = sb / se, d = raw * sc;
  const float mn = min_depth * sc, mx = max_depth * sc;
  if (!BWD) {
    if (live_px && sub == 0) {
      depth[i] = d;
      disp[i] = (1.f / d - 1.f / mx) / (1.f / mn - 1.f / mx);
    }
  } else {
    const float gd = gdepth ? __ldg(gdepth + ic) : 0.f, gs = gdisp ? __ldg(gdisp + ic) : 0.f;
    const float g = (gd + gs * (-1.f / (d * d)) / (1.f / mn - 1.f / mx)) * sc;     // d L / d raw
    const float rse = 1.f / se;
    float o4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o4[k] = (l[k] >= -10.f && l[k] <= 10.f) ? g * (e[k] * rse) * (bv[k] - raw) : 0.f;
    if (live_px) reinterpret_cast<float4*>(glogits + i * n)[sub] = make_float4(o4[0], o4[1], o4[2], o4[3]);
  }
}
static bool head_cl_ok(int n, int channels_last, int sigmoid_head, const void* logits, const void* bins, const void* out) {
  const int L = n / 4;
  return channels_last && !sigmoid_head && n % 4 == 0 && L >= 1 && L <= 32 && (L & (L - 1)) == 0 &&
         ((uintptr_t)logits & 15) == 0 && ((uintptr_t)bins & 15) == 0 && ((uintptr_t)out & 15) == 0;
}

extern "C" int fsnet_depth_head_fwd(const float* logits, const float* bins, const float* scale, int B, int n, int h, int w,
                                    int channels_last, int sigmoid_head, float min_depth, float max_depth,
                                    float* depth, float* disp, void* stream) {
  FSNET_REQUIRE(logits && depth && disp && (sigmoid_head || bins), "fsnet_depth_head_fwd: null pointer");
  FSNET_REQUIRE(B > 0 && n > 0 && h > 0 && w > 0 && (!sigmoid_head || n == 1), "fsnet_depth_head_fwd: bad shape");
  size_t total = (size_t)B * h * w;
  if (head_cl_ok(n, channels_last, sigmoid_head, logits, bins, nullptr)) {
    const int L = n / 4;
    depth_head_cl_kernel<0><<<(unsigned)((total * L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        logits, bins, scale, B, n, h * w, L, min_depth, max_depth, depth, disp, nullptr, nullptr, nullptr);
    FSNET_LAUNCH_OK();
    return FSNET_OK;
  }
  depth_head_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      logits, bins, scale, B, n, h * w, channels_last, sigmoid_head, min_depth, max_depth, depth, disp);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_depth_head_bwd(const float* logits, const float* bins, const float* scale, int B, int n, int h, int w,
                                    int channels_last, int sigmoid_head, float min_depth, float max_depth,
                                    const float* grad_depth, const float* grad_disp, float* grad_logits, void* stream) {
  FSNET_REQUIRE(logits && grad_logits && (sigmoid_head || bins), "fsnet_depth_head_bwd: null pointer");
  FSNET_REQUIRE(B > 0 && n > 0 && h > 0 && w > 0 && (!sigmoid_head || n == 1), "fsnet_depth_head_bwd: bad shape");
  size_t total = (size_t)B * h * w;
  if (head_cl_ok(n, channels_last, sigmoid_head, logits, bins, grad_logits)) {
    const int L = n / 4;
    depth_head_cl_kernel<1><<<(unsigned)((total * L + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        logits, bins, scale, B, n, h * w, L, min_depth, max_depth, nullptr, nullptr, grad_depth, grad_disp, grad_logits);
    FSNET_LAUNCH_OK();
    return FSNET_OK;
  }
  depth_head_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      logits, bins, scale, B, n, h * w, channels_last, sigmoid_head, min_depth, max_depth, grad_depth, grad_disp, grad_logits);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_loss_finalize(const double* acc, int S, double* out, void* stream) {
  FSNET_REQUIRE(acc && out && S > 0 && S <= 8, "fsnet_loss_finalize: bad arguments");
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, S, out);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_pose_matrix(const float* axisangle, const float* translation, int B, int invert, float* T, void* stream) {
  FSNET_REQUIRE(axisangle && translation && T && B > 0, "fsnet_pose_matrix: bad arguments");
  pose_matrix_kernel<<<ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(axisangle, translation, B, invert, T);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_pose_matrix_bwd(const float* axisangle, const float* translation, const float* grad_T, int B, int invert,
                                     float* grad_axisangle, float* grad_translation, void* stream) {
  FSNET_REQUIRE(axisangle && translation && grad_T && grad_axisangle && grad_translation && B > 0, "fsnet_pose_matrix_bwd: bad arguments");
  pose_matrix_bwd_kernel<<<ceil_div(B * 6, 64), 64, 0, (cudaStream_t)stream>>>(axisangle, translation, grad_T, B, invert, grad_axisangle, grad_translation);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_box_pool(const float* img, int planes, int H, int W, int k, float* out, void* stream) {
  FSNET_REQUIRE(img && out && planes > 0 && k >= 1 && H % k == 0 && W % k == 0 && H >= k && W >= k, "fsnet_box_pool: bad arguments");
  const size_t total = (size_t)planes * (H / k) * (W / k);
  FSNET_REQUIRE(total < (1ull << 32), "fsnet_box_pool: tensor too large for 32-bit indexing");
  box_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, planes, H, W, k, out);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
