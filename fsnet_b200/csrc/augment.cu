// Training augmentation of a batch of uint8 frame triplets on the device: affine warp or resize + pad (bilinear) +
// horizontal mirror + colour chain (brightness / contrast / saturation through HSV, in the drawn order) + normalisation, and
// the nearest-neighbour warp of the validity mask -- the pixel work of the reference's CPU list
// (vision_base/data/augmentations/augmentations.py:91-109,200-226,377-498,527-592), whose arithmetic is OpenCV's:
//   * cv2.warpAffine: source coordinates in 10-bit fixed point, rint(coef * 1024) evaluated in double, rounded to 1/32 pixel
//     (bilinear) or to the pixel (nearest); bilinear weights = float products of (1 - k/32, k/32); constant border 0;
//   * cv2.resize: coordinate (d + 0.5) * scale - 0.5 in double, float fraction, horizontal then vertical pass; nearest = floor(d * scale);
//   * cv2.cvtColor RGB<->HSV on float32: H in [0, 360), S = (V - min) / (|V| + eps), no clipping.
// Every float operation is written with explicit round-to-nearest intrinsics so that no FMA contraction changes the result
// relative to the numpy restatement (oracle/augment_oracle.py), which is bit-exact with cv2 for the warp.
// HBM bound and tiny: 9 B read + 72 B written per output pixel of a triplet; one launch per batch.
#include <cfloat>
#include "common.cuh"

namespace fsnet {
namespace {

constexpr int kPlan = 16;      // [0:6] geometry, 6 mirror, 7:10 op codes, 10:13 op values, 13 h0, 14 w0, 15 geometry mode
constexpr int OP_BRIGHTNESS = 1, OP_CONTRAST = 2, OP_SATURATION = 3;

__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  v = fmaxf(fmaxf(r, g), b);
  const float diff = __fsub_rn(v, fminf(fminf(r, g), b));
  s = __fdiv_rn(diff, __fadd_rn(fabsf(v), FLT_EPSILON));
  const float d = __fdiv_rn(60.f, __fadd_rn(diff, FLT_EPSILON));
  if (v == r) h = __fmul_rn(__fsub_rn(g, b), d);
  else if (v == g) h = __fadd_rn(__fmul_rn(__fsub_rn(b, r), d), 120.f);
  else h = __fadd_rn(__fmul_rn(__fsub_rn(r, g), d), 240.f);
  if (h < 0.f) h = __fadd_rn(h, 360.f);
}

__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
  if (s == 0.f) { r = g = b = v; return; }
  h = __fmul_rn(h, (float)(6.0 / 360.0));
  if (h < 0.f) h = __fadd_rn(h, __fmul_rn(6.f, ceilf(__fdiv_rn(-h, 6.f))));
  if (h >= 6.f) h = __fsub_rn(h, __fmul_rn(6.f, floorf(__fdiv_rn(h, 6.f))));
  float sector_f = floorf(h);
  float f = __fsub_rn(h, sector_f);
  int sector = (int)sector_f;
  if (!(sector >= 0 && sector < 6)) { sector = 0; f = 0.f; }
  const float t0 = v;
  const float t1 = __fmul_rn(v, __fsub_rn(1.f, s));
  const float t2 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, f)));
  const float t3 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f))));
  switch (sector) {            // (b, g, r) = tab[{1,3,0}, {1,0,2}, {3,0,1}, {0,2,1}, {0,1,3}, {2,1,0}]
    case 0: b = t1; g = t3; r = t0; break;
    case 1: b = t1; g = t0; r = t2; break;
    case 2: b = t3; g = t0; r = t1; break;
    case 3: b = t0; g = t2; r = t1; break;
    case 4: b = t0; g = t1; r = t3; break;
    default: b = t2; g = t1; r = t0; break;
  }
}

__device__ __forceinline__ long long fixed(double a, double t, double c) {      // rint((a * t + c) * 1024), products unfused
  return __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(a, t), c), 1024.0));
}

__global__ void __launch_bounds__(256) augment_frames_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ mask,
                                                             const double* __restrict__ plan, int B, int F, int H0, int W0,
                                                             int H, int W, const float* __restrict__ mean_std,
                                                             float* __restrict__ image, float* __restrict__ original,
                                                             double* __restrict__ mask_out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= W || y >= H) return;
  const double* p = plan + (size_t)b * kPlan;
  const bool mirror = p[6] != 0.0;
  const int h0 = (int)p[13], w0 = (int)p[14];
  const bool resize = (int)p[15] == 1;
  const int xm = mirror ? W - 1 - x : x;                     // RandomMirror acts on the finished (padded) frame
  const double xs = (double)xm, ys = (double)y;
  int sx, sy, sx1, sy1;                                        // tap columns / rows
  float wa, wb, wc, wd;                                        // affine: the four tap weights; resize: (1-fx, fx, 1-fy, fy)
  bool ok00, ok01, ok10, ok11;
  if (!resize) {
    // cv2.warpAffine: adelta / bdelta per column, X0 / Y0 per row (+ the rounding offset of the interpolation mode)
    const long long adelta = __double2ll_rn(__dmul_rn(__dmul_rn(p[0], xs), 1024.0));
    const long long bdelta = __double2ll_rn(__dmul_rn(__dmul_rn(p[3], xs), 1024.0));
    const long long X0 = fixed(p[1], ys, p[2]), Y0 = fixed(p[4], ys, p[5]);
    if (mask_out != nullptr) {
      const long long mx = (X0 + 512 + adelta) >> 10, my = (Y0 + 512 + bdelta) >> 10;
      const bool in = mx >= 0 && mx < w0 && my >= 0 && my < h0;
      mask_out[((size_t)b * H + y) * W + x] = in ? (double)mask[((size_t)b * H0 + my) * W0 + mx] : 0.0;
    }
    const long long X = (X0 + 16 + adelta) >> 5, Y = (Y0 + 16 + bdelta) >> 5;
    sx = (int)max(min(X >> 5, 32767LL), -32768LL);            // saturate_cast<short>
    sy = (int)max(min(Y >> 5, 32767LL), -32768LL);
    sx1 = sx + 1;
    sy1 = sy + 1;
    const float fx = (float)(X & 31) * (1.f / 32.f), fy = (float)(Y & 31) * (1.f / 32.f);
    wa = __fmul_rn(__fsub_rn(1.f, fy), __fsub_rn(1.f, fx));
    wb = __fmul_rn(__fsub_rn(1.f, fy), fx);
    wc = __fmul_rn(fy, __fsub_rn(1.f, fx));
    wd = __fmul_rn(fy, fx);
    const bool okx0 = sx >= 0 && sx < w0, okx1 = sx1 >= 0 && sx1 < w0, oky0 = sy >= 0 && sy < h0, oky1 = sy1 >= 0 && sy1 < h0;
    ok00 = oky0 && okx0; ok01 = oky0 && okx1; ok10 = oky1 && okx0; ok11 = oky1 && okx1;
  } else {
    // cv2.resize to (w_eff, h_eff) at the top-left of a zero canvas: coordinate (d + 0.5) * scale - 0.5 in double, float fraction
    const int w_eff = (int)p[2], h_eff = (int)p[3];
    const bool in = xm < w_eff && y < h_eff;
    if (mask_out != nullptr) {
      const int mx = min((int)floor(__dmul_rn(xs, p[0])), w0 - 1), my = min((int)floor(__dmul_rn(ys, p[1])), h0 - 1);
      mask_out[((size_t)b * H + y) * W + x] = in ? (double)mask[((size_t)b * H0 + my) * W0 + mx] : 0.0;
    }
    const double cx = __dadd_rn(__dmul_rn(__dadd_rn(xs, 0.5), p[0]), -0.5), cy = __dadd_rn(__dmul_rn(__dadd_rn(ys, 0.5), p[1]), -0.5);
    const double flx = floor(cx), fly = floor(cy);
    sx = (int)flx;
    sy = (int)fly;
    float fx = (float)__dadd_rn(cx, -flx), fy = (float)__dadd_rn(cy, -fly);
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= w0 - 1) { sx = w0 - 1; fx = 0.f; }
    if (sy < 0) { sy = 0; fy = 0.f; }
    if (sy >= h0 - 1) { sy = h0 - 1; fy = 0.f; }
    sx1 = min(sx + 1, w0 - 1);
    sy1 = min(sy + 1, h0 - 1);
    wa = __fsub_rn(1.f, fx); wb = fx; wc = __fsub_rn(1.f, fy); wd = fy;
    ok00 = ok01 = ok10 = ok11 = in;
  }
  const float mean[3] = {mean_std[0], mean_std[1], mean_std[2]}, stdv[3] = {mean_std[3], mean_std[4], mean_std[5]};
  const size_t plane = (size_t)H * W, pix = (size_t)y * W + x;
  for (int f = 0; f < F; ++f) {
    const uint8_t* src = frames + ((size_t)b * F + f) * H0 * W0 * 3;
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t00 = ok00 ? (float)src[((size_t)sy * W0 + sx) * 3 + c] : 0.f;
      const float t01 = ok01 ? (float)src[((size_t)sy * W0 + sx1) * 3 + c] : 0.f;
      const float t10 = ok10 ? (float)src[((size_t)sy1 * W0 + sx) * 3 + c] : 0.f;
      const float t11 = ok11 ? (float)src[((size_t)sy1 * W0 + sx1) * 3 + c] : 0.f;
      if (!resize)
        rgb[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t00, wa), __fmul_rn(t01, wb)), __fmul_rn(t10, wc)), __fmul_rn(t11, wd));
      else      // horizontal pass, then vertical pass
        rgb[c] = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(t00, wa), __fmul_rn(t01, wb)), wc),
                           __fmul_rn(__fadd_rn(__fmul_rn(t10, wa), __fmul_rn(t11, wb)), wd));
    }
    float* o = original + ((size_t)f * B + b) * 3 * plane + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = __fdiv_rn(rgb[c], 255.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int code = (int)p[7 + k];
      const double value = p[10 + k];
      if (code == OP_BRIGHTNESS) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = __fadd_rn(rgb[c], (float)value);
      } else if (code == OP_CONTRAST) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = __fmul_rn(rgb[c], (float)value);
      } else if (code == OP_SATURATION) {
        float h, s, v;
        rgb_to_hsv(rgb[0], rgb[1], rgb[2], h, s, v);
        if (value == value) s = __fmul_rn(s, (float)value);          // NaN = factor not drawn: the round trip alone
        hsv_to_rgb(h, s, v, rgb[0], rgb[1], rgb[2]);
      }
    }
    float* im = image + ((size_t)f * B + b) * 3 * plane + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) im[c * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(rgb[c], 255.f), mean[c]), stdv[c]);
  }
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

extern "C" int fsnet_augment_frames(const uint8_t* frames, const uint8_t* mask, const double* plan, int B, int F, int H0, int W0,
                                    int H, int W, const float* mean_std, float* image, float* original, double* mask_out,
                                    void* stream) {
  FSNET_REQUIRE(frames && plan && mean_std && image && original, "fsnet_augment_frames: null pointer");
  FSNET_REQUIRE(B > 0 && F > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0, "fsnet_augment_frames: empty shape (B=%d F=%d %dx%d -> %dx%d)",
                B, F, H0, W0, H, W);
  FSNET_REQUIRE(B <= 65535, "fsnet_augment_frames: batch %d exceeds the grid limit", B);
  FSNET_REQUIRE((mask != nullptr) == (mask_out != nullptr), "fsnet_augment_frames: mask input and output go together");
  dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), B);
  augment_frames_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(frames, mask, plan, B, F, H0, W0, H, W, mean_std, image, original,
                                                                  mask_out);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
