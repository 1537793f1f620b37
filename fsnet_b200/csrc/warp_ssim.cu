// Fused photometric-reprojection loss for one scale: depth up-sample -> back-project -> project ->
// bilinear border gather of both source frames -> nearest overlap-mask gather -> 3x3 reflect-padded
// SSIM + L1 -> min over {identity, reprojection} x {+1,-1} -> patched-mask weighted sums; and its
// backward by recomputation.  Replaces monodepth2_decoder.py:61-128,205-292 of the reference
// (see include/fsnet_b200.h for the per-entry citations, tests/loss_math_ref.py for the derivation).
//
// Work decomposition ("marching warps"): one warp owns a strip of columns and walks down a chunk of
// rows.  Lane l holds column x0+l-HALO; horizontal 3-tap sums take the neighbours' raw values with
// warp shuffles, vertical 3-tap sums are rolling registers, so the 3x3 SSIM window never touches
// shared memory and no block-level barrier exists.  The halo lanes/rows recompute the warp of the
// reflected pixel, exactly what ReflectionPad2d(1) on the *warped* image means.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

namespace fsnet {
namespace {

#ifndef LOSS_BWD_OCC
#define LOSS_BWD_OCC 3
#endif
constexpr int kWarps = 4;                 // warps per block
constexpr float k81C1 = 81.f * 1e-4f;     // 81 * 0.01^2   (sums instead of means: everything scaled by 9^2)
constexpr float k81C2 = 81.f * 9e-4f;     // 81 * 0.03^2

struct LossParams {
  const float* depth; int hs, ws;
  const float* tgt; const float* src0; const float* src1;   // fp32 NCHW (identity kernel only)
  const float4* packed; float4* packed_out;                 // [3][B,H,W] RGBX: target, source 0, source 1
  const void* mask; int mask_dtype;
  const float* cam; const float* ident; const float* noise; const float* motion;
  unsigned flags; int B, H, W;
  double* accum; uint8_t* sel; float* pred0; float* ident_out;
  const double* accum_in; const float* gout; float* grad_depth; float* grad_P;
  int rows_per_item, n_strips, n_chunks;
  float sy, sx;                           // (hs-1)/(H-1), (ws-1)/(W-1): align_corners=True scales
  const float4* lut; const int* lut_idx;  // MEI camera: per-pixel ray table (X, Y, Z, mask) and sample -> table index
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  i = i < 0 ? -i : i;
  i = i >= n ? 2 * n - 2 - i : i;
  return min(max(i, 0), n - 1);
}

__device__ __forceinline__ const void* mask_dtype_ptr_add(const void* m, int dtype, size_t n) {
  return reinterpret_cast<const char*>(m) + n * (dtype == FSNET_MASK_F64 ? 8 : 4);
}
__device__ __forceinline__ float load_mask(const void* m, int dtype, size_t i) {
  if (dtype == FSNET_MASK_F64) return (float)__ldg(reinterpret_cast<const double*>(m) + i);
  return __ldg(reinterpret_cast<const float*>(m) + i);
}

// align_corners=True bilinear read of the scale-s depth map at full-resolution pixel (y, x)
struct UpW { int i0, i1; float l; };
__device__ __forceinline__ UpW up_weights(int o, float scale, int n_in) {
  UpW w;
  float s = scale * (float)o;
  w.i0 = min((int)s, n_in - 1);
  w.i1 = w.i0 < n_in - 1 ? w.i0 + 1 : w.i0;
  w.l = s - (float)w.i0;
  return w;
}
__device__ __forceinline__ float depth_at(const float* d, int ws, const UpW& wy, const UpW& wx) {
  const float* r0 = d + (size_t)wy.i0 * ws;
  const float* r1 = d + (size_t)wy.i1 * ws;
  float top = (1.f - wx.l) * __ldg(r0 + wx.i0) + wx.l * __ldg(r0 + wx.i1);
  float bot = (1.f - wx.l) * __ldg(r1 + wx.i0) + wx.l * __ldg(r1 + wx.i1);
  return (1.f - wy.l) * top + wy.l * bot;
}

struct Geo { float r[3]; float c[3]; };      // ray inv(K)(x,y,1) and camera point r*D
__device__ __forceinline__ Geo geometry(const float* ik, float x, float y, float D) {
  Geo g;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g.r[i] = fmaf(ik[3 * i], x, fmaf(ik[3 * i + 1], y, ik[3 * i + 2]));
    g.c[i] = g.r[i] * D;
  }
  return g;
}

// ---- MEI (unified omnidirectional) camera: FishEyeDecoder, mei_fisheye_utils.py:14-51 ---------------------
// cam block of a frame in MEI mode: [0..6] = gamma1, gamma2, u0, v0, xi, k1, k2; [9..20] = T[:3, :4].
// The ray of a pixel comes from the calibration's look-up table (X, Y, Z) instead of inv(K)(x, y, 1).
__device__ __forceinline__ Geo geometry_lut(const float4& L, float D) {
  Geo g;
  g.r[0] = L.x; g.r[1] = L.y; g.r[2] = L.z;
#pragma unroll
  for (int i = 0; i < 3; ++i) g.c[i] = g.r[i] * D;
  return g;
}
struct Mei { float inv_n, n, mx, my, inv_d, ro2, dd; };
// _cam2image (mei_fisheye_utils.py:23-51): unit sphere -> shifted plane -> radial distortion -> pixel
__device__ __forceinline__ Mei mei_project(const float* in, float px, float py, float pz, float& u, float& v) {
  Mei m;
  m.n = sqrtf(fmaf(px, px, fmaf(py, py, pz * pz)));
  m.inv_n = __frcp_rn(m.n + 1e-6f);
  const float qx = px * m.inv_n, qy = py * m.inv_n, qz = pz * m.inv_n;
  m.inv_d = __frcp_rn((qz + in[4]) + 1e-6f);
  m.mx = qx * m.inv_d; m.my = qy * m.inv_d;
  m.ro2 = fmaf(m.mx, m.mx, m.my * m.my);
  m.dd = 1.f + in[5] * m.ro2 + in[6] * m.ro2 * m.ro2;
  u = fmaf(in[0], m.mx * m.dd, in[2]);
  v = fmaf(in[1], m.my * m.dd, in[3]);
  return m;
}
// forward-mode derivative of (u, v) along a = d p / d D
__device__ __forceinline__ void mei_jvp(const float* in, const Mei& m, float px, float py, float pz,
                                        float ax, float ay, float az, float& du, float& dv) {
  const float dn = (px * ax + py * ay + pz * az) * __frcp_rn(fmaxf(m.n, 1e-30f));
  const float dinv = -dn * m.inv_n * m.inv_n;
  const float dqx = fmaf(ax, m.inv_n, px * dinv), dqy = fmaf(ay, m.inv_n, py * dinv), dqz = fmaf(az, m.inv_n, pz * dinv);
  const float dmx = (dqx - m.mx * dqz) * m.inv_d, dmy = (dqy - m.my * dqz) * m.inv_d;
  const float ddd = (in[5] + 2.f * in[6] * m.ro2) * 2.f * (m.mx * dmx + m.my * dmy);
  du = in[0] * fmaf(dmx, m.dd, m.mx * ddd);
  dv = in[1] * fmaf(dmy, m.dd, m.my * ddd);
}
// reverse mode: (gu, gv) -> d loss / d p
__device__ __forceinline__ void mei_vjp(const float* in, const Mei& m, float px, float py, float pz,
                                        float gu, float gv, float (&gp)[3]) {
  const float gdx = in[0] * gu, gdy = in[1] * gv;
  const float gro2 = (gdx * m.mx + gdy * m.my) * (in[5] + 2.f * in[6] * m.ro2);
  const float gmx = fmaf(gdx, m.dd, 2.f * m.mx * gro2), gmy = fmaf(gdy, m.dd, 2.f * m.my * gro2);
  const float gqx = gmx * m.inv_d, gqy = gmy * m.inv_d, gqz = -(gmx * m.mx + gmy * m.my) * m.inv_d;
  const float ginv = gqx * px + gqy * py + gqz * pz;
  const float gn_over_n = -ginv * m.inv_n * m.inv_n * __frcp_rn(fmaxf(m.n, 1e-30f));
  gp[0] = fmaf(gqx, m.inv_n, px * gn_over_n);
  gp[1] = fmaf(gqy, m.inv_n, py * gn_over_n);
  gp[2] = fmaf(gqz, m.inv_n, pz * gn_over_n);
}

// One source frame at one pixel.  GRAD=0: colour + validity.  GRAD=1: also d pred/d ix, d pred/d iy
// (already zeroed where grid_sample's border clamp blocks the gradient) and d(u,v)/dD.
template <int GRAD>
struct Sample {
  float pred[3];
  bool valid;
  float dix[GRAD ? 3 : 1], diy[GRAD ? 3 : 1];
  float du, dv;                // d u / d D, d v / d D
  float px, py, rz;            // projected point (x, y) and 1/(z+eps)
};

template <int GRAD>
__device__ __forceinline__ Sample<GRAD> sample_frame(const float* __restrict__ src, const float* P, const Geo& g,
                                                      const void* mask, int mask_dtype, bool want_valid,
                                                      int H, int W) {
  Sample<GRAD> o;
  const size_t HW = (size_t)H * W;
  float px = fmaf(P[0], g.c[0], fmaf(P[1], g.c[1], fmaf(P[2], g.c[2], P[3])));
  float py = fmaf(P[4], g.c[0], fmaf(P[5], g.c[1], fmaf(P[6], g.c[2], P[7])));
  float pz = fmaf(P[8], g.c[0], fmaf(P[9], g.c[1], fmaf(P[10], g.c[2], P[11])));
  float rz = __frcp_rn(pz + 1e-7f);
  float ix = px * rz, iy = py * rz;      // == grid_sample's un-normalised coordinate (align_corners=True)
  const float xm = (float)(W - 1), ym = (float)(H - 1);
  o.valid = true;
  if (want_valid) {                       // nearest, zeros padding, "== 1" (monodepth2_decoder.py:113-116)
    float xn = rintf(ix), yn = rintf(iy);
    bool inb = (xn >= 0.f) && (xn <= xm) && (yn >= 0.f) && (yn <= ym);
    o.valid = inb && (mask == nullptr || load_mask(mask, mask_dtype, (size_t)(int)yn * W + (int)xn) == 1.f);
  }
  float ixc = fminf(fmaxf(ix, 0.f), xm), iyc = fminf(fmaxf(iy, 0.f), ym);   // padding_mode='border'
  float x0f = floorf(ixc), y0f = floorf(iyc);
  float fx = ixc - x0f, fy = iyc - y0f;
  int x0 = (int)x0f, y0 = (int)y0f;
  int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  const float* r0 = src + (size_t)y0 * W;
  const float* r1 = src + (size_t)y1 * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float nw = __ldg(r0 + c * HW + x0), ne = __ldg(r0 + c * HW + x1);
    float sw = __ldg(r1 + c * HW + x0), se = __ldg(r1 + c * HW + x1);
    float top = fmaf(fx, ne - nw, nw), bot = fmaf(fx, se - sw, sw);
    o.pred[c] = fmaf(fy, bot - top, top);
    if (GRAD) {
      float dx_top = ne - nw, dx_bot = se - sw;
      o.dix[c] = fmaf(fy, dx_bot - dx_top, dx_top);
      o.diy[c] = bot - top;
    }
  }
  o.px = px; o.py = py; o.rz = rz;
  o.du = 0.f; o.dv = 0.f;
  if (GRAD) {
    // clip_coordinates_set_grad: zero gradient on and outside the border
    bool mx = (ix > 0.f) && (ix < xm), my = (iy > 0.f) && (iy < ym);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o.dix[c] = mx ? o.dix[c] : 0.f;
      o.diy[c] = my ? o.diy[c] : 0.f;
    }
    float ax = fmaf(P[0], g.r[0], fmaf(P[1], g.r[1], P[2] * g.r[2]));
    float ay = fmaf(P[4], g.r[0], fmaf(P[5], g.r[1], P[6] * g.r[2]));
    float az = fmaf(P[8], g.r[0], fmaf(P[9], g.r[1], P[10] * g.r[2]));
    o.du = (ax - ix * az) * rz;          // (ax*z - px*az)/z^2
    o.dv = (ay - iy * az) * rz;
  }
  return o;
}


// ---- row-skewed loads (software pipeline) -------------------------------------------------------------
// The images are read from the RGBX-packed copy written once per step by the identity kernel: one 128-bit
// load per bilinear corner.  issue_* only computes addresses and issues loads; finish_* consumes them one
// loop iteration later, so a row's gathers are in flight while the previous row's SSIM arithmetic runs.
struct DepthLoads { float d00, d01, d10, d11, ly; float4 L; };
// lut_row: this sample's ray table at column xr (MEI camera) or nullptr (pinhole)
__device__ __forceinline__ DepthLoads issue_depth(const float* d, int ws, int hs, float sy, int yr, const UpW& wx,
                                                  const float4* lut_col = nullptr, int W = 0) {
  UpW wy = up_weights(yr, sy, hs);
  const float* r0 = d + (size_t)wy.i0 * ws;
  const float* r1 = d + (size_t)wy.i1 * ws;
  DepthLoads o;
  o.d00 = __ldg(r0 + wx.i0); o.d01 = __ldg(r0 + wx.i1); o.d10 = __ldg(r1 + wx.i0); o.d11 = __ldg(r1 + wx.i1); o.ly = wy.l;
  o.L = lut_col ? __ldg(lut_col + (size_t)yr * W) : make_float4(0.f, 0.f, 0.f, 0.f);
  return o;
}
__device__ __forceinline__ float finish_depth(const DepthLoads& o, float lx) {
  float top = (1.f - lx) * o.d00 + lx * o.d01, bot = (1.f - lx) * o.d10 + lx * o.d11;
  return (1.f - o.ly) * top + o.ly * bot;
}

struct FrameLoads {
  float4 nw, ne, sw, se;
  float fx, fy, ix, iy;
  float du, dv;                // d(u, v) / d D (gradient variant only)
  float mval; bool inb;
};
// `in` = the 9 leading floats of the frame's camera block (inv(K) for the pinhole camera, MEI intrinsics otherwise);
// lut_b = this sample's ray table (its .w = validity of the source pixel, folded into the overlap mask,
// monodepth2_decoder.py:409) or nullptr.
template <int GRAD, int CAM>
__device__ __forceinline__ FrameLoads issue_frame(const float4* __restrict__ src, const float* in, const float* P, const Geo& g,
                                                  const void* mask, int mask_dtype, bool want_valid, int H, int W,
                                                  const float4* lut_b) {
  FrameLoads o;
  float px = fmaf(P[0], g.c[0], fmaf(P[1], g.c[1], fmaf(P[2], g.c[2], P[3])));
  float py = fmaf(P[4], g.c[0], fmaf(P[5], g.c[1], fmaf(P[6], g.c[2], P[7])));
  float pz = fmaf(P[8], g.c[0], fmaf(P[9], g.c[1], fmaf(P[10], g.c[2], P[11])));
  float ix, iy, rz = 0.f;
  Mei mei;
  if (CAM == 0) {
    rz = __frcp_rn(pz + 1e-7f);
    ix = px * rz; iy = py * rz;
  } else {
    mei = mei_project(in, px, py, pz, ix, iy);
  }
  const float xm = (float)(W - 1), ym = (float)(H - 1);
  o.inb = true; o.mval = 1.f;
  if (want_valid) {
    float xn = rintf(ix), yn = rintf(iy);
    o.inb = (xn >= 0.f) && (xn <= xm) && (yn >= 0.f) && (yn <= ym);
    if (o.inb && mask != nullptr) o.mval = load_mask(mask, mask_dtype, (size_t)(int)yn * W + (int)xn);
    if (CAM == 1 && o.inb) o.mval *= __ldg(reinterpret_cast<const float*>(lut_b + (size_t)(int)yn * W + (int)xn) + 3);
  }
  float ixc = fminf(fmaxf(ix, 0.f), xm), iyc = fminf(fmaxf(iy, 0.f), ym);
  float x0f = floorf(ixc), y0f = floorf(iyc);
  o.fx = ixc - x0f; o.fy = iyc - y0f;
  int x0 = (int)x0f, y0 = (int)y0f;
  int dx = x0 < W - 1 ? 1 : 0, dy = y0 < H - 1 ? W : 0;
  const float4* p00 = src + (size_t)y0 * W + x0;
  o.nw = __ldg(p00); o.ne = __ldg(p00 + dx); o.sw = __ldg(p00 + dy); o.se = __ldg(p00 + dy + dx);
  o.ix = ix; o.iy = iy;
  o.du = o.dv = 0.f;
  if (GRAD) {
    const float ax = fmaf(P[0], g.r[0], fmaf(P[1], g.r[1], P[2] * g.r[2]));
    const float ay = fmaf(P[4], g.r[0], fmaf(P[5], g.r[1], P[6] * g.r[2]));
    const float az = fmaf(P[8], g.r[0], fmaf(P[9], g.r[1], P[10] * g.r[2]));
    if (CAM == 0) {
      o.du = (ax - ix * az) * rz;          // (ax*z - px*az)/z^2
      o.dv = (ay - iy * az) * rz;
    } else {
      mei_jvp(in, mei, px, py, pz, ax, ay, az, o.du, o.dv);
    }
  }
  return o;
}
template <int GRAD>
__device__ __forceinline__ Sample<GRAD> finish_frame(const FrameLoads& o, int H, int W) {
  Sample<GRAD> s;
  const float nw[3] = {o.nw.x, o.nw.y, o.nw.z}, ne[3] = {o.ne.x, o.ne.y, o.ne.z};
  const float sw[3] = {o.sw.x, o.sw.y, o.sw.z}, se[3] = {o.se.x, o.se.y, o.se.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float top = fmaf(o.fx, ne[c] - nw[c], nw[c]), bot = fmaf(o.fx, se[c] - sw[c], sw[c]);
    s.pred[c] = fmaf(o.fy, bot - top, top);
    if (GRAD) {
      float dx_top = ne[c] - nw[c], dx_bot = se[c] - sw[c];
      s.dix[c] = fmaf(o.fy, dx_bot - dx_top, dx_top);
      s.diy[c] = bot - top;
    }
  }
  s.valid = o.inb && (o.mval == 1.f);
  s.px = o.ix; s.py = o.iy; s.rz = 0.f;
  s.du = 0.f; s.dv = 0.f;
  if (GRAD) {
    bool mx = (o.ix > 0.f) && (o.ix < (float)(W - 1)), my = (o.iy > 0.f) && (o.iy < (float)(H - 1));
#pragma unroll
    for (int c = 0; c < 3; ++c) { s.dix[c] = mx ? s.dix[c] : 0.f; s.diy[c] = my ? s.diy[c] : 0.f; }
    s.du = o.du;
    s.dv = o.dv;
  }
  return s;
}
struct RowLoads { float4 t; FrameLoads f0, f1; float D; };
template <int GRAD, int CAM>
__device__ __forceinline__ RowLoads issue_row(const LossParams& p, const float4* tg, const float4* s0, const float4* s1,
                                              const float* ik, const float* P0, const float* P1, const void* mask_b,
                                              bool overlap, int yr, int xr, const DepthLoads& dl, float lx, const float4* lut_b) {
  RowLoads r;
  const float D = finish_depth(dl, lx);
  r.t = __ldg(tg + (size_t)yr * p.W + xr);
  Geo g = CAM == 0 ? geometry(ik, (float)xr, (float)yr, D) : geometry_lut(dl.L, D);
  r.f0 = issue_frame<GRAD, CAM>(s0, ik, P0, g, mask_b, p.mask_dtype, overlap, p.H, p.W, lut_b);
  r.f1 = issue_frame<GRAD, CAM>(s1, ik, P1, g, mask_b, p.mask_dtype, overlap, p.H, p.W, lut_b);
  r.D = D;
  return r;
}

// SSIM loss value of (x = pred, t = target) from the 3x3 SUMS (not means).
__device__ __forceinline__ float ssim_sums(float Sx, float Sxx, float Sxt, float St, float Stt) {
  float sxst = Sx * St;
  float sq = fmaf(Sx, Sx, St * St);
  float A1 = fmaf(2.f, sxst, k81C1);
  float A2 = fmaf(18.f, Sxt, k81C2) - 2.f * sxst;
  float B1 = sq + k81C1;
  float B2 = fmaf(9.f, Sxx + Stt, k81C2) - sq;
  return __saturatef(fmaf(-0.5f, __fdividef(A1 * A2, B1 * B2), 0.5f));
}
// value and partial derivatives with respect to Sx, Sxx, Sxt (tests/loss_math_ref.py::ssim_and_partials)
__device__ __forceinline__ float ssim_sums_grad(float Sx, float Sxx, float Sxt, float St, float Stt,
                                                float& dSx, float& dSxx, float& dSxt) {
  float sxst = Sx * St;
  float sq = fmaf(Sx, Sx, St * St);
  float A1 = fmaf(2.f, sxst, k81C1);
  float A2 = fmaf(18.f, Sxt, k81C2) - 2.f * sxst;
  float B1 = sq + k81C1;
  float B2 = fmaf(9.f, Sxx + Stt, k81C2) - sq;
  float inv = __fdividef(1.f, B1 * B2);
  float Q = A1 * A2 * inv;
  float raw = fmaf(-0.5f, Q, 0.5f);
  float live = (raw >= 0.f && raw <= 1.f) ? -0.5f : 0.f;
  dSx = live * 2.f * inv * (St * (A2 - A1) - Q * Sx * (B2 - B1));
  dSxx = live * (-9.f) * __fdividef(Q, B2);
  dSxt = live * 18.f * A1 * inv;
  return __saturatef(raw);
}

// horizontal 3-tap sums of the 24 SSIM moments from this lane's raw values and its two neighbours'
// raw[0..2] = target, raw[3..5] = image 0, raw[6..8] = image 1
__device__ __forceinline__ void hsum24(const float (&raw)[9], float (&h)[24]) {
  float l[9], r[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    l[i] = __shfl_up_sync(0xffffffffu, raw[i], 1);
    r[i] = __shfl_down_sync(0xffffffffu, raw[i], 1);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    h[c] = l[c] + raw[c] + r[c];                                           // St
    h[3 + c] = fmaf(l[c], l[c], fmaf(raw[c], raw[c], r[c] * r[c]));        // Stt
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int j = 3 + 3 * k + c;
      h[6 + 9 * k + c] = l[j] + raw[j] + r[j];                                       // Sx
      h[6 + 9 * k + 3 + c] = fmaf(l[j], l[j], fmaf(raw[j], raw[j], r[j] * r[j]));    // Sxx
      h[6 + 9 * k + 6 + c] = fmaf(l[j], l[c], fmaf(raw[j], raw[c], r[j] * r[c]));    // Sxt
    }
  }
}

struct Item { int b, y_begin, y_end, strip; };
__device__ __forceinline__ bool decode_item(const LossParams& p, Item& it) {
  int item = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (item >= p.B * p.n_strips * p.n_chunks) return false;
  it.strip = item % p.n_strips;
  int rest = item / p.n_strips;
  int chunk = rest % p.n_chunks;
  it.b = rest / p.n_chunks;
  it.y_begin = chunk * p.rows_per_item;
  it.y_end = min(it.y_begin + p.rows_per_item, p.H);
  return true;
}

// ------------------------------------------------------------------------------------------------
// forward.  MODE 0: identity photometric map (pred_f := src_f at the same pixel, result stored);
//           MODE 1: reprojection loss (pred_f := warped src_f, result reduced into accum).
// Columns per warp: 30 (lanes 0 and 31 are the reflect / neighbour halo).
// ------------------------------------------------------------------------------------------------
template <int MODE, int CAM>
__global__ void __launch_bounds__(kWarps * 32, 4) loss_fwd_kernel(LossParams p) {
  __shared__ float s_cam[kWarps][42];
  Item it;
  if (!decode_item(p, it)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, W = p.W;
  const size_t HW = (size_t)H * W;
  const int b = it.b;
  const int x = it.strip * 30 + lane - 1;
  const int xr = reflect_idx(x, W);
  const bool out_lane = lane >= 1 && lane <= 30 && x < W;
  const void* mask = p.mask;
  const bool overlap = (p.flags & FSNET_FLAG_OVERLAP_MASK) != 0;
  const bool use_ident = (p.flags & FSNET_FLAG_MOTION_MASK) == 0;
  const void* mask_b = mask ? mask_dtype_ptr_add(mask, p.mask_dtype, (size_t)b * HW) : nullptr;

  float A1[24], A2[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) { A1[i] = 0.f; A2[i] = 0.f; }
  float l1_prev[2] = {0.f, 0.f};
  bool valid_prev[2] = {true, true};
  float acc_num = 0.f, acc_den = 0.f;
  const int y_first = it.y_begin - 1, y_last = it.y_end;

  // ---- MODE 0 state: plain NCHW loads -------------------------------------------------------------
  const float* tgt = MODE == 0 ? p.tgt + (size_t)b * 3 * HW : nullptr;
  const float* src0 = MODE == 0 ? p.src0 + (size_t)b * 3 * HW : nullptr;
  const float* src1 = MODE == 0 ? p.src1 + (size_t)b * 3 * HW : nullptr;
  // ---- MODE 1 state: packed images + software pipeline ----------------------------------------------
  const float* ik = s_cam[warp];
  const float* P0 = s_cam[warp] + 9;        // frame 0: inv(K) at [0,9), P at [9,21)
  const float* P1 = s_cam[warp] + 21 + 9;   // frame 1: inv(K) at [21,30), P at [30,42)
  UpW wx = {0, 0, 0.f};
  const float* depth = nullptr;
  const float4 *tg4 = nullptr, *s04 = nullptr, *s14 = nullptr;
  DepthLoads dl_next;
  const float4* lut_b = nullptr;            // MEI camera: this sample's ray table
  const float4* lut_col = nullptr;
  if (MODE == 1) {
    for (int i = lane; i < 42; i += 32) s_cam[warp][i] = __ldg(p.cam + (size_t)b * 42 + i);
    __syncwarp();
    wx = up_weights(xr, p.sx, p.ws);
    depth = p.depth + (size_t)b * p.hs * p.ws;
    tg4 = p.packed + ((size_t)0 * p.B + b) * HW;
    s04 = p.packed + ((size_t)1 * p.B + b) * HW;
    s14 = p.packed + ((size_t)2 * p.B + b) * HW;
    if (CAM == 1) {
      lut_b = p.lut + (size_t)(p.lut_idx ? __ldg(p.lut_idx + b) : 0) * HW;
      lut_col = lut_b + xr;
    }
    dl_next = issue_depth(depth, p.ws, p.hs, p.sy, reflect_idx(y_first, H), wx, lut_col, W);
  }

#pragma unroll 1
  for (int yy = y_first; yy <= y_last; ++yy) {
    const int yr = reflect_idx(yy, H);
    float raw[9];
    float l1[2];
    bool valid[2] = {true, true};
    // centre-pixel inputs (row yy-1) depend on nothing computed here: issue their loads first
    float c_i0 = 0.f, c_i1 = 0.f, c_n0 = 0.f, c_n1 = 0.f, c_m = 1.f;
    if (MODE == 1 && out_lane && yy >= it.y_begin + 1) {
      const size_t pix = (size_t)(yy - 1) * W + x;
      if (use_ident) {
        c_i0 = __ldg(p.ident + ((size_t)b * 2 + 0) * HW + pix);
        c_i1 = __ldg(p.ident + ((size_t)b * 2 + 1) * HW + pix);
        if (p.noise) {
          c_n0 = __ldg(p.noise + ((size_t)b * 2 + 0) * HW + pix);
          c_n1 = __ldg(p.noise + ((size_t)b * 2 + 1) * HW + pix);
        }
      }
      if (mask) c_m = load_mask(mask_b, p.mask_dtype, pix);
    }
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        raw[c] = __ldg(tgt + c * HW + (size_t)yr * W + xr);
        raw[3 + c] = __ldg(src0 + c * HW + (size_t)yr * W + xr);
        raw[6 + c] = __ldg(src1 + c * HW + (size_t)yr * W + xr);
      }
      if (p.packed_out != nullptr && out_lane && yy >= it.y_begin && yy < it.y_end) {
        const size_t pix = (size_t)yy * W + x;
        const float mw = (p.flags & FSNET_FLAG_PACKED_MASK) ? (mask_b ? load_mask(mask_b, p.mask_dtype, pix) : 1.f) : 0.f;
        p.packed_out[((size_t)0 * p.B + b) * HW + pix] = make_float4(raw[0], raw[1], raw[2], mw);
        p.packed_out[((size_t)1 * p.B + b) * HW + pix] = make_float4(raw[3], raw[4], raw[5], mw);
        p.packed_out[((size_t)2 * p.B + b) * HW + pix] = make_float4(raw[6], raw[7], raw[8], mw);
      }
    } else {
      // this row's depth (and ray) was requested one iteration ago: the gathers can go out immediately
      RowLoads rl = issue_row<0, CAM>(p, tg4, s04, s14, ik, P0, P1, mask_b, overlap, yr, xr, dl_next, wx.l, lut_b);
      dl_next = issue_depth(depth, p.ws, p.hs, p.sy, reflect_idx(min(yy + 1, y_last), H), wx, lut_col, W);
      Sample<0> s0 = finish_frame<0>(rl.f0, H, W);
      Sample<0> s1 = finish_frame<0>(rl.f1, H, W);
      raw[0] = rl.t.x; raw[1] = rl.t.y; raw[2] = rl.t.z;
#pragma unroll
      for (int c = 0; c < 3; ++c) { raw[3 + c] = s0.pred[c]; raw[6 + c] = s1.pred[c]; }
      valid[0] = s0.valid; valid[1] = s1.valid;
      if (p.pred0 != nullptr && b == 0 && out_lane && yy >= it.y_begin && yy < it.y_end) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p.pred0[c * HW + (size_t)yy * W + x] = s0.pred[c];
          p.pred0[(3 + c) * HW + (size_t)yy * W + x] = s1.pred[c];
        }
      }
    }
    l1[0] = fabsf(raw[0] - raw[3]) + fabsf(raw[1] - raw[4]) + fabsf(raw[2] - raw[5]);
    l1[1] = fabsf(raw[0] - raw[6]) + fabsf(raw[1] - raw[7]) + fabsf(raw[2] - raw[8]);

    float h[24];
    hsum24(raw, h);
    if (yy >= it.y_begin + 1) {            // three rows in: centre row yc = yy - 1
      const int yc = yy - 1;
      float ph[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s += ssim_sums(A2[6 + 9 * k + c] + h[6 + 9 * k + c], A2[6 + 9 * k + 3 + c] + h[6 + 9 * k + 3 + c],
                         A2[6 + 9 * k + 6 + c] + h[6 + 9 * k + 6 + c], A2[c] + h[c], A2[3 + c] + h[3 + c]);
        }
        ph[k] = fmaf(0.85f / 3.f, s, (0.15f / 3.f) * l1_prev[k]);
      }
      if (out_lane) {
        const size_t pix = (size_t)yc * W + x;
        if (MODE == 0) {
          p.ident_out[((size_t)b * 2 + 0) * HW + pix] = ph[0];
          p.ident_out[((size_t)b * 2 + 1) * HW + pix] = ph[1];
        } else {
          if (overlap) {
            ph[0] = valid_prev[0] ? ph[0] : 100.f;
            ph[1] = valid_prev[1] ? ph[1] : 100.f;
          }
          float best;
          int arg;
          if (use_ident) {
            const float i0 = fmaf(c_n0, 1e-5f, c_i0), i1 = fmaf(c_n1, 1e-5f, c_i1);
            best = i0; arg = 0;
            if (i1 < best) { best = i1; arg = 1; }
            if (ph[0] < best) { best = ph[0]; arg = 2; }
            if (ph[1] < best) { best = ph[1]; arg = 3; }
          } else {
            best = ph[0]; arg = 0;
            if (ph[1] < best) { best = ph[1]; arg = 1; }
          }
          acc_num = fmaf(best, c_m, acc_num);
          acc_den += c_m;
          if (p.sel) p.sel[(size_t)b * HW + pix] = (uint8_t)arg;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) { A2[i] = A1[i] + h[i]; A1[i] = h[i]; }
    l1_prev[0] = l1[0]; l1_prev[1] = l1[1];
    valid_prev[0] = valid[0]; valid_prev[1] = valid[1];
  }
  if (MODE == 1) {
    double n = warp_sum((double)acc_num), d = warp_sum((double)acc_den);
    if (lane == 0) {
      atomicAdd(p.accum, n);
      atomicAdd(p.accum + 1, d);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward.  Columns per warp: 28 (two halo lanes each side: one for SSIM, one for its adjoint).
// Pipeline per raw row yy: A) raw values + derivative bundle at row yy  B) SSIM partials and winner at
// yc = yy-1  C) adjoint box filter -> d pred, chain to depth (and pose) at yq = yy-2.
// POSE=1 additionally reduces d loss / d P (12 numbers per frame).
// ------------------------------------------------------------------------------------------------
template <int POSE, int CAM>
__global__ void __launch_bounds__(kWarps * 32, LOSS_BWD_OCC) loss_bwd_kernel(LossParams p) {
  constexpr int NV = POSE ? 22 : 15;        // delayed values per pixel
  __shared__ float s_cam[kWarps][42];
  __shared__ float s_delay[kWarps][3][NV][32];
  Item it;
  if (!decode_item(p, it)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, W = p.W;
  const size_t HW = (size_t)H * W;
  const int b = it.b;
  const int x = it.strip * 28 + lane - 2;
  const int xr = reflect_idx(x, W);
  const bool x_in = x >= 0 && x < W;
  const bool out_lane = lane >= 2 && lane <= 29 && x < W;
  const float4* tg4 = p.packed + ((size_t)0 * p.B + b) * HW;
  const float4* s04 = p.packed + ((size_t)1 * p.B + b) * HW;
  const float4* s14 = p.packed + ((size_t)2 * p.B + b) * HW;
  const bool overlap = (p.flags & FSNET_FLAG_OVERLAP_MASK) != 0;
  const bool use_ident = (p.flags & FSNET_FLAG_MOTION_MASK) == 0;
  const void* mask_b = p.mask ? mask_dtype_ptr_add(p.mask, p.mask_dtype, (size_t)b * HW) : nullptr;

  for (int i = lane; i < 42; i += 32) s_cam[warp][i] = __ldg(p.cam + (size_t)b * 42 + i);
  __syncwarp();
  const float* ik = s_cam[warp];
  const float* P0 = s_cam[warp] + 9;
  const float* P1 = s_cam[warp] + 30;
  const UpW wx = up_weights(xr, p.sx, p.ws);
  const float* depth = p.depth + (size_t)b * p.hs * p.ws;
  const bool full_res = (p.hs == H && p.ws == W);
  // d total / d (min value at a pixel) = gout * mask / (sum(mask) + 1e-6)
  const float gbase = (float)((double)__ldg(p.gout) / (__ldg(p.accum_in + 1) + 1e-6));

  float A1[24], A2[24], B1[18], B2[18];
#pragma unroll
  for (int i = 0; i < 24; ++i) { A1[i] = 0.f; A2[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 18; ++i) { B1[i] = 0.f; B2[i] = 0.f; }
  float l1_prev[2] = {0.f, 0.f};
  bool valid_prev[2] = {true, true};
  float gf_prev[2] = {0.f, 0.f};
  float gP[POSE ? 24 : 1];
#pragma unroll
  for (int i = 0; i < (POSE ? 24 : 1); ++i) gP[i] = 0.f;
  float acc_num = 0.f;                       // forward value of the owned pixels (fused forward+backward launches, p.accum != null)

  const int y_first = it.y_begin - 2, y_last = it.y_end + 1;
  const float4* lut_b = CAM == 1 ? p.lut + (size_t)(p.lut_idx ? __ldg(p.lut_idx + b) : 0) * HW : nullptr;
  const float4* lut_col = CAM == 1 ? lut_b + xr : nullptr;
  DepthLoads dl_next = issue_depth(depth, p.ws, p.hs, p.sy, reflect_idx(y_first, H), wx, lut_col, W);

#pragma unroll 1
  for (int yy = y_first; yy <= y_last; ++yy) {
    // centre-pixel inputs (row yy-1) depend on nothing computed here: issue their loads first
    const bool centre_live = yy >= it.y_begin && (yy - 1) >= 0 && (yy - 1) < H && x_in && lane >= 1 && lane <= 30;
    float c_i0 = 0.f, c_i1 = 0.f, c_n0 = 0.f, c_n1 = 0.f, c_m = 1.f, c_gate = 1.f;
    if (centre_live) {
      const size_t pix = (size_t)(yy - 1) * W + x;
      if (use_ident) {
        c_i0 = __ldg(p.ident + ((size_t)b * 2 + 0) * HW + pix);
        c_i1 = __ldg(p.ident + ((size_t)b * 2 + 1) * HW + pix);
        if (p.noise) {
          c_n0 = __ldg(p.noise + ((size_t)b * 2 + 0) * HW + pix);
          c_n1 = __ldg(p.noise + ((size_t)b * 2 + 1) * HW + pix);
        }
      } else {
        c_gate = 1.f - __ldg(p.motion + (size_t)b * HW + pix);
      }
      if (p.mask) c_m = load_mask(mask_b, p.mask_dtype, pix);
    }
    // ---- A: raw values at row yy (reflected / clamped) ------------------------------------------------
    float raw[9];
    float l1[2];
    bool valid[2];
    {
      RowLoads rl = issue_row<1, CAM>(p, tg4, s04, s14, ik, P0, P1, mask_b, overlap, reflect_idx(yy, H), xr, dl_next, wx.l, lut_b);
      dl_next = issue_depth(depth, p.ws, p.hs, p.sy, reflect_idx(min(yy + 1, y_last), H), wx, lut_col, W);
      const float D = rl.D;
      Sample<1> s0 = finish_frame<1>(rl.f0, H, W);
      Sample<1> s1 = finish_frame<1>(rl.f1, H, W);
      raw[0] = rl.t.x; raw[1] = rl.t.y; raw[2] = rl.t.z;
#pragma unroll
      for (int c = 0; c < 3; ++c) { raw[3 + c] = s0.pred[c]; raw[6 + c] = s1.pred[c]; }
      valid[0] = s0.valid; valid[1] = s1.valid;
      float(*slot)[32] = s_delay[warp][(yy + 3) % 3];
#pragma unroll
      for (int i = 0; i < 9; ++i) slot[i][lane] = raw[i];
      if (POSE) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          slot[9 + c][lane] = s0.dix[c];  slot[12 + c][lane] = s0.diy[c];
          slot[15 + c][lane] = s1.dix[c]; slot[18 + c][lane] = s1.diy[c];
        }
        slot[21][lane] = D;
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          slot[9 + c][lane] = fmaf(s0.dix[c], s0.du, s0.diy[c] * s0.dv);
          slot[12 + c][lane] = fmaf(s1.dix[c], s1.du, s1.diy[c] * s1.dv);
        }
      }
    }
    l1[0] = fabsf(raw[0] - raw[3]) + fabsf(raw[1] - raw[4]) + fabsf(raw[2] - raw[5]);
    l1[1] = fabsf(raw[0] - raw[6]) + fabsf(raw[1] - raw[7]) + fabsf(raw[2] - raw[8]);
    float h[24];
    hsum24(raw, h);

    // ---- B: SSIM partials and arg-min at the centre row yc = yy - 1 ------------------------------
    const int yc = yy - 1;
    float w[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) w[i] = 0.f;
    float gf[2] = {0.f, 0.f};
    if (centre_live) {
      float ph[2], da[2][3], db[2][3], dc[2][3];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s += ssim_sums_grad(A2[6 + 9 * k + c] + h[6 + 9 * k + c], A2[6 + 9 * k + 3 + c] + h[6 + 9 * k + 3 + c],
                              A2[6 + 9 * k + 6 + c] + h[6 + 9 * k + 6 + c], A2[c] + h[c], A2[3 + c] + h[3 + c],
                              da[k][c], db[k][c], dc[k][c]);
        }
        ph[k] = fmaf(0.85f / 3.f, s, (0.15f / 3.f) * l1_prev[k]);
        if (overlap && !valid_prev[k]) ph[k] = 100.f;
      }
      int win;                               // 0 / 1 = reprojection frame that wins, -1 = none
      const float gate = c_gate;
      float best;
      if (use_ident) {
        const float i0 = fmaf(c_n0, 1e-5f, c_i0), i1 = fmaf(c_n1, 1e-5f, c_i1);
        best = fminf(i0, i1);
        win = -1;
        if (ph[0] < best) { best = ph[0]; win = 0; }
        if (ph[1] < best) { best = ph[1]; win = 1; }
      } else {
        win = ph[1] < ph[0] ? 1 : 0;
        best = fminf(ph[0], ph[1]);
      }
      // the forward sum counts every pixel once: the rows and columns this warp owns
      if (p.accum != nullptr && out_lane && yc >= it.y_begin && yc < it.y_end) acc_num = fmaf(best, c_m, acc_num);
      if (win >= 0 && (!overlap || valid_prev[win])) {
        float gv = gbase * c_m * gate;
        float ws_ = (0.85f / 3.f) * gv;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k == win) {
            gf[k] = gv;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              w[9 * k + c] = ws_ * da[k][c];
              w[9 * k + 3 + c] = ws_ * db[k][c];
              w[9 * k + 6 + c] = ws_ * dc[k][c];
            }
          }
        }
      }
    }
    // ---- C: adjoint of the reflect-padded box filter, then the chain to depth at yq = yy - 2 -----
    const int yq = yy - 2;
    float G[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      float wl = __shfl_up_sync(0xffffffffu, w[i], 1);
      float wr = __shfl_down_sync(0xffffffffu, w[i], 1);
      float hw = wl + w[i] + wr;
      if (x == 1) hw += wl;                  // column 0 reflects onto column 1
      if (x == W - 2) hw += wr;              // column W-1 reflects onto column W-2
      float gsum = B2[i] + hw;
      if (yq == 1) gsum += B2[i] - B1[i];    // row 0 reflects onto row 1   (B2 - B1 = HW(row 0))
      if (yq == H - 2) gsum += hw;           // row H-1 reflects onto row H-2
      G[i] = gsum;
      B2[i] = B1[i] + hw;
      B1[i] = hw;
    }
    if (yq >= it.y_begin && yq < it.y_end && out_lane) {
      float(*slot)[32] = s_delay[warp][(yq + 3) % 3];
      float t[3], gD = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) t[c] = slot[c][lane];
      float gu[2] = {0.f, 0.f}, gv2[2] = {0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float pr = slot[3 + 3 * k + c][lane];
          float d = t[c] - pr;
          float sgn = (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f);
          float gp = fmaf(2.f * pr, G[9 * k + 3 + c], fmaf(t[c], G[9 * k + 6 + c], G[9 * k + c]));
          gp = fmaf(-(0.15f / 3.f) * gf_prev[k], sgn, gp);
          if (POSE) {
            gu[k] = fmaf(gp, slot[9 + 6 * k + c][lane], gu[k]);
            gv2[k] = fmaf(gp, slot[12 + 6 * k + c][lane], gv2[k]);
          } else {
            gD = fmaf(gp, slot[9 + 3 * k + c][lane], gD);
          }
        }
      }
      if (POSE) {
        float D = slot[21][lane];
        Geo g = CAM == 0 ? geometry(ik, (float)x, (float)yq, D) : geometry_lut(__ldg(lut_b + (size_t)yq * W + x), D);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float* P = k == 0 ? P0 : P1;
          float px = fmaf(P[0], g.c[0], fmaf(P[1], g.c[1], fmaf(P[2], g.c[2], P[3])));
          float py = fmaf(P[4], g.c[0], fmaf(P[5], g.c[1], fmaf(P[6], g.c[2], P[7])));
          float pz = fmaf(P[8], g.c[0], fmaf(P[9], g.c[1], fmaf(P[10], g.c[2], P[11])));
          float ax = fmaf(P[0], g.r[0], fmaf(P[1], g.r[1], P[2] * g.r[2]));
          float ay = fmaf(P[4], g.r[0], fmaf(P[5], g.r[1], P[6] * g.r[2]));
          float az = fmaf(P[8], g.r[0], fmaf(P[9], g.r[1], P[10] * g.r[2]));
          float gp[3];                       // d loss / d p (projected camera point)
          if (CAM == 0) {
            float rz = __frcp_rn(pz + 1e-7f);
            float u = px * rz, v = py * rz;
            gp[0] = gu[k] * rz; gp[1] = gv2[k] * rz; gp[2] = -(gu[k] * u + gv2[k] * v) * rz;
          } else {
            float u, v;
            Mei mei = mei_project(ik, px, py, pz, u, v);
            mei_vjp(ik, mei, px, py, pz, gu[k], gv2[k], gp);
          }
          gD = fmaf(gp[0], ax, fmaf(gp[1], ay, fmaf(gp[2], az, gD)));
          float hh[4] = {g.c[0], g.c[1], g.c[2], 1.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            gP[12 * k + j] = fmaf(gp[0], hh[j], gP[12 * k + j]);
            gP[12 * k + 4 + j] = fmaf(gp[1], hh[j], gP[12 * k + 4 + j]);
            gP[12 * k + 8 + j] = fmaf(gp[2], hh[j], gP[12 * k + 8 + j]);
          }
        }
      }
      float* gd = p.grad_depth + (size_t)b * p.hs * p.ws;
      if (full_res) {
        gd[(size_t)yq * W + x] = gD;
      } else {                               // transposed align_corners=True bilinear up-sample
        UpW wy = up_weights(yq, p.sy, p.hs);
        UpW wxx = up_weights(x, p.sx, p.ws);
        float a = gD * (1.f - wy.l), c2 = gD * wy.l;
        atomicAdd(gd + (size_t)wy.i0 * p.ws + wxx.i0, a * (1.f - wxx.l));
        atomicAdd(gd + (size_t)wy.i0 * p.ws + wxx.i1, a * wxx.l);
        atomicAdd(gd + (size_t)wy.i1 * p.ws + wxx.i0, c2 * (1.f - wxx.l));
        atomicAdd(gd + (size_t)wy.i1 * p.ws + wxx.i1, c2 * wxx.l);
      }
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) { A2[i] = A1[i] + h[i]; A1[i] = h[i]; }
    l1_prev[0] = l1[0]; l1_prev[1] = l1[1];
    valid_prev[0] = valid[0]; valid_prev[1] = valid[1];
    gf_prev[0] = gf[0]; gf_prev[1] = gf[1];
  }
  if (POSE) {
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      float v = warp_sum(gP[i]);
      if (lane == 0) atomicAdd(p.grad_P + (size_t)b * 24 + i, v);
    }
  }
  if (p.accum != nullptr) {
    const double n = warp_sum((double)acc_num);
    if (lane == 0) atomicAdd(p.accum, n);
  }
}

// ------------------------------------------------------------------------------------------------
// fused forward+backward, "frame pair" decomposition (round 2; replaces loss_bwd_kernel<0, CAM> in training steps).
//
// A CTA is TWO warps working on the same 30-column strip / row chunk: warp k owns source frame k -- its projection, gather,
// the 15 SSIM moments of (target, warped frame k), the SSIM partials and the 9-channel adjoint box filter.  Per row the two
// warps exchange one number per pixel (their photometric loss) through shared memory to agree on the arg-min of
// [identity(+1), identity(-1), reprojection(+1), reprojection(-1)] (monodepth2_decoder.py:261-263).  Halving the per-thread
// state (rolling 3x3 sums of 15 + 9 instead of 24 + 18 quantities) takes the kernel from 168 to < 100 registers, i.e. from 12 to
// 20 resident warps per SM: the old kernel was issue / latency bound at 44 % issue utilisation (profiles/r1_summary.md C.3).
//
// Halo of ONE pixel instead of two: SSIM partials w(p) are evaluated only at the pixels the CTA owns; the adjoint box filter
// G(q) = sum_{p in N(q), p owned} w(p) is evaluated on the owned region grown by one pixel, multiplied by the LOCAL derivative
// d pred(q) / d D and ADDED (red.global.add.f32) into the gradient at the pixel the lane actually read: reflect(q).  Partial
// sums of neighbouring CTAs meet in memory; ReflectionPad2d's fold-back (monodepth_utils.py:189) needs no special case, the halo
// lane at x = -1 holds pixel 1 and adds there.  The projection is folded per sample, p = D * (P[:, :3] K^-1 (x, y, 1)) + P[:, 3]:
// 6 FMAs per pixel and frame instead of 21 (the reference multiplies inv_K first, monodepth_utils.py:139-141,155-159; the
// coordinates differ by rounding, ~1e-7 relative).
// ------------------------------------------------------------------------------------------------
#ifndef LOSS_PAIR_OCC
#define LOSS_PAIR_OCC 8
#endif
#ifndef LOSS_PAIR_PIPELINE
#define LOSS_PAIR_PIPELINE 0
#endif
constexpr int kPairCols = 30;

// horizontal 3-tap sums of the 15 moments of (target, one warped frame): St, Stt, Sx, Sxx, Sxt per channel
__device__ __forceinline__ void hsum15(const float (&t)[3], const float (&x)[3], float (&h)[15]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float tl = __shfl_up_sync(0xffffffffu, t[c], 1), tr = __shfl_down_sync(0xffffffffu, t[c], 1);
    const float xl = __shfl_up_sync(0xffffffffu, x[c], 1), xr = __shfl_down_sync(0xffffffffu, x[c], 1);
    h[c] = tl + t[c] + tr;
    h[3 + c] = fmaf(tl, tl, fmaf(t[c], t[c], tr * tr));
    h[6 + c] = xl + x[c] + xr;
    h[9 + c] = fmaf(xl, xl, fmaf(x[c], x[c], xr * xr));
    h[12 + c] = fmaf(xl, tl, fmaf(x[c], t[c], xr * tr));
  }
}

// One row of inputs of the pair kernel, requested one loop iteration before it is consumed (software pipeline: every global
// load of an iteration is issued in ONE phase and has a whole iteration of arithmetic to land; the first version issued and
// consumed four dependent load groups per row and sat at 5.4 long-scoreboard stalls per issue, profiles/r2_loss_pair.md).
struct PairRow {
  float4 nw, ne, sw, se, tq;      // bilinear corners of the source frame (RGBX), target pixel
  float fx, fy, du, dv;           // bilinear fractions, d(u, v) / d D (zeroed where grid_sample's border clamp blocks the gradient)
  float mv;                       // MEI camera: validity of the ray table at the nearest source pixel (1 otherwise)
  int sel;                        // which corner is the nearest source pixel (bit 0 east, bit 1 south), 4 = out of bounds
  float id, nz;                   // centre-pixel inputs of the NEXT iteration: identity term, tie-break noise (or the motion mask)
};

template <int CAM>
__global__ void __launch_bounds__(64, LOSS_PAIR_OCC) loss_pair_kernel(LossParams p) {
  __shared__ float s_cam[2][20];                // per frame: M = P[:, :3] K^-1 (9), t = P[:, 3] (3), MEI intrinsics (7)
  __shared__ float s_ring[2][4][9][32];         // per frame: a ring of four rows (three live) of (target 3, warped 3, d warped / d D 3)
  __shared__ float s_ex[2][2][2][32];           // [row parity][frame][photometric | identity][lane]
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, W = p.W, HW = H * W;
  int item = blockIdx.x;
  const int strip = item % p.n_strips; item /= p.n_strips;
  const int chunk = item % p.n_chunks;
  const int b = item / p.n_chunks;
  const int y_begin = chunk * p.rows_per_item, y_end = min(y_begin + p.rows_per_item, H);
  const int x = strip * kPairCols + lane - 1;
  const int xr = reflect_idx(x, W);
  const bool own_col = lane >= 1 && lane <= kPairCols && x < W;
  const bool live_col = x <= W;                 // further right than the reflected halo column: nothing to add
  const bool overlap = (p.flags & FSNET_FLAG_OVERLAP_MASK) != 0;
  const bool use_ident = (p.flags & FSNET_FLAG_MOTION_MASK) == 0;
  const bool full_res = (p.hs == H && p.ws == W);

  {
    const float* cb = p.cam + ((size_t)b * 2 + k) * 21;
    if (lane < 9) {
      const int i = lane / 3, j = lane - 3 * i;
      float m;
      if (CAM == 0) m = fmaf(__ldg(cb + 9 + 4 * i), __ldg(cb + j), fmaf(__ldg(cb + 10 + 4 * i), __ldg(cb + 3 + j), __ldg(cb + 11 + 4 * i) * __ldg(cb + 6 + j)));
      else m = __ldg(cb + 9 + 4 * i + j);
      s_cam[k][lane] = m;
    } else if (lane < 12) {
      s_cam[k][lane] = __ldg(cb + 9 + 4 * (lane - 9) + 3);
    } else if (lane < 19) {
      s_cam[k][lane] = __ldg(cb + lane - 12);
    }
  }
  __syncwarp();
  const float* cm = s_cam[k];
  const float* in = s_cam[k] + 12;
  // pinhole: q = M (x, y, 1) = qc + y * M[:, 1]
  float qc0 = 0.f, qc1 = 0.f, qc2 = 0.f;
  if (CAM == 0) {
    qc0 = fmaf(cm[0], (float)xr, cm[2]); qc1 = fmaf(cm[3], (float)xr, cm[5]); qc2 = fmaf(cm[6], (float)xr, cm[8]);
  }
  const float xm = (float)(W - 1), ym = (float)(H - 1);

  const float4* tg4 = p.packed + (size_t)b * HW;
  const float4* sk4 = p.packed + ((size_t)(1 + k) * p.B + b) * HW;
  const float* depth = p.depth + (size_t)b * p.hs * p.ws;
  const void* mask_b = p.mask ? mask_dtype_ptr_add(p.mask, p.mask_dtype, (size_t)b * HW) : nullptr;
  const float* ident_k = use_ident ? p.ident + ((size_t)b * 2 + k) * HW : nullptr;
  const float* noise_k = (use_ident && p.noise) ? p.noise + ((size_t)b * 2 + k) * HW : nullptr;
  const float* motion_b = use_ident ? nullptr : p.motion + (size_t)b * HW;
  const float4* lut_b = CAM == 1 ? p.lut + (size_t)(p.lut_idx ? __ldg(p.lut_idx + b) : 0) * HW : nullptr;
  const UpW wx = up_weights(xr, p.sx, p.ws);
  float* gd = p.grad_depth + (size_t)b * p.hs * p.ws;
  const float gbase = (float)((double)__ldg(p.gout) / (__ldg(p.accum_in + 1) + 1e-6));

  // depth (and MEI ray) of a row: raw loads now, blended when consumed
  struct DepthRaw { float d00, d01, d10, d11, ly; float4 ray; };
  auto request_depth = [&](int yr) {
    DepthRaw r;
    r.d01 = r.d10 = r.d11 = r.ly = 0.f;
    r.ray = make_float4(0.f, 0.f, 0.f, 0.f);
    if (full_res) {
      r.d00 = __ldg(depth + yr * W + xr);
    } else {
      const UpW wy = up_weights(yr, p.sy, p.hs);
      const float* r0 = depth + wy.i0 * p.ws;
      const float* r1 = depth + wy.i1 * p.ws;
      r.d00 = __ldg(r0 + wx.i0); r.d01 = __ldg(r0 + wx.i1); r.d10 = __ldg(r1 + wx.i0); r.d11 = __ldg(r1 + wx.i1); r.ly = wy.l;
    }
    if (CAM == 1) r.ray = __ldg(lut_b + yr * W + xr);
    return r;
  };
  auto blend_depth = [&](const DepthRaw& r) {
    if (full_res) return r.d00;
    const float top = (1.f - wx.l) * r.d00 + wx.l * r.d01, bot = (1.f - wx.l) * r.d10 + wx.l * r.d11;
    return (1.f - r.ly) * top + r.ly * bot;
  };
  // project the pixel of raw row yy with depth D and issue every load that row needs; centre inputs of row yy - 1 ride along
  auto request_row = [&](int yy, float D, const float4& ray, bool want_centre) {
    PairRow r;
    const int yr = reflect_idx(yy, H);
    r.tq = __ldg(tg4 + yr * W + xr);
    float q0, q1, q2;
    if (CAM == 0) {
      const float fy_ = (float)yr;
      q0 = fmaf(cm[1], fy_, qc0); q1 = fmaf(cm[4], fy_, qc1); q2 = fmaf(cm[7], fy_, qc2);
    } else {
      q0 = fmaf(cm[0], ray.x, fmaf(cm[1], ray.y, cm[2] * ray.z));
      q1 = fmaf(cm[3], ray.x, fmaf(cm[4], ray.y, cm[5] * ray.z));
      q2 = fmaf(cm[6], ray.x, fmaf(cm[7], ray.y, cm[8] * ray.z));
    }
    const float px = fmaf(D, q0, cm[9]), py = fmaf(D, q1, cm[10]), pz = fmaf(D, q2, cm[11]);
    float ix, iy, du, dv;
    if (CAM == 0) {
      const float rz = __fdividef(1.f, pz + 1e-7f);
      ix = px * rz; iy = py * rz;
      du = (q0 - ix * q2) * rz; dv = (q1 - iy * q2) * rz;
    } else {
      const Mei mei = mei_project(in, px, py, pz, ix, iy);
      mei_jvp(in, mei, px, py, pz, q0, q1, q2, du, dv);
    }
    const float ixc = fminf(fmaxf(ix, 0.f), xm), iyc = fminf(fmaxf(iy, 0.f), ym);   // padding_mode='border'
    const float x0f = floorf(ixc), y0f = floorf(iyc);
    r.fx = ixc - x0f; r.fy = iyc - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int dx = x0 < W - 1 ? 1 : 0, dy = y0 < H - 1 ? W : 0;
    const float4* p00 = sk4 + y0 * W + x0;
    r.nw = __ldg(p00); r.ne = __ldg(p00 + dx); r.sw = __ldg(p00 + dy); r.se = __ldg(p00 + dy + dx);
    // overlap mask: nearest sample of patched_mask, zeros padding, "== 1" (monodepth2_decoder.py:110-116).  The nearest pixel
    // (round half to even, like grid_sample) is one of the four corners just requested; their 4th component holds the mask.
    // sel: bit 0 = east column, bit 1 = south row, 4 = out of bounds (mask reads as 0)
    r.sel = 4;
    r.mv = 1.f;
    if (overlap) {
      const float xn = rintf(ix), yn = rintf(iy);
      const bool inb = (xn >= 0.f) && (xn <= xm) && (yn >= 0.f) && (yn <= ym);
      r.sel = inb ? ((xn != x0f ? 1 : 0) | (yn != y0f ? 2 : 0)) : 4;
      if (CAM == 1 && inb) r.mv = __ldg(reinterpret_cast<const float*>(lut_b + (int)yn * W + (int)xn) + 3);
    }
    // clip_coordinates_set_grad: no coordinate gradient on or outside the border
    r.du = (ix > 0.f && ix < xm) ? du : 0.f;
    r.dv = (iy > 0.f && iy < ym) ? dv : 0.f;
    r.id = 0.f; r.nz = 0.f;
    if (want_centre) {                                       // centre pixel of the iteration that consumes this row: (x, yy - 1)
      const int pix = (yy - 1) * W + x;
      if (use_ident) {
        r.id = __ldg(ident_k + pix);
        if (noise_k) r.nz = __ldg(noise_k + pix);
      } else {
        r.nz = __ldg(motion_b + pix);                        // motion mask rides in nz
      }
    }
    return r;
  };

  // ---- the row walk ---------------------------------------------------------------------------------------------------
  // Iteration yy: (A) moments of raw row yy, (B) SSIM + arg-min at centre row yy - 1, (C) adjoint + chain at row yy - 2.
  // Rows y_begin-1, y_begin run A only, rows y_begin+1 .. y_end all three, two more iterations C only: three code paths
  // instantiated from one generic lambda, so the steady-state loop carries no phase predicates.  Vertical 3-row sums are
  // kept as PAIR sums (hP = h(yy-2) + h(yy-1)): S = hP + h(yy), then hP = h(yy-1) + h(yy) -- with the loop unrolled by two the
  // roles of the two row buffers alternate and nothing is ever moved between registers (the rotation of 2 x 15 + 2 x 9
  // registers per row was 84 of the 686 instructions of an iteration, profiles/r2_loss_pair_ncu.json).
  float hX[15], hY[15], hP[15], gX[9], gY[9], gP[9];
#pragma unroll
  for (int i = 0; i < 15; ++i) { hX[i] = 0.f; hY[i] = 0.f; hP[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 9; ++i) { gX[i] = 0.f; gY[i] = 0.f; gP[i] = 0.f; }
  float l1_prev = 0.f, gf_prev = 0.f, acc_num = 0.f, c_m_prev = 1.f;
  bool valid_prev = true;
  const int y_first = y_begin - 1;
  DepthRaw dnext = request_depth(reflect_idx(y_first, H));
  using T_ = std::true_type;
  using F_ = std::false_type;

  auto iter = [&](auto fA, auto fB, auto fC, auto fPar, int yy) {
    constexpr bool rowA = decltype(fA)::value, rowB = decltype(fB)::value, rowC = decltype(fC)::value;
    constexpr int par = decltype(fPar)::value;
    float(&hcur)[15] = par ? hY : hX;            // this row's horizontal sums land here; the other buffer holds the previous row's
    float(&hprev)[15] = par ? hX : hY;
    float(&gcur)[9] = par ? gY : gX;
    float(&gprev)[9] = par ? gX : gY;
    const bool centre = rowB && own_col;
    float c_id = 0.f, c_gate = 1.f, c_m_next = 1.f;
    const float c_m = c_m_prev;                          // centre row yy - 1: target mask, L1 and validity from one iteration ago
    const float l1_c = l1_prev;
    const bool valid_c = valid_prev;
    float l1 = 0.f;
    bool valid = true;
    float ph = 0.f, da[3], db[3], dc[3];
    if (rowA) {
      // ---- load phase: this row's gathers (its depth was requested one iteration ago), then the next row's depth ---
      const PairRow cur = request_row(yy, blend_depth(dnext), dnext.ray, centre);
      dnext = request_depth(reflect_idx(min(yy + 1, y_end), H));
      // ---- A: bilinear blend of the corners ------------------------------------------------------------------------
      // (every component of the 128-bit pixels is used: a dead 4th component lets the register allocator reuse that register
      // while the load is in flight, and the write-after-write hazard exposes the full memory latency behind every gather)
      if (use_ident) c_id = fmaf(cur.nz, 1e-5f, cur.id); else c_gate = 1.f - cur.nz;
      c_m_next = cur.tq.w;                                   // patched mask of the pixel that is the centre one iteration later
      {
        const float mw = (cur.sel & 2) ? ((cur.sel & 1) ? cur.se.w : cur.sw.w) : ((cur.sel & 1) ? cur.ne.w : cur.nw.w);
        valid = !overlap || ((cur.sel < 4) && mw * cur.mv == 1.f);
      }
      const float tt[3] = {cur.tq.x, cur.tq.y, cur.tq.z};
      const float a_nw[3] = {cur.nw.x, cur.nw.y, cur.nw.z}, a_ne[3] = {cur.ne.x, cur.ne.y, cur.ne.z};
      const float a_sw[3] = {cur.sw.x, cur.sw.y, cur.sw.z}, a_se[3] = {cur.se.x, cur.se.y, cur.se.z};
      float pred[3];
      float(*slot)[32] = s_ring[k][yy & 3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float dxt = a_ne[c] - a_nw[c], dxb = a_se[c] - a_sw[c];
        const float top = fmaf(cur.fx, dxt, a_nw[c]), bot = fmaf(cur.fx, dxb, a_sw[c]);
        const float dyv = bot - top;
        pred[c] = fmaf(cur.fy, dyv, top);
        const float dix = fmaf(cur.fy, dxb - dxt, dxt);
        slot[c][lane] = tt[c];
        slot[3 + c][lane] = pred[c];
        slot[6 + c][lane] = fmaf(dix, cur.du, dyv * cur.dv);    // d pred_c / d D
        l1 += fabsf(tt[c] - pred[c]);
      }
      hsum15(tt, pred, hcur);
      // ---- B (first half): SSIM value and partials at the centre row from the three rows of horizontal sums ---
      if (rowB) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          s += ssim_sums_grad(hP[6 + c] + hcur[6 + c], hP[9 + c] + hcur[9 + c], hP[12 + c] + hcur[12 + c],
                              hP[c] + hcur[c], hP[3 + c] + hcur[3 + c], da[c], db[c], dc[c]);
        ph = fmaf(0.85f / 3.f, s, (0.15f / 3.f) * l1_c);
        if (overlap && !valid_c) ph = 100.f;
      }
#pragma unroll
      for (int i = 0; i < 15; ++i) hP[i] = hprev[i] + hcur[i];
      l1_prev = l1; valid_prev = valid; c_m_prev = c_m_next;
    }
    // ---- B (second half): arg-min over [identity(+1), identity(-1), reprojection(+1), reprojection(-1)] ----------
    float ws_ = 0.f, gf = 0.f;
    if (rowB) {
      s_ex[par][k][0][lane] = ph;
      s_ex[par][k][1][lane] = c_id;
      __syncthreads();
      const float ph_o = s_ex[par][1 - k][0][lane], id_o = s_ex[par][1 - k][1][lane];
      const float p0 = k == 0 ? ph : ph_o, p1 = k == 0 ? ph_o : ph;
      float best;
      int arg;
      if (use_ident) {
        const float i0 = k == 0 ? c_id : id_o, i1 = k == 0 ? id_o : c_id;
        best = i0; arg = 0;
        if (i1 < best) { best = i1; arg = 1; }
        if (p0 < best) { best = p0; arg = 2; }
        if (p1 < best) { best = p1; arg = 3; }
        arg -= 2;
      } else {
        arg = p1 < p0 ? 1 : 0;
        best = fminf(p0, p1);
      }
      if (centre) {
        if (k == 0 && p.accum != nullptr) acc_num = fmaf(best, c_m, acc_num);
        if (arg == k && (!overlap || valid_c)) {
          gf = gbase * c_m * c_gate;
          ws_ = (0.85f / 3.f) * gf;
        }
      }
    }
    // ---- C: adjoint of the 3x3 box filter over the OWNED centres, chain to the depth at row yq = yy - 2 ----------
    // (rows in which no lane of the warp won carry no gradient: the products are skipped)
    if (rowB && __any_sync(0xffffffffu, ws_ != 0.f)) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // (halo lanes 0 and 31 hold w = 0 and a shuffle from outside the warp returns the lane's own value: no edge case)
        const float w0 = ws_ * da[c], w1 = ws_ * db[c], w2 = ws_ * dc[c];
        gcur[c] = __shfl_up_sync(0xffffffffu, w0, 1) + w0 + __shfl_down_sync(0xffffffffu, w0, 1);
        gcur[3 + c] = __shfl_up_sync(0xffffffffu, w1, 1) + w1 + __shfl_down_sync(0xffffffffu, w1, 1);
        gcur[6 + c] = __shfl_up_sync(0xffffffffu, w2, 1) + w2 + __shfl_down_sync(0xffffffffu, w2, 1);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) gcur[i] = 0.f;
    }
    if (rowC) {
      const int yq = yy - 2;
      float(*slot)[32] = s_ring[k][yq & 3];
      float gD = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float t = slot[c][lane], pr = slot[3 + c][lane];
        const float d = t - pr;
        const float sgn = (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f);
        float gp = fmaf(2.f * pr, gP[3 + c] + gcur[3 + c], fmaf(t, gP[6 + c] + gcur[6 + c], gP[c] + gcur[c]));
        gp = fmaf(-(0.15f / 3.f) * gf_prev, sgn, gp);
        gD = fmaf(gp, slot[6 + c][lane], gD);
      }
      if (live_col && gD != 0.f) {
        const int yw = reflect_idx(yq, H);
        if (full_res) {
          atomicAdd(gd + yw * W + xr, gD);
        } else {                                 // transposed align_corners=True bilinear up-sample
          const UpW wy = up_weights(yw, p.sy, p.hs);
          const float a = gD * (1.f - wy.l), c2 = gD * wy.l;
          atomicAdd(gd + wy.i0 * p.ws + wx.i0, a * (1.f - wx.l));
          atomicAdd(gd + wy.i0 * p.ws + wx.i1, a * wx.l);
          atomicAdd(gd + wy.i1 * p.ws + wx.i0, c2 * (1.f - wx.l));
          atomicAdd(gd + wy.i1 * p.ws + wx.i1, c2 * wx.l);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) gP[i] = gprev[i] + gcur[i];
    gf_prev = gf;
  };
  using P0_ = std::integral_constant<int, 0>;
  using P1_ = std::integral_constant<int, 1>;
  iter(T_{}, F_{}, F_{}, P0_{}, y_first);
  iter(T_{}, F_{}, F_{}, P1_{}, y_first + 1);
  int yy = y_begin + 1;
#pragma unroll 1
  for (; yy + 1 <= y_end; yy += 2) {
    iter(T_{}, T_{}, T_{}, P0_{}, yy);
    iter(T_{}, T_{}, T_{}, P1_{}, yy + 1);
  }
  if (yy <= y_end) {                               // odd number of rows in the chunk: the buffers' roles stay alternating
    iter(T_{}, T_{}, T_{}, P0_{}, yy);
    iter(F_{}, F_{}, T_{}, P1_{}, yy + 1);
    iter(F_{}, F_{}, T_{}, P0_{}, yy + 2);
  } else {
    iter(F_{}, F_{}, T_{}, P0_{}, yy);
    iter(F_{}, F_{}, T_{}, P1_{}, yy + 1);
  }
  if (k == 0 && p.accum != nullptr) {
    const double n = warp_sum((double)acc_num);
    if (lane == 0) atomicAdd(p.accum, n);
  }
}

__global__ void camera_setup_kernel(const float* __restrict__ P2, const float* __restrict__ T0,
                                    const float* __restrict__ T1, int B, float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2) return;
  int b = i >> 1, f = i & 1;
  const float* K = P2 + (size_t)b * 12;      // rows of 4; K = P2[:3,:3]
  const float* T = (f == 0 ? T0 : T1) + (size_t)b * 16;
  float* o = cam + (size_t)i * 21;
  double k[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) k[r][c] = (double)K[r * 4 + c];
  double c00 = k[1][1] * k[2][2] - k[1][2] * k[2][1];
  double c01 = k[1][2] * k[2][0] - k[1][0] * k[2][2];
  double c02 = k[1][0] * k[2][1] - k[1][1] * k[2][0];
  double det = k[0][0] * c00 + k[0][1] * c01 + k[0][2] * c02;
  double id = 1.0 / det;
  o[0] = (float)(c00 * id);
  o[1] = (float)((k[0][2] * k[2][1] - k[0][1] * k[2][2]) * id);
  o[2] = (float)((k[0][1] * k[1][2] - k[0][2] * k[1][1]) * id);
  o[3] = (float)(c01 * id);
  o[4] = (float)((k[0][0] * k[2][2] - k[0][2] * k[2][0]) * id);
  o[5] = (float)((k[0][2] * k[1][0] - k[0][0] * k[1][2]) * id);
  o[6] = (float)(c02 * id);
  o[7] = (float)((k[0][1] * k[2][0] - k[0][0] * k[2][1]) * id);
  o[8] = (float)((k[0][0] * k[1][1] - k[0][1] * k[1][0]) * id);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int j = 0; j < 3; ++j) s = fmaf(K[r * 4 + j], T[j * 4 + c], s);   // K[r][3] = 0 in the 4x4 embedding
      o[9 + r * 4 + c] = s;
    }
}

// MEI camera block: [0..6] = gamma1, gamma2, u0, v0, xi, k1, k2 (P2 and calib_meta), [9..20] = T[:3, :4]
// (FishEyeDecoder multiplies the back-projected point by cam_T_cam directly, monodepth2_decoder.py:379-381).
__global__ void camera_setup_mei_kernel(const float* __restrict__ P2, const double* __restrict__ calib,
                                        const float* __restrict__ T0, const float* __restrict__ T1, int B,
                                        float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2) return;
  int b = i >> 1, f = i & 1;
  const float* K = P2 + (size_t)b * 12;
  const float* T = (f == 0 ? T0 : T1) + (size_t)b * 16;
  float* o = cam + (size_t)i * 21;
  o[0] = K[0]; o[1] = K[5]; o[2] = K[2]; o[3] = K[6];
  o[4] = (float)calib[b * 3 + 0]; o[5] = (float)calib[b * 3 + 1]; o[6] = (float)calib[b * 3 + 2];
  o[7] = 0.f; o[8] = 0.f;
  for (int j = 0; j < 12; ++j) o[9 + j] = T[j];
}

// ---- MEI ray table: MeiCameraProjection.image2cam's cached LUT (mei_fisheye_utils.py:139-170) -------------
// header [B,8] = the calibration a table slot was last built for (gamma1, gamma2, u0, v0, xi, k1, k2, dirty).
// plan: sample b shares the table of the first sample with an identical calibration (lut_idx[b]); a slot is
// rebuilt only when its calibration changed, so the steady-state cost per step is this one tiny launch plus
// an early-exit build launch -- no host synchronisation (the reference calls .item() per sample, :151-154).
__global__ void mei_lut_plan_kernel(const float* __restrict__ P2, const double* __restrict__ calib, int B,
                                    double* __restrict__ header, int* __restrict__ lut_idx) {
  __shared__ double s_cal[256][7];
  const int b = threadIdx.x;
  if (b < B) {
    const float* K = P2 + (size_t)b * 12;
    s_cal[b][0] = K[0]; s_cal[b][1] = K[5]; s_cal[b][2] = K[2]; s_cal[b][3] = K[6];
    s_cal[b][4] = calib[b * 3 + 0]; s_cal[b][5] = calib[b * 3 + 1]; s_cal[b][6] = calib[b * 3 + 2];
  }
  __syncthreads();
  if (b >= B) return;
  int first = b;
  for (int j = 0; j < b; ++j) {
    bool same = true;
    for (int k = 0; k < 7; ++k) same = same && (s_cal[j][k] == s_cal[b][k]);
    if (same) { first = j; break; }
  }
  lut_idx[b] = first;
  double* h = header + (size_t)b * 8;
  bool dirty = false;
  if (first == b) {
    for (int k = 0; k < 7; ++k) dirty = dirty || !(h[k] == s_cal[b][k]);
    for (int k = 0; k < 7; ++k) h[k] = s_cal[b][k];
  }
  h[7] = dirty ? 1.0 : 0.0;
}

__device__ __forceinline__ double mei_radial(double k1, double k2, double r1, double r0) {
  const double r2 = r0 * r0;
  return r0 - r1 / (1.0 + k1 * r2 + k2 * (r2 * r2));
}
__device__ __forceinline__ double mei_mirror(double r0, double xi, double Z) {
  return r0 * r0 - (1.0 - Z * Z) / ((xi + Z) * (xi + Z));
}
__global__ void mei_lut_build_kernel(const double* __restrict__ header, const int* __restrict__ lut_idx, int H, int W,
                                     float4* __restrict__ lut) {
  const int b = blockIdx.y;
  const double* h = header + (size_t)b * 8;
  if (lut_idx[b] != b || h[7] == 0.0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int yy = i / W, xx = i % W;
  const double k1 = h[5], k2 = h[6], xi = h[4];
  float X = __fdiv_rn((float)xx - (float)h[2], (float)h[0]);      // gamma / principal point are fp32 P2 entries
  float Y = __fdiv_rn((float)yy - (float)h[3], (float)h[1]);
  const double r1 = (double)__fsqrt_rn(__fadd_rn(__fmul_rn(X, X), __fmul_rn(Y, Y)));
  const double tol = 1e-6;
  // Newton with a finite-difference derivative (mei_fisheye_utils.py:70-79)
  double r0 = r1;
  for (int it = 0; it < 100; ++it) {
    const double f = mei_radial(k1, k2, r1, r0);
    if (fabs(f) < tol) break;
    const double df = (mei_radial(k1, k2, r1, r0 + tol) - f) / tol;
    r0 = r0 - f / df;
  }
  // bisection on Z in [0, 1] (:85-101)
  const double y0 = mei_mirror(r0, xi, 0.0), y1 = mei_mirror(r0, xi, 1.0);
  bool flag = !(y0 * y1 > 0.0);
  double z = -1.0;
  if (flag) {
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 100; ++it) {
      z = (lo + hi) / 2;
      const double f = mei_mirror(r0, xi, z);
      if (fabs(f) < tol) break;
      if (f * mei_mirror(r0, xi, lo) < 0.0) hi = z; else lo = z;
    }
  }
  float Z = (float)z;
  float m = flag ? 1.f : 0.f;
  if (Z < 0.05f) m = 0.f;
  if (m == 0.f) { X = -1.f; Y = -1.f; Z = -1.f; }
  const float zx = __fadd_rn(Z, (float)xi);
  lut[(size_t)b * H * W + i] = make_float4(__fmul_rn(X, zx), __fmul_rn(Y, zx), Z, m);
}

// FishEyeDecoder.get_prediction (monodepth2_decoder.py:415-420): depth = z of the back-projected ray
__global__ void mei_depth_kernel(const float* __restrict__ norm, const float4* __restrict__ lut, const int* __restrict__ lut_idx,
                                 int B, int HW, float* __restrict__ depth) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * HW) return;
  const int b = (int)(i / HW), px = (int)(i % HW);
  const int t = lut_idx ? lut_idx[b] : 0;
  depth[i] = __ldg(reinterpret_cast<const float*>(lut + (size_t)t * HW + px) + 2) * norm[i];
}

// sum of the patched mask (the normaliser of monodepth2_decoder.py:292); it does not depend on the depth, so the fused
// forward+backward launches get it up front.  out[1] = sum(mask) (or `count` when there is no mask); out[0] is left alone.
__global__ void __launch_bounds__(256) mask_sum_kernel(const void* __restrict__ mask, int mask_dtype, size_t n, double count,
                                                       double* __restrict__ out, int n_out, int out_stride) {
  double acc = 0.0;
  if (mask != nullptr) {
    float a = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += load_mask(mask, mask_dtype, i);
    acc = warp_sum((double)a);
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    acc = count;
  }
  if ((threadIdx.x & 31) == 0 && acc != 0.0)
    for (int k = 0; k < n_out; ++k) atomicAdd(out + (size_t)k * out_stride + 1, acc);
}

// Rows per warp-item: minimise (number of waves) x (rows + halo rows) given how many warps are resident
// (148 SMs x warps/SM allowed by the kernel's registers), so that no second, mostly empty wave is left.
int plan(LossParams& p, int cols_per_warp, int halo_rows, int warps_per_sm) {
  p.n_strips = ceil_div(p.W, cols_per_warp);
  const long cap = 148L * warps_per_sm;
  int best_rows = 8;
  double best_cost = 1e30;
  for (int rows = 8; rows <= 64; ++rows) {
    const long items = (long)p.B * p.n_strips * ceil_div(p.H, rows);
    const long waves = (items + cap - 1) / cap;
    // rows actually walked by the longest item, plus a small bias towards more (smaller) items for balance
    const double cost = (double)waves * (rows + halo_rows) * (1.0 + 0.002 * rows) * ((double)ceil_div(p.H, rows) * rows / p.H);
    if (cost < best_cost) { best_cost = cost; best_rows = rows; }
  }
  p.rows_per_item = best_rows;
  p.n_chunks = ceil_div(p.H, best_rows);
  p.sy = p.H > 1 ? (float)(p.hs - 1) / (float)(p.H - 1) : 0.f;
  p.sx = p.W > 1 ? (float)(p.ws - 1) / (float)(p.W - 1) : 0.f;
  return ceil_div(p.B * p.n_strips * p.n_chunks, kWarps);
}

// frame-pair kernel: one CTA per (sample, 30-column strip, row chunk).  The kernel is issue bound (71 % issue utilisation, ncu
// r2c7), so the cost of a plan is its instruction count: every chunk walks rows + 3 rows (two halo rows of gathers, one more of the
// adjoint) -> long chunks, as long as the CTAs still fill ~80 % of the 148 SMs x LOSS_PAIR_OCC resident slots.
int plan_pair(LossParams& p) {
  p.n_strips = ceil_div(p.W, kPairCols);
  const double cap = 148.0 * LOSS_PAIR_OCC;
  int best_rows = 8;
  double best_cost = 1e30;
  for (int rows = 8; rows <= 96; ++rows) {
    const int chunks = ceil_div(p.H, rows);
    const double items = (double)p.B * p.n_strips * chunks;
    const double walked = (double)chunks * (rows + 3) / p.H;                 // rows walked per image row
    const double fill = items >= 0.8 * cap ? 1.0 : 0.8 * cap / items;          // too few CTAs: latency is no longer hidden
    const double cost = walked * fill;
    if (cost < best_cost - 1e-9) { best_cost = cost; best_rows = rows; }
  }
  p.rows_per_item = best_rows;
  p.n_chunks = ceil_div(p.H, best_rows);
  p.sy = p.H > 1 ? (float)(p.hs - 1) / (float)(p.H - 1) : 0.f;
  p.sx = p.W > 1 ? (float)(p.ws - 1) / (float)(p.W - 1) : 0.f;
  return p.B * p.n_strips * p.n_chunks;
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

extern "C" int fsnet_camera_setup(const float* P2, const float* T0, const float* T1, int B, float* cam, void* stream) {
  FSNET_REQUIRE(P2 && T0 && T1 && cam && B > 0, "fsnet_camera_setup: bad arguments");
  camera_setup_kernel<<<ceil_div(B * 2, 64), 64, 0, (cudaStream_t)stream>>>(P2, T0, T1, B, cam);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_identity_photometric_masked(const float* tgt, const float* src0, const float* src1, const void* mask, int mask_dtype,
                                                 int B, int H, int W, float* ident, float* packed, void* stream) {
  FSNET_REQUIRE(tgt && src0 && src1 && ident && packed, "fsnet_identity_photometric_masked: null pointer");
  FSNET_REQUIRE(((uintptr_t)packed & 15) == 0, "fsnet_identity_photometric_masked: packed buffer must be 16-byte aligned");
  FSNET_REQUIRE(B > 0 && H >= 3 && W >= 3, "fsnet_identity_photometric_masked: need B>0, H>=3, W>=3 (got %d,%d,%d)", B, H, W);
  FSNET_REQUIRE((mask == nullptr) == (mask_dtype == FSNET_MASK_NONE) && mask_dtype >= 0 && mask_dtype <= 2,
                "fsnet_identity_photometric_masked: mask pointer / dtype mismatch");
  LossParams p = {};
  p.tgt = tgt; p.src0 = src0; p.src1 = src1; p.B = B; p.H = H; p.W = W; p.hs = H; p.ws = W;
  p.mask = mask; p.mask_dtype = mask_dtype; p.flags = FSNET_FLAG_PACKED_MASK;
  p.ident_out = ident; p.packed_out = reinterpret_cast<float4*>(packed);
  int blocks = plan(p, 30, 2, 16);
  loss_fwd_kernel<0, 0><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_identity_photometric(const float* tgt, const float* src0, const float* src1,
                                          int B, int H, int W, float* ident, float* packed, void* stream) {
  FSNET_REQUIRE(tgt && src0 && src1 && ident, "fsnet_identity_photometric: null pointer");
  FSNET_REQUIRE(packed == nullptr || ((uintptr_t)packed & 15) == 0, "fsnet_identity_photometric: packed buffer must be 16-byte aligned");
  FSNET_REQUIRE(B > 0 && H >= 3 && W >= 3, "fsnet_identity_photometric: need B>0, H>=3, W>=3 (got %d,%d,%d)", B, H, W);
  LossParams p = {};
  p.tgt = tgt; p.src0 = src0; p.src1 = src1; p.B = B; p.H = H; p.W = W; p.hs = H; p.ws = W;
  p.ident_out = ident; p.packed_out = reinterpret_cast<float4*>(packed);
  int blocks = plan(p, 30, 2, 16);
  loss_fwd_kernel<0, 0><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

static int check_common(const float* depth_s, int hs, int ws, const float* packed,
                        const void* mask, int mask_dtype, const float* cam, const float* ident, const float* motion,
                        unsigned flags, int B, int H, int W) {
  FSNET_REQUIRE(depth_s && packed && cam, "fsnet_warp_ssim: null pointer");
  FSNET_REQUIRE(((uintptr_t)packed & 15) == 0, "fsnet_warp_ssim: packed images must be 16-byte aligned");
  FSNET_REQUIRE(B > 0 && H >= 3 && W >= 3 && hs >= 1 && ws >= 1 && hs <= H && ws <= W,
                "fsnet_warp_ssim: bad shape B=%d H=%d W=%d hs=%d ws=%d", B, H, W, hs, ws);
  FSNET_REQUIRE((mask == nullptr) == (mask_dtype == FSNET_MASK_NONE), "fsnet_warp_ssim: mask pointer / dtype mismatch");
  FSNET_REQUIRE(mask_dtype >= 0 && mask_dtype <= 2, "fsnet_warp_ssim: unknown mask dtype %d", mask_dtype);
  if (flags & FSNET_FLAG_MOTION_MASK) FSNET_REQUIRE(motion != nullptr, "fsnet_warp_ssim: motion-mask flag without a motion mask");
  else FSNET_REQUIRE(ident != nullptr, "fsnet_warp_ssim: identity terms required (no motion mask)");
  return FSNET_OK;
}

static int launch_fwd(int cam_model, const float* lut, const int* lut_idx,
                      const float* depth_s, int hs, int ws, const float* packed,
                      const void* mask, int mask_dtype, const float* cam,
                      const float* ident, const float* noise, const float* motion, unsigned flags,
                      int B, int H, int W, double* accum, uint8_t* sel, float* pred0, void* stream) {
  int rc = check_common(depth_s, hs, ws, packed, mask, mask_dtype, cam, ident, motion, flags, B, H, W);
  if (rc) return rc;
  FSNET_REQUIRE(accum, "fsnet_warp_ssim_fwd: null accumulator");
  LossParams p = {};
  p.depth = depth_s; p.hs = hs; p.ws = ws; p.packed = reinterpret_cast<const float4*>(packed);
  p.mask = mask; p.mask_dtype = mask_dtype; p.cam = cam; p.ident = ident; p.noise = noise; p.motion = motion;
  p.flags = flags; p.B = B; p.H = H; p.W = W; p.accum = accum; p.sel = sel; p.pred0 = pred0;
  p.lut = reinterpret_cast<const float4*>(lut); p.lut_idx = lut_idx;
  int blocks = plan(p, 30, 2, 16);
  if (cam_model == 0) loss_fwd_kernel<1, 0><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(p);
  else loss_fwd_kernel<1, 1><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

static int launch_bwd(int cam_model, const float* lut, const int* lut_idx,
                      const float* depth_s, int hs, int ws, const float* packed,
                      const void* mask, int mask_dtype, const float* cam,
                      const float* ident, const float* noise, const float* motion, unsigned flags,
                      int B, int H, int W, const double* accum, const float* gout,
                      float* grad_depth, float* grad_P, void* stream, double* accum_out = nullptr) {
  int rc = check_common(depth_s, hs, ws, packed, mask, mask_dtype, cam, ident, motion, flags, B, H, W);
  if (rc) return rc;
  FSNET_REQUIRE(accum && gout && grad_depth, "fsnet_warp_ssim_bwd: null pointer");
  LossParams p = {};
  p.depth = depth_s; p.hs = hs; p.ws = ws; p.packed = reinterpret_cast<const float4*>(packed);
  p.mask = mask; p.mask_dtype = mask_dtype; p.cam = cam; p.ident = ident; p.noise = noise; p.motion = motion;
  p.flags = flags; p.B = B; p.H = H; p.W = W; p.accum_in = accum; p.gout = gout;
  p.grad_depth = grad_depth; p.grad_P = grad_P; p.accum = accum_out;
  p.lut = reinterpret_cast<const float4*>(lut); p.lut_idx = lut_idx;
  cudaStream_t st = (cudaStream_t)stream;
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = getenv("FSNET_LOSS_PAIR"); pair_env = e ? atoi(e) : 1; }
  if (pair_env && grad_P == nullptr && accum_out != nullptr && (flags & FSNET_FLAG_PACKED_MASK)) {
    // fused forward+backward of a training step: the frame-pair kernel (every gradient write is an add: the caller zeroes grad_depth)
    const int items = plan_pair(p);
    if (cam_model == 0) loss_pair_kernel<0><<<items, 64, 0, st>>>(p);
    else loss_pair_kernel<1><<<items, 64, 0, st>>>(p);
    FSNET_LAUNCH_OK();
    return FSNET_OK;
  }
  int blocks = plan(p, 28, 4, 4 * LOSS_BWD_OCC);
  if (cam_model == 0) {
    if (grad_P) loss_bwd_kernel<1, 0><<<blocks, kWarps * 32, 0, st>>>(p);
    else loss_bwd_kernel<0, 0><<<blocks, kWarps * 32, 0, st>>>(p);
  } else {
    if (grad_P) loss_bwd_kernel<1, 1><<<blocks, kWarps * 32, 0, st>>>(p);
    else loss_bwd_kernel<0, 1><<<blocks, kWarps * 32, 0, st>>>(p);
  }
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_warp_ssim_fwd(const float* depth_s, int hs, int ws, const float* packed,
                                   const void* mask, int mask_dtype, const float* cam,
                                   const float* ident, const float* noise, const float* motion, unsigned flags,
                                   int B, int H, int W, double* accum, uint8_t* sel, float* pred0, void* stream) {
  return launch_fwd(0, nullptr, nullptr, depth_s, hs, ws, packed, mask, mask_dtype, cam, ident, noise, motion, flags,
                    B, H, W, accum, sel, pred0, stream);
}

extern "C" int fsnet_warp_ssim_bwd(const float* depth_s, int hs, int ws, const float* packed,
                                   const void* mask, int mask_dtype, const float* cam,
                                   const float* ident, const float* noise, const float* motion, unsigned flags,
                                   int B, int H, int W, const double* accum, const float* gout,
                                   float* grad_depth, float* grad_P, void* stream) {
  return launch_bwd(0, nullptr, nullptr, depth_s, hs, ws, packed, mask, mask_dtype, cam, ident, noise, motion, flags,
                    B, H, W, accum, gout, grad_depth, grad_P, stream);
}

extern "C" int fsnet_warp_ssim_mei_fwd(const float* lut, const int* lut_idx,
                                       const float* norm_s, int hs, int ws, const float* packed,
                                       const void* mask, int mask_dtype, const float* cam,
                                       const float* ident, const float* noise, const float* motion, unsigned flags,
                                       int B, int H, int W, double* accum, uint8_t* sel, float* pred0, void* stream) {
  FSNET_REQUIRE(lut && ((uintptr_t)lut & 15) == 0, "fsnet_warp_ssim_mei_fwd: ray table must be non-null and 16-byte aligned");
  return launch_fwd(1, lut, lut_idx, norm_s, hs, ws, packed, mask, mask_dtype, cam, ident, noise, motion, flags,
                    B, H, W, accum, sel, pred0, stream);
}

extern "C" int fsnet_warp_ssim_mei_bwd(const float* lut, const int* lut_idx,
                                       const float* norm_s, int hs, int ws, const float* packed,
                                       const void* mask, int mask_dtype, const float* cam,
                                       const float* ident, const float* noise, const float* motion, unsigned flags,
                                       int B, int H, int W, const double* accum, const float* gout,
                                       float* grad_norm, float* grad_T, void* stream) {
  FSNET_REQUIRE(lut && ((uintptr_t)lut & 15) == 0, "fsnet_warp_ssim_mei_bwd: ray table must be non-null and 16-byte aligned");
  return launch_bwd(1, lut, lut_idx, norm_s, hs, ws, packed, mask, mask_dtype, cam, ident, noise, motion, flags,
                    B, H, W, accum, gout, grad_norm, grad_T, stream);
}

extern "C" int fsnet_camera_setup_mei(const float* P2, const double* calib, const float* T0, const float* T1, int B,
                                      float* cam, void* stream) {
  FSNET_REQUIRE(P2 && calib && T0 && T1 && cam && B > 0, "fsnet_camera_setup_mei: bad arguments");
  camera_setup_mei_kernel<<<ceil_div(B * 2, 64), 64, 0, (cudaStream_t)stream>>>(P2, calib, T0, T1, B, cam);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_mei_lut(const float* P2, const double* calib, int B, int H, int W,
                             double* header, int* lut_idx, float* lut, void* stream) {
  FSNET_REQUIRE(P2 && calib && header && lut_idx && lut, "fsnet_mei_lut: null pointer");
  FSNET_REQUIRE(B > 0 && B <= 256 && H > 0 && W > 0, "fsnet_mei_lut: bad shape (B <= 256) B=%d H=%d W=%d", B, H, W);
  FSNET_REQUIRE(((uintptr_t)lut & 15) == 0, "fsnet_mei_lut: ray table must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  mei_lut_plan_kernel<<<1, 256, 0, st>>>(P2, calib, B, header, lut_idx);
  FSNET_LAUNCH_OK();
  dim3 grid(ceil_div(H * W, 128), B);
  mei_lut_build_kernel<<<grid, 128, 0, st>>>(header, lut_idx, H, W, reinterpret_cast<float4*>(lut));
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_mei_depth(const float* norm, const float* lut, const int* lut_idx, int B, int H, int W,
                               float* depth, void* stream) {
  FSNET_REQUIRE(norm && lut && depth && B > 0 && H > 0 && W > 0, "fsnet_mei_depth: bad arguments");
  const long n = (long)B * H * W;
  mei_depth_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(norm, reinterpret_cast<const float4*>(lut), lut_idx, B, H * W, depth);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_mask_sum(const void* mask, int mask_dtype, long long n, double* out, int n_out, int out_stride, void* stream) {
  FSNET_REQUIRE(out && n > 0 && n_out > 0 && out_stride >= 2, "fsnet_mask_sum: bad arguments");
  FSNET_REQUIRE((mask == nullptr) == (mask_dtype == FSNET_MASK_NONE) && mask_dtype >= 0 && mask_dtype <= 2, "fsnet_mask_sum: mask pointer / dtype mismatch");
  const int grid = mask ? (int)((n + 256 * 16 - 1) / (256 * 16) < 148 * 8 ? (n + 256 * 16 - 1) / (256 * 16) : 148 * 8) : 1;
  mask_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, mask_dtype, (size_t)n, (double)n, out, n_out, out_stride);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_warp_ssim_fwdbwd(const float* lut, const int* lut_idx,
                                      const float* depth_s, int hs, int ws, const float* packed,
                                      const void* mask, int mask_dtype, const float* cam,
                                      const float* ident, const float* noise, const float* motion, unsigned flags,
                                      int B, int H, int W, double* accum, const float* gout,
                                      float* grad_depth, float* grad_P, void* stream) {
  FSNET_REQUIRE(lut == nullptr || ((uintptr_t)lut & 15) == 0, "fsnet_warp_ssim_fwdbwd: ray table must be 16-byte aligned");
  return launch_bwd(lut ? 1 : 0, lut, lut_idx, depth_s, hs, ws, packed, mask, mask_dtype, cam, ident, noise, motion, flags,
                    B, H, W, accum, gout, grad_depth, grad_P, stream, accum);
}
