// Host stand-ins for the two tensor-core entry points of the C ABI (include/fsnet_b200.h: fsnet_conv, fsnet_conv_wgrad), written
// from the ABI's contract, NOT an emulation of the tcgen05 / TMA kernels (those can only be checked on a B200: tests/test_conv_gpu.py).
// They let the executor (fsnet_b200/engine.py) and every other kernel of a training step run end to end in the CPU suite.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fsnet_b200.h"

// the host-side planning code of conv_tc.cu itself (argument checks, tiling, pipeline depth, tensor-map encodings), device code
// stripped: emulate.py::translate_plan
extern "C" int fsnet_conv_plan(const fsnet_view* in, int use_ring, const void* w_hi, const void* w_lo, int Cout, int KH, int KW, int stride,
                               int pad, int nprod, const float* bias, int relu, const fsnet_view* out, int accumulate, double* stats,
                               void* stream);
extern "C" int fsnet_conv_wgrad_plan(const fsnet_view* x, int use_ring, const fsnet_view* dy, int KH, int KW, int stride, int pad, float* acc,
                                     void* stream);
extern "C" { long long fsnet_emulated_plans = 0; }          // number of planner calls that passed (read by the tests)

namespace {
inline bool plan_only() {
  const char* e = getenv("FSNET_EMULATE_PLAN_ONLY");
  return e && e[0] == '1';
}
inline float bf(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
struct PlaneView {
  const uint16_t* hi;
  const uint16_t* lo;
  int n, h, w, c, ring, ct, coff, ph, pw;
  explicit PlaneView(const fsnet_view* v)
      : hi((const uint16_t*)v->ptr), n(v->n), h(v->h), w(v->w), c(v->c), ring(v->ring), ct(v->c_total), coff(v->c_off),
        ph(v->h + 2 * v->ring), pw(v->w + 2 * v->ring) {
    lo = hi + (size_t)n * ph * pw * ct;
  }
  // element offset of pixel (img, y, x) -- y / x relative to the interior, may reach into the ring
  size_t at(int img, int y, int x) const { return (((size_t)img * ph + y + ring) * pw + x + ring) * ct + coff; }
};
}  // namespace

extern "C" int fsnet_conv(const fsnet_view* in, int use_ring, const void* w_hi_, const void* w_lo_, int Cout, int KH, int KW, int stride,
                          int pad, int nprod, const float* bias, int relu, const fsnet_view* out, int accumulate, double* stats, void*) {
  const int rc = fsnet_conv_plan(in, use_ring, w_hi_, w_lo_, Cout, KH, KW, stride, pad, nprod, bias, relu, out, accumulate, stats, nullptr);
  if (rc != FSNET_OK) return rc;
  fsnet_emulated_plans++;
  if (plan_only()) return FSNET_OK;
  if (!in || !in->ptr || !w_hi_ || !out || !out->ptr) return FSNET_ERR_INVALID;
  if (!(nprod == 1 || (nprod == 3 && w_lo_))) return FSNET_ERR_INVALID;
  if ((in->c % 16 && in->c != 8) || Cout % 16 || (use_ring && in->ring < pad)) return FSNET_ERR_INVALID;   // 8: the network stem
  const PlaneView x(in);
  const int Cin = x.c, Ho = (x.h + 2 * pad - KH) / stride + 1, Wo = (x.w + 2 * pad - KW) / stride + 1;
  if (out->n != x.n || out->h != Ho || out->w != Wo || out->c != Cout) return FSNET_ERR_INVALID;
  const uint16_t* w_hi = (const uint16_t*)w_hi_;
  const uint16_t* w_lo = (const uint16_t*)w_lo_;
  float* o = (float*)out->ptr;
  const int oph = out->h + 2 * out->ring, opw = out->w + 2 * out->ring;
  std::vector<float> xh(Cin), xl(Cin);
  std::vector<double> acc(Cout);
  for (int img = 0; img < x.n; ++img)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox) {
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int kh = 0; kh < KH; ++kh)
          for (int kw = 0; kw < KW; ++kw) {
            const int iy = oy * stride + kh - pad, ix = ox * stride + kw - pad;
            const bool interior = iy >= 0 && iy < x.h && ix >= 0 && ix < x.w;
            if (!interior && !use_ring) continue;                      // zero padding
            const size_t base = x.at(img, iy, ix);                     // use_ring: the materialised ring is read as data
            for (int ci = 0; ci < Cin; ++ci) { xh[ci] = bf(x.hi[base + ci]); xl[ci] = nprod == 3 ? bf(x.lo[base + ci]) : 0.f; }
            for (int co = 0; co < Cout; ++co) {
              const size_t wb = (((size_t)co * KH + kh) * KW + kw) * Cin;
              double s = 0.0;
              if (nprod == 3)
                for (int ci = 0; ci < Cin; ++ci) {
                  const float wh = bf(w_hi[wb + ci]);
                  s += (double)xh[ci] * wh + (double)xl[ci] * wh + (double)xh[ci] * bf(w_lo[wb + ci]);
                }
              else
                for (int ci = 0; ci < Cin; ++ci) s += (double)xh[ci] * bf(w_hi[wb + ci]);
              acc[co] += s;
            }
          }
        float* dst = o + (((size_t)img * oph + oy + out->ring) * opw + ox + out->ring) * out->c_total + out->c_off;
        for (int co = 0; co < Cout; ++co) {
          float v = (float)acc[co];
          if (bias) v += bias[co];
          if (relu) v = fmaxf(v, 0.f);
          if (stats) { stats[co] += v; stats[Cout + co] += (double)v * v; }
          dst[co] = accumulate ? dst[co] + v : v;
        }
      }
  return FSNET_OK;
}

extern "C" int fsnet_conv_wgrad(const fsnet_view* x_, int use_ring, const fsnet_view* dy_, int KH, int KW, int stride, int pad, float* acc,
                                void*) {
  const int rc = fsnet_conv_wgrad_plan(x_, use_ring, dy_, KH, KW, stride, pad, acc, nullptr);
  if (rc != FSNET_OK) return rc;
  fsnet_emulated_plans++;
  if (plan_only()) return FSNET_OK;
  if (!x_ || !dy_ || !x_->ptr || !dy_->ptr || !acc) return FSNET_ERR_INVALID;
  const PlaneView x(x_), dy(dy_);
  const int Cin = x.c, Cout = dy.c, Ho = dy.h, Wo = dy.w;
  if ((x.h + 2 * pad - KH) / stride + 1 != Ho || (x.w + 2 * pad - KW) / stride + 1 != Wo) return FSNET_ERR_INVALID;
  if (use_ring && x.ring < pad) return FSNET_ERR_INVALID;
  std::vector<double> sum((size_t)Cout * KH * KW * Cin, 0.0);
  std::vector<float> xv(Cin);
  for (int img = 0; img < x.n; ++img)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox) {
        const size_t db = dy.at(img, oy, ox);
        for (int kh = 0; kh < KH; ++kh)
          for (int kw = 0; kw < KW; ++kw) {
            const int iy = oy * stride + kh - pad, ix = ox * stride + kw - pad;
            const bool interior = iy >= 0 && iy < x.h && ix >= 0 && ix < x.w;
            if (!interior && !use_ring) continue;
            const size_t xb = x.at(img, iy, ix);
            for (int ci = 0; ci < Cin; ++ci) xv[ci] = bf(x.hi[xb + ci]);
            for (int co = 0; co < Cout; ++co) {
              const float g = bf(dy.hi[db + co]);
              if (g == 0.f) continue;
              double* s = &sum[(((size_t)co * KH + kh) * KW + kw) * Cin];
              for (int ci = 0; ci < Cin; ++ci) s[ci] += (double)g * xv[ci];
            }
          }
      }
  for (size_t i = 0; i < sum.size(); ++i) acc[i] += (float)sum[i];
  return FSNET_OK;
}
