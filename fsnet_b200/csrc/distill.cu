// Distillation loss of the second training stage: value and unit gradients in one pass.
// Reference: MonoDepth2Decoder.compute_distill_loss, monodepth2_decoder.py:185-203 (scaled branch) with the sigmoid of
// MultiChannelDepthDecoderUncertain.forward (depth_encoder.py:190) folded in:
//     u = sigmoid(l);   loss = mean_i( |t_i - p_i| / u_i + log(u_i + 1e-5) )        (plain mean |t - p| without l)
// HBM bound: 12 B read + 8 B written per pixel, one launch per scale.
#include "common.cuh"

namespace fsnet {
namespace {

constexpr int kThreads = 256;

// One pixel: returns its loss term and writes d term / d p, d term / d l (both already divided by n) and u = sigmoid(l).
__device__ __forceinline__ float distill_term(float p, float t, bool has_u, float l, float inv_n, float& gp, float& gl, float& u) {
  const float d = t - p, e = fabsf(d);
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);          // d|d|/dd as autograd defines it (0 at 0)
  gp = -sgn * inv_n;
  gl = 0.f;
  u = 1.f;
  if (!has_u) return e;
  u = 1.f / (1.f + expf(-l));
  gp = gp / u;
  gl = (1.f / (u + 1e-5f) - e / (u * u)) * (u * (1.f - u)) * inv_n;
  return e / u + logf(u + 1e-5f);
}

__global__ void __launch_bounds__(kThreads) distill_loss_kernel(const float* __restrict__ pred, const float* __restrict__ teacher,
                                                               const float* __restrict__ ulogit, long long n, float inv_n,
                                                               double* __restrict__ out, float* __restrict__ grad_pred,
                                                               float* __restrict__ grad_ulogit, float* __restrict__ uncertain) {
  __shared__ double part[kThreads / 32];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    float gp, gl, u;
    acc += (double)distill_term(ldg(pred + i), ldg(teacher + i), ulogit != nullptr, ulogit != nullptr ? ldg(ulogit + i) : 0.f, inv_n,
                                gp, gl, u);
    if (grad_pred != nullptr) grad_pred[i] = gp;
    if (grad_ulogit != nullptr) grad_ulogit[i] = gl;
    if (uncertain != nullptr) uncertain[i] = u;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v * (double)inv_n);
  }
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

extern "C" int fsnet_distill_loss(const float* pred, const float* teacher, const float* ulogit, long long n, double* out,
                                  float* grad_pred, float* grad_ulogit, float* uncertain, void* stream) {
  FSNET_REQUIRE(pred && teacher && out, "fsnet_distill_loss: null pointer");
  FSNET_REQUIRE(n > 0, "fsnet_distill_loss: empty input (n = %lld)", n);
  FSNET_REQUIRE(ulogit || (!grad_ulogit && !uncertain), "fsnet_distill_loss: uncertainty outputs requested without ulogit");
  const long long want = (n + kThreads - 1) / kThreads;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);              // <= 8 resident blocks per SM, grid-stride beyond
  distill_loss_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(pred, teacher, ulogit, n, 1.f / (float)n, out, grad_pred,
                                                                  grad_ulogit, uncertain);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
