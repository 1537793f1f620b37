// Fused multi-tensor gradient-norm clip + Adam: two launches replace clip_grad_norm_ + Adam.step()
// (base_training_hooks.py:46-49, optimizers.py:8 of the reference; ~330 parameter tensors x (norm, scale,
// 4 Adam element-wise ops) in stock PyTorch).  HBM bound: 28 bytes per parameter element (read g, p, m, v;
// write p, m, v).  The tensor table lives in device memory so that the step can be captured in a CUDA graph.
#include "common.cuh"

namespace fsnet {
namespace {

constexpr int kChunk = 4096;                 // elements per block-iteration

__device__ __forceinline__ int find_tensor(const fsnet_adam_tensor* __restrict__ tab, int n, long long chunk) {
  int lo = 0, hi = n - 1;                    // last tensor whose first chunk <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk_start <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const fsnet_adam_tensor* __restrict__ tab, int n_tensors, long long n_chunks,
                                                         double* __restrict__ sumsq) {
  float acc = 0.f;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int t = find_tensor(tab, n_tensors, ch);
    const fsnet_adam_tensor e = tab[t];
    const long long base = (ch - e.chunk_start) * kChunk;
    const long long end = min(base + (long long)kChunk, e.n);
    const float* g = e.g;
    if ((((uintptr_t)g) & 15) == 0) {
      for (long long i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
        if (i + 4 <= end) {
          const float4 v = *reinterpret_cast<const float4*>(g + i);
          acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
        } else {
          for (long long j = i; j < end; ++j) acc = fmaf(g[j], g[j], acc);
        }
      }
    } else {
      for (long long i = base + threadIdx.x; i < end; i += 256) acc = fmaf(g[i], g[i], acc);
    }
  }
  __shared__ double s_part[8];
  double d = warp_sum((double)acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += s_part[i];
    if (s != 0.0) atomicAdd(sumsq, s);
  }
}

// hyper [8] fp64 on the device: lr, beta1, beta2, eps, weight_decay, max_norm (<= 0: no clipping), step (incremented here), unused
// (fp64 so that the bias corrections 1 - beta^t match the Python doubles of torch.optim.Adam)
__global__ void adam_tick_kernel(double* __restrict__ hyper) { hyper[6] += 1.0; }

__global__ void __launch_bounds__(256) adam_step_kernel(const fsnet_adam_tensor* __restrict__ tab, int n_tensors, long long n_chunks,
                                                        const double* __restrict__ sumsq, const double* __restrict__ hyper) {
  const float b1 = (float)hyper[1], b2 = (float)hyper[2], eps = (float)hyper[3], wd = (float)hyper[4], max_norm = (float)hyper[5];
  float coef = 1.f;
  if (max_norm > 0.f) {                       // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max = 1)
    const float norm = (float)sqrt(*sumsq);
    coef = fminf(max_norm / (norm + 1e-6f), 1.f);
  }
  const double bc1 = 1.0 - pow(hyper[1], hyper[6]), bc2 = 1.0 - pow(hyper[2], hyper[6]);
  const float step_size = (float)(hyper[0] / bc1), rs_bc2 = (float)(1.0 / sqrt(bc2));
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int t = find_tensor(tab, n_tensors, ch);
    const fsnet_adam_tensor e = tab[t];
    const long long base = (ch - e.chunk_start) * kChunk;
    const long long end = min(base + (long long)kChunk, e.n);
    for (long long i = base + threadIdx.x; i < end; i += 256) {
      float g = e.g[i] * coef;
      float p = e.p[i];
      if (wd != 0.f) g = fmaf(wd, p, g);
      const float m = fmaf(b1, e.m[i], (1.f - b1) * g);
      const float v = fmaf(b2, e.v[i], (1.f - b2) * g * g);
      e.m[i] = m; e.v[i] = v;
      e.p[i] = p - step_size * (m / (sqrtf(v) * rs_bc2 + eps));
    }
  }
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

extern "C" int fsnet_grad_sumsq(const fsnet_adam_tensor* table, int n_tensors, long long n_chunks, double* sumsq, void* stream) {
  FSNET_REQUIRE(table && sumsq && n_tensors > 0 && n_chunks > 0, "fsnet_grad_sumsq: bad arguments");
  const int grid = (int)(n_chunks < 148 * 8 ? n_chunks : 148 * 8);
  grad_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table, n_tensors, n_chunks, sumsq);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_adam_step(const fsnet_adam_tensor* table, int n_tensors, long long n_chunks, const double* sumsq, double* hyper,
                               void* stream) {
  FSNET_REQUIRE(table && sumsq && hyper && n_tensors > 0 && n_chunks > 0, "fsnet_adam_step: bad arguments");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper);
  FSNET_LAUNCH_OK();
  const int grid = (int)(n_chunks < 148 * 8 ? n_chunks : 148 * 8);
  adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table, n_tensors, n_chunks, sumsq, hyper);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
