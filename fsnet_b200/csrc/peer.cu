// SyncBatchNorm statistics over NVLink / NVSwitch PEER MEMORY instead of NCCL (reference: torch.nn.SyncBatchNorm as wrapped by
// scripts/train.py:100-102; one all_gather per BatchNorm layer and direction).
//
// A data-parallel step of the depth network exchanges 2C fp64 numbers 68 times (ResNet-18 + decoder: 34 layers, forward statistics
// and the two backward sums).  Each exchange sits on the critical path between a convolution and the kernel that normalises its
// output, and an NCCL all-reduce of 1-8 KB costs ~15-25 us of launch + protocol latency at 8 GPUs: ~1.4 ms of a 9.6 ms step
// (SCALE_r01).  Here the exchange is a ONE-SHOT all-reduce written into the kernel that consumes the statistics:
// every rank stores its 2C numbers straight into a slot of every peer's buffer (the buffers are mapped into all ranks' address
// spaces: symmetric memory, torch.distributed._symmetric_memory -> cuMemMap over NVLink), raises a flag per peer
// (st.release.sys), waits for the world's flags in its OWN flag array (ld.acquire.sys), sums the slots in rank order (so every
// rank gets bit-identical statistics) and goes on to compute scale / shift / running statistics -- one launch, ~3-4 us, no host
// involvement, capturable into the step's CUDA graph.
//
// Protocol per slot (a slot belongs to one layer and direction; it is reused every step):
//   seq   = ++local_counter[slot]                       (all ranks execute the same sequence of exchanges, so counters agree)
//   data  : peer r's buffer [slot][my_rank][0..n) <- my values          (plain stores to mapped peer memory)
//   fence : __threadfence_system(); __syncthreads();
//   flag  : peer r's flags [slot][my_rank] <- seq                       (release, system scope)
//   wait  : my flags [slot][r] >= seq for every r                        (acquire, system scope)
//   sum   : my buffer [slot][r][i] over r, fixed order
// A peer can not overwrite a slot before this rank has read it: its next write to the same slot happens a whole step later, and
// between the two lie exchanges of other slots that need this rank's flags, which this rank only raises after the kernel that read
// the slot has finished (stream order).
#include "common.cuh"

namespace fsnet {
namespace {

struct PeerArgs {
  void* const* bufs;      // device array [world]: every rank's data buffer (fp64)
  void* const* flags;     // device array [world]: every rank's flag buffer (u32)
  int rank, world;
  long long slot_off;     // in doubles: this slot's [world][n] region inside every data buffer
  int flag_off;           // this slot's [world] flags inside every flag buffer
  unsigned* seq;          // this slot's local use counter
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// whole CTA: data[0..n) := sum over ranks of data[0..n)
__device__ void peer_allreduce_sum(double* data, int n, const PeerArgs& a) {
  __shared__ unsigned s_seq;
  if (threadIdx.x == 0) s_seq = ++(*a.seq);
  __syncthreads();
  const unsigned seq = s_seq;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = data[i];
    for (int r = 0; r < a.world; ++r) reinterpret_cast<double*>(a.bufs[r])[a.slot_off + (long long)a.rank * n + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    st_release_sys(reinterpret_cast<unsigned*>(a.flags[threadIdx.x]) + a.flag_off + a.rank, seq);
    const unsigned* mine = reinterpret_cast<const unsigned*>(a.flags[a.rank]) + a.flag_off + threadIdx.x;
    while ((int)(ld_acquire_sys(mine) - seq) < 0) {}
  }
  __syncthreads();
  const double* own = reinterpret_cast<const double*>(a.bufs[a.rank]) + a.slot_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += __ldcv(own + (long long)r * n + i);      // written by peers: not through a stale L1 line
    data[i] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(double* data, int n, PeerArgs a) { peer_allreduce_sum(data, n, a); }

// bn_finalize (act_tc.cu) with the cross-rank sum of the statistics in front: same arithmetic as nn.SyncBatchNorm / nn.BatchNorm2d
// on the global batch (count = elements per channel over ALL ranks).
__global__ void __launch_bounds__(256) bn_finalize_sync_kernel(double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                                               float* __restrict__ running_mean, float* __restrict__ running_var,
                                                               long long* __restrict__ num_batches, float momentum, float eps, int C,
                                                               float* __restrict__ scale_shift, float* __restrict__ mean_invstd, PeerArgs a) {
  peer_allreduce_sum(stats, 2 * C, a);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0) var = 0;
    const float mean = (float)m;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const float bias = conv_bias ? conv_bias[c] : 0.f;
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + bias);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale_shift[c] = g * invstd;
    scale_shift[C + c] = b - mean * g * invstd;
    if (mean_invstd) { mean_invstd[c] = mean; mean_invstd[C + c] = invstd; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) stats[c] = 0.0;        // the convolution epilogue accumulates into it
  if (threadIdx.x == 0 && num_batches) *num_batches += 1;
}

int check_peer(const fsnet_peer* pe, int n, const char* who) {
  FSNET_REQUIRE(pe && pe->bufs && pe->flags && pe->seq, "%s: null peer descriptor", who);
  FSNET_REQUIRE(pe->world >= 2 && pe->world <= 64 && pe->rank >= 0 && pe->rank < pe->world, "%s: bad rank %d / world %d", who, pe->rank, pe->world);
  FSNET_REQUIRE(pe->slot_off >= 0 && pe->flag_off >= 0 && n > 0, "%s: bad slot", who);
  return FSNET_OK;
}

PeerArgs to_args(const fsnet_peer* pe) {
  PeerArgs a;
  a.bufs = reinterpret_cast<void* const*>(pe->bufs); a.flags = reinterpret_cast<void* const*>(pe->flags);
  a.rank = pe->rank; a.world = pe->world; a.slot_off = pe->slot_off; a.flag_off = pe->flag_off; a.seq = reinterpret_cast<unsigned*>(pe->seq);
  return a;
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

extern "C" int fsnet_peer_allreduce_f64(double* data, int n, const fsnet_peer* peer, void* stream) {
  FSNET_REQUIRE(data, "fsnet_peer_allreduce_f64: null data");
  int rc = check_peer(peer, n, "fsnet_peer_allreduce_f64");
  if (rc) return rc;
  peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(data, n, to_args(peer));
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

extern "C" int fsnet_bn_finalize_sync(double* stats, double count, const float* gamma, const float* beta, const float* conv_bias,
                                      float* running_mean, float* running_var, long long* num_batches, float momentum, float eps,
                                      int C, float* scale_shift, float* mean_invstd, const fsnet_peer* peer, void* stream) {
  FSNET_REQUIRE(stats && scale_shift && C > 0, "fsnet_bn_finalize_sync: bad arguments");
  int rc = check_peer(peer, 2 * C, "fsnet_bn_finalize_sync");
  if (rc) return rc;
  bn_finalize_sync_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, conv_bias, running_mean, running_var, num_batches,
                                                              momentum, eps, C, scale_shift, mean_invstd, to_args(peer));
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
