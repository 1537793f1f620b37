// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   D[128 pixels, BN out-channels] += A[128 pixels, KC in-channels of tap (r,s)] * W[BN, KC]^T
//
// * activations live in HBM as NHWC bf16 "planes": hi = bf16(x), lo = bf16(x - hi).  With NPROD = 3 the
//   kernel issues hi*hi + lo*hi + hi*lo, which carries ~16 mantissa bits through the tensor core (the
//   north_star parity of 1e-3 on disparity fails with plain bf16 and is marginal with TF32);
//   NPROD = 1 is the plain bf16 path used for gradients.
// * the A tile of a tap is ONE 4-D TMA box [1, TH, TW, KC] of the NHWC tensor at the shifted
//   coordinate: out-of-bounds rows/columns are zero-filled by the TMA unit (= zero padding); strided
//   convolutions use the tensor map's element strides; replicate padding reads a tensor whose
//   one-pixel ring has been materialised by the producer kernel.
// * 128B/64B/32B shared-memory swizzle chosen from KC = 64/32/16 channels per stage, K-major UMMA
//   descriptors, accumulators double-buffered in TMEM so the epilogue of tile i overlaps the main loop
//   of tile i+1 (persistent CTAs, one per SM).
// * epilogue: tcgen05.ld -> (+bias) -> fp32 NHWC store, plus per-channel sum / sum-of-squares of the
//   tile (train-mode BatchNorm statistics) reduced with a shuffle transpose and accumulated in fp64.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 =
// epilogue (TMEM lane quadrant = warp_id % 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace fsnet {
namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvParams {
  int N, H, W, Cin;            // logical input size
  int Ho, Wo, Cout;
  int KH, KW, stride, pad, dil, org;   // org = 1 when the tensor map covers a materialised padding ring
  int TH, TW;                  // output tile (TH*TW <= 128 pixels)
  int tiles_x, tiles_y, n_tiles, total_tiles;
  int BN;                      // output channels per tile (multiple of 16, <= 128)
  int KC, cchunks, kiters;     // channels per stage, Cin/KC, taps*cchunks
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes, tx_bytes;
  uint32_t tmem_cols;
  uint32_t sbo, layout_type;   // UMMA smem descriptor fields for this swizzle
  float* out;                  // [N,Ho,Wo,Cout] fp32
  const float* bias;           // [Cout] or null
  double* stats;               // [2*Cout] (sum, sumsq) or null
  int relu;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride byte offset: 8 rows
  d |= (uint64_t)1 << 46;                              // version
  d |= (uint64_t)(layout_type & 7) << 61;              // swizzle mode
  return d;
}

// 16-column transpose-reduce: lane (row) values v[0..15] -> column sums; lanes 2c and 2c+1 both end up
// with the sum of column c in v[0].
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bool up = lane & 16;
    float send = up ? v[j] : v[j + 8], keep = up ? v[j + 8] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bool up = lane & 8;
    float send = up ? v[j] : v[j + 4], keep = up ? v[j + 4] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    bool up = lane & 4;
    float send = up ? v[j] : v[j + 2], keep = up ? v[j + 2] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    bool up = lane & 2;
    float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];      // column index = ((lane>>4)&1)*8 + ((lane>>3)&1)*4 + ((lane>>2)&1)*2 + ((lane>>1)&1)
}

template <int NPROD>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  // dynamic smem: [stages * stage_bytes | 2*Cout floats of statistics]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* s_stats = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (NPROD == 3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.Cout; i += kThreads) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
        int tx = mt % p.tiles_x; int rest = mt / p.tiles_x;
        int ty = rest % p.tiles_y; int img = rest / p.tiles_y;
        const int x_base = tx * p.TW * p.stride - p.pad + p.org;
        const int y_base = ty * p.TH * p.stride - p.pad + p.org;
        for (int kit = 0; kit < p.kiters; ++kit) {
          int tap = kit / p.cchunks, cc = kit - tap * p.cchunks;
          int r = tap / p.KW, s = tap - r * p.KW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          mbar_expect_tx(&full_bar[stage], p.tx_bytes);
          tma_load_4d(st, &map_a_hi, &full_bar[stage], cc * p.KC, x_base + s * p.dil, y_base + r * p.dil, img);
          tma_load_2d(st + p.a_bytes, &map_b_hi, &full_bar[stage], tap * p.Cin + cc * p.KC, nt * p.BN);
          if (NPROD == 3) {
            tma_load_4d(st + p.a_bytes + p.b_bytes, &map_a_lo, &full_bar[stage], cc * p.KC, x_base + s * p.dil, y_base + r * p.dil, img);
            tma_load_2d(st + 2 * p.a_bytes + p.b_bytes, &map_b_lo, &full_bar[stage], tap * p.Cin + cc * p.KC, nt * p.BN);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const int ksteps = p.KC / 16;
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int kit = 0; kit < p.kiters; ++kit) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint64_t a_hi = make_desc(st, p.sbo, p.layout_type);
          const uint64_t b_hi = make_desc(st + p.a_bytes, p.sbo, p.layout_type);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(d_tmem, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (kit | k) != 0);
          if (NPROD == 3) {
            const uint64_t a_lo = make_desc(st + p.a_bytes + p.b_bytes, p.sbo, p.layout_type);
            const uint64_t b_lo = make_desc(st + 2 * p.a_bytes + p.b_bytes, p.sbo, p.layout_type);
            for (int k = 0; k < ksteps; ++k) umma_bf16(d_tmem, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1);
            for (int k = 0; k < ksteps; ++k) umma_bf16(d_tmem, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1);
          }
          umma_commit(&empty_bar[stage]);           // frees the smem stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);               // accumulator complete
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                         // TMEM lane quadrant of this warp
    const int m = q * 32 + lane;                    // row of the tile = pixel
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      int tx = mt % p.tiles_x; int rest = mt / p.tiles_x;
      int ty = rest % p.tiles_y; int img = rest / p.tiles_y;
      const int py = m / p.TW, px = m - py * p.TW;
      const int oy = ty * p.TH + py, ox = tx * p.TW + px;
      const bool valid = (m < p.TH * p.TW) && oy < p.Ho && ox < p.Wo;
      float* orow = p.out + (((size_t)img * p.Ho + oy) * p.Wo + ox) * p.Cout + nt * p.BN;
      mbar_wait(&tmem_full[acc], ((uint32_t)it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(taddr + c0, raw);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = valid ? __uint_as_float(raw[j]) : 0.f;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + nt * p.BN + c0 + j);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (p.stats) {
          float sq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = valid ? v[j] : 0.f; sq[j] = v[j] * v[j]; }
          float s1 = colsum16(v, lane), s2 = colsum16(sq, lane);
          if ((lane & 1) == 0) {
            int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            atomicAdd(&s_stats[nt * p.BN + c0 + col], s1);
            atomicAdd(&s_stats[p.Cout + nt * p.BN + c0 + col], s2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps
      for (int i = threadIdx.x - 64; i < 2 * p.Cout; i += 128) {
        float s = s_stats[i];
        if (s != 0.f) atomicAdd(p.stats + i, (double)s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
    fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int pick_kc(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }

void pick_tile(int Ho, int Wo, int& TH, int& TW) {
  double best = -1;
  TH = 1; TW = 1;
  for (int tw = 1; tw <= 128 && tw <= Wo; ++tw) {
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th > 256) th = 256;
    double fill = (double)(th * tw) / 128.0;
    double cover = ((double)Ho / ((double)ceil_div(Ho, th) * th)) * ((double)Wo / ((double)ceil_div(Wo, tw) * tw));
    double score = fill * cover + 1e-6 * tw;       // prefer wide tiles on ties (longer contiguous rows)
    if (score > best) { best = score; TH = th; TW = tw; }
  }
}

}  // namespace
}  // namespace fsnet

using namespace fsnet;

// ---------------------------------------------------------------------------------------------------
// fsnet_conv_fwd -- see include/fsnet_b200.h
// ---------------------------------------------------------------------------------------------------
extern "C" int fsnet_conv_fwd(const void* in_hi, const void* in_lo, int N, int H, int W, int Cin, int pitch_w, int pitch_h,
                              int ring, const void* w_hi, const void* w_lo, int Cout, int KH, int KW, int stride, int pad,
                              int use_ring, int nprod, const float* bias, int relu, float* out, double* stats, void* stream) {
  FSNET_REQUIRE(in_hi && w_hi && out, "fsnet_conv_fwd: null pointer");
  FSNET_REQUIRE(nprod == 1 || (nprod == 3 && in_lo && w_lo), "fsnet_conv_fwd: nprod must be 1 or 3 (3 needs the lo planes)");
  FSNET_REQUIRE(N > 0 && H > 0 && W > 0 && Cin % 16 == 0 && Cout % 16 == 0, "fsnet_conv_fwd: channels must be multiples of 16 (Cin=%d Cout=%d)", Cin, Cout);
  FSNET_REQUIRE(stride == 1 || stride == 2, "fsnet_conv_fwd: stride %d unsupported", stride);
  FSNET_REQUIRE(!use_ring || (ring >= pad), "fsnet_conv_fwd: replicate padding needs a materialised ring >= pad");
  EncodeTiledFn enc = encode_fn();
  FSNET_REQUIRE(enc != nullptr, "fsnet_conv_fwd: cuTensorMapEncodeTiled not available from the driver");

  ConvParams p = {};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.dil = 1;
  p.Ho = (H + 2 * pad - KH) / stride + 1;
  p.Wo = (W + 2 * pad - KW) / stride + 1;
  p.org = use_ring ? ring : 0;
  pick_tile(p.Ho, p.Wo, p.TH, p.TW);
  p.tiles_x = ceil_div(p.Wo, p.TW); p.tiles_y = ceil_div(p.Ho, p.TH);
  p.BN = Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16));
  if (Cout <= 128) p.BN = Cout;
  FSNET_REQUIRE(p.BN % 16 == 0 && p.BN <= 128 && Cout % p.BN == 0, "fsnet_conv_fwd: cannot tile Cout=%d", Cout);
  p.n_tiles = Cout / p.BN;
  p.total_tiles = N * p.tiles_x * p.tiles_y * p.n_tiles;
  p.KC = pick_kc(Cin); p.cchunks = Cin / p.KC; p.kiters = KH * KW * p.cchunks;
  p.a_bytes = 128u * p.KC * 2;
  p.b_bytes = ((uint32_t)p.BN * p.KC * 2 + 1023u) & ~1023u;
  p.stage_bytes = (nprod == 3 ? 2u : 1u) * (p.a_bytes + p.b_bytes);
  p.tx_bytes = (nprod == 3 ? 2u : 1u) * ((uint32_t)p.TH * p.TW * p.KC * 2 + (uint32_t)p.BN * p.KC * 2);
  const uint32_t stats_bytes = stats ? 2u * Cout * 4u : 0u;
  int stages = (int)((200u * 1024u - stats_bytes) / p.stage_bytes);
  p.stages = stages > kMaxStages ? kMaxStages : stages;
  FSNET_REQUIRE(p.stages >= 2, "fsnet_conv_fwd: tile does not fit shared memory");
  uint32_t cols = 32;
  while (cols < 2u * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  p.sbo = 8u * p.KC * 2;
  p.layout_type = p.KC == 64 ? 2u : (p.KC == 32 ? 4u : 6u);
  p.out = out; p.bias = bias; p.stats = stats; p.relu = relu;

  // activation maps: dims (C, W', H', N) over the logical image (zero padding through OOB fill) or over
  // the ringed tensor (materialised replicate padding)
  const int Wm = use_ring ? W + 2 * ring : W, Hm = use_ring ? H + 2 * ring : H;
  const char* base_hi = (const char*)in_hi;
  const char* base_lo = (const char*)in_lo;
  if (!use_ring) {            // caller passes the pointer to the ring origin; step to the interior
    size_t off = ((size_t)ring * pitch_w + ring) * Cin * 2;
    base_hi += off;
    if (base_lo) base_lo += off;
  }
  cuuint64_t adim[4] = {(cuuint64_t)Cin, (cuuint64_t)Wm, (cuuint64_t)Hm, (cuuint64_t)N};
  cuuint64_t astr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)pitch_w * Cin * 2, (cuuint64_t)pitch_h * pitch_w * Cin * 2};
  cuuint32_t abox[4] = {(cuuint32_t)p.KC, (cuuint32_t)(p.TW * stride), (cuuint32_t)(p.TH * stride), 1};
  cuuint32_t aes[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  FSNET_REQUIRE(abox[1] <= 256 && abox[2] <= 256, "fsnet_conv_fwd: TMA box too large");
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  CUresult r = enc(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base_hi, adim, astr, abox, aes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv_fwd: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  ma_lo = ma_hi;
  if (nprod == 3) {
    r = enc(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base_lo, adim, astr, abox, aes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv_fwd: cuTensorMapEncodeTiled(A lo) failed with %d", (int)r);
  }
  const int Ktot = KH * KW * Cin;
  cuuint64_t bdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t bstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t bbox[2] = {(cuuint32_t)p.KC, (cuuint32_t)p.BN};
  cuuint32_t bes[2] = {1, 1};
  r = enc(&mb_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_hi, bdim, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
          swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv_fwd: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  mb_lo = mb_hi;
  if (nprod == 3) {
    r = enc(&mb_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_lo, bdim, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv_fwd: cuTensorMapEncodeTiled(B lo) failed with %d", (int)r);
  }

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  const size_t smem = (size_t)p.stages * p.stage_bytes + stats_bytes + 1024;
  auto kern = nprod == 3 ? conv_tc_kernel<3> : conv_tc_kernel<1>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[nprod == 3]) {
    FSNET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_set[nprod == 3] = true;
  }
  kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
