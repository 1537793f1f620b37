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
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 =
// epilogue (TMEM lane quadrant = warp_id % 4, two warps per quadrant on alternate 16-column groups).
//
// Three kernels in this file: conv_tc_kernel (the per-tap scheme above: stem, stride-2 and 1x1 layers, small maps; optional
// cluster multicast), conv_halo_kernel (every 3x3 / stride-1 layer: one halo tile per chunk feeds all nine taps through shifted
// operand descriptors, weights in their own ring or resident) and wgrad_tc_kernel (weight gradient, MN-major operands).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace fsnet {
namespace {

constexpr int kThreads = 320;          // warp 0 TMA producer, 1 MMA issuer, 2-9 epilogue
constexpr int kMaxStages = 8;

struct ConvParams {
  int N, H, W, Cin;            // logical input size
  int Ho, Wo, Cout;
  int KH, KW, stride, pad, dil, org;   // org = 1 when the tensor map covers a materialised padding ring
  int TH, TW;                  // output tile (TH*TW <= 128 pixels)
  int tiles_x, tiles_y, n_tiles, total_tiles;
  int BN;                      // output channels per tile (multiple of 16, <= 128)
  int KC, cchunks, kiters;     // channels per stage, Cin/KC, taps*cchunks
  int fold;                    // x-taps folded into K (cchunks = ceil(KW*Cin / 64)).  1: one stage per (kernel row, chunk);
                               // 2: one stage per chunk holding the (TH+KH-1)-row halo tile once, kernel rows = smem row offsets
  uint32_t b_each, tap_bytes;  // fold 2: bytes of one kernel row's weight box, smem offset between kernel rows of A (TW * 128)
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes, tx_bytes;
  uint32_t tmem_cols;
  uint32_t sbo, layout_type;   // UMMA smem descriptor fields for this swizzle
  float* out;                  // fp32 NHWC view: element (n,y,x,c) at ((n*out_ph + y+out_ring)*out_pw + x+out_ring)*out_ct + out_coff + c
  int out_pw, out_ph, out_ring, out_ct, out_coff, accumulate;
  const float* bias;           // [Cout] or null
  double* stats;               // [2*Cout] (sum, sumsq) or null
  int relu;
  // thread-block cluster of cm x cn CTAs working on cm pixel tiles x cn channel tiles at once: the A tile (pixels) is multicast to the
  // cn CTAs of a cluster row, the B tile (weights) to the cm CTAs of a cluster column (L2 -> SM delivery bounds these kernels)
  int cm, cn, super_n, total_super, m_tiles;
  // halo path (conv_halo_kernel): 3x3 / stride 1, 64-channel K chunks.  The (TH+2) x (TW+2) halo of a 128-pixel tile is loaded ONCE per
  // chunk and plane; tap (r, s) is the same shared-memory tile read from a start address shifted by whole pixels.
  // halo = 1: 16 rows x 8 pixels (box [64, 10, 18]);  halo = 2: 8 rows x 16 pixels, y fastest in shared memory (box [64 ch, 10 rows, 18 cols])
  int halo, a_stages, b_stages, b_resident;
  uint32_t a_plane_bytes, a_stage_bytes, b_plane_bytes, b_stage_bytes, a_tx, b_tx;
  int dbg;                     // diagnostics (FSNET_CONV_DBG, tools/bench_conv.py): 1 skip the accumulate read, 2 skip the stores, 4 skip the statistics
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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

// elected (warp-converged) variants, see umma_bf16_elect
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variants: the box lands at the same shared-memory offset of every CTA in `mask` and completes on each one's own barrier
__device__ __forceinline__ void tma_load_4d_mc_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_elect(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;\n}\n"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
// Warp-converged variants: the whole (converged) warp executes the statement, one elected lane issues the instruction.
// Keeping the issuing code free of thread-dependent control flow lets the compiler hold descriptors in uniform
// registers instead of the ELECT / R2UR.BROADCAST / BRA.U.ANY sequence it emits inside `if (lane == 0)`.
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred pe, pa;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 pa, %4, 0;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// One 64-element K chunk of a stage in ONE statement: a single election, then the 4 (NPROD = 1) or 12 (NPROD = 3: hi*hi, lo*hi,
// hi*lo) K-steps back to back, descriptor advances (32 B = 2 units per K-step) done in PTX.  The tensor pipe's queue is shallow:
// whatever the issuing warp executes between two tcgen05.mma (an election per MMA, 64-bit descriptor assembly per tap: ~400 cycles
// per stage in the first version) showed up as idle pipe time on top of the MMAs (profiles/r2_conv_halo.md, DBG ablation).
template <int NPROD, int KS = 4>
__device__ __forceinline__ void umma_chunk_elect(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                 uint32_t idesc, uint32_t accumulate_first) {
  // KS = K-steps of 16 elements in the chunk (4: 64-channel rows, 2: 32, 1: 16); products hi*hi, then lo*hi, then hi*lo
#define FSNET_MMA(A, B, P) "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], " A ", " B ", %5, " P ";\n"
  if (NPROD == 3) {
    if (KS == 4) {
      asm volatile(
          "{\n"
          ".reg .pred pe, pz, pt;\n"
          ".reg .b64 ah1, ah2, ah3, al1, al2, al3, bh1, bh2, bh3, bl1, bl2, bl3;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.ne.b32 pz, %6, 0;\n"
          "setp.eq.b32 pt, %6, %6;\n"
          "add.u64 ah1, %1, 2;\n add.u64 ah2, %1, 4;\n add.u64 ah3, %1, 6;\n"
          "add.u64 al1, %2, 2;\n add.u64 al2, %2, 4;\n add.u64 al3, %2, 6;\n"
          "add.u64 bh1, %3, 2;\n add.u64 bh2, %3, 4;\n add.u64 bh3, %3, 6;\n"
          "add.u64 bl1, %4, 2;\n add.u64 bl2, %4, 4;\n add.u64 bl3, %4, 6;\n"
          FSNET_MMA("%1", "%3", "pz") FSNET_MMA("ah1", "bh1", "pt") FSNET_MMA("ah2", "bh2", "pt") FSNET_MMA("ah3", "bh3", "pt")
          FSNET_MMA("%2", "%3", "pt") FSNET_MMA("al1", "bh1", "pt") FSNET_MMA("al2", "bh2", "pt") FSNET_MMA("al3", "bh3", "pt")
          FSNET_MMA("%1", "%4", "pt") FSNET_MMA("ah1", "bl1", "pt") FSNET_MMA("ah2", "bl2", "pt") FSNET_MMA("ah3", "bl3", "pt")
          "}\n" ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first) : "memory");
    } else if (KS == 2) {
      asm volatile(
          "{\n"
          ".reg .pred pe, pz, pt;\n"
          ".reg .b64 ah1, al1, bh1, bl1;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.ne.b32 pz, %6, 0;\n"
          "setp.eq.b32 pt, %6, %6;\n"
          "add.u64 ah1, %1, 2;\n add.u64 al1, %2, 2;\n add.u64 bh1, %3, 2;\n add.u64 bl1, %4, 2;\n"
          FSNET_MMA("%1", "%3", "pz") FSNET_MMA("ah1", "bh1", "pt")
          FSNET_MMA("%2", "%3", "pt") FSNET_MMA("al1", "bh1", "pt")
          FSNET_MMA("%1", "%4", "pt") FSNET_MMA("ah1", "bl1", "pt")
          "}\n" ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first) : "memory");
    } else {
      asm volatile(
          "{\n"
          ".reg .pred pe, pz, pt;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.ne.b32 pz, %6, 0;\n"
          "setp.eq.b32 pt, %6, %6;\n"
          FSNET_MMA("%1", "%3", "pz") FSNET_MMA("%2", "%3", "pt") FSNET_MMA("%1", "%4", "pt")
          "}\n" ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first) : "memory");
    }
  } else {
    // (operand numbering as above: %2 / %4, the lo planes, are unused)
    if (KS == 4) {
      asm volatile(
          "{\n"
          ".reg .pred pe, pz, pt;\n"
          ".reg .b64 ah1, ah2, ah3, bh1, bh2, bh3;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.ne.b32 pz, %6, 0;\n"
          "setp.eq.b32 pt, %6, %6;\n"
          "add.u64 ah1, %1, 2;\n add.u64 ah2, %1, 4;\n add.u64 ah3, %1, 6;\n"
          "add.u64 bh1, %3, 2;\n add.u64 bh2, %3, 4;\n add.u64 bh3, %3, 6;\n"
          FSNET_MMA("%1", "%3", "pz") FSNET_MMA("ah1", "bh1", "pt") FSNET_MMA("ah2", "bh2", "pt") FSNET_MMA("ah3", "bh3", "pt")
          "}\n" ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first) : "memory");
    } else if (KS == 2) {
      asm volatile(
          "{\n"
          ".reg .pred pe, pz, pt;\n"
          ".reg .b64 ah1, bh1;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.ne.b32 pz, %6, 0;\n"
          "setp.eq.b32 pt, %6, %6;\n"
          "add.u64 ah1, %1, 2;\n add.u64 bh1, %3, 2;\n"
          FSNET_MMA("%1", "%3", "pz") FSNET_MMA("ah1", "bh1", "pt")
          "}\n" ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first) : "memory");
    } else {
      asm volatile(
          "{\n"
          ".reg .pred pe, pz;\n"
          "elect.sync _|pe, 0xffffffff;\n"
          "setp.ne.b32 pz, %6, 0;\n"
          FSNET_MMA("%1", "%3", "pz")
          "}\n" ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first) : "memory");
    }
  }
#undef FSNET_MMA
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrival on the barrier at this offset in every CTA of `mask` (frees a multicast stage in all its producers)
__device__ __forceinline__ void umma_commit_mc_elect(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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

// Epilogue of the convolution kernels, run by EIGHT warps (2..9): accumulator -> (+bias, ReLU) -> fp32 NHWC store (optionally
// accumulating) + per-channel sum / sum-of-squares of the tile.  A warp reads the TMEM lane quadrant warp % 4 (its 32 pixels);
// the two warps of a quadrant take alternate 16-column groups.  With four warps and shared-memory fp64 atomics for the statistics
// the epilogue (~12 K cycles per 128 x 64 tile, 38 % of it in the atomics' compare-and-swap loops) was slower than the tile's
// MMAs (~5.4 K cycles) and bounded the halo kernel (ncu: profiles/r2_conv_halo.md).
// Statistics: every (quadrant, column) has ONE owning lane that adds its per-tile column sum (shuffle tree over 32 pixels, fp32,
// deterministic) into a private fp64 slot -- no atomics, no order dependence inside the CTA -- and the slots are reduced over
// the quadrants and flushed to global memory (fp64 atomics) when the CTA moves to another channel tile and at the end.
constexpr int kEpiWarps = 8;
__device__ __forceinline__ void epilogue_flush_stats(const ConvParams& p, double* s_slots, int nt, int tid_e) {
  asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight epilogue warps: all slots of this channel tile are written
  if (tid_e < 2 * p.BN) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { s += s_slots[q * 2 * p.BN + tid_e]; s_slots[q * 2 * p.BN + tid_e] = 0.0; }
    const int ch = tid_e < p.BN ? nt * p.BN + tid_e : p.Cout + nt * p.BN + (tid_e - p.BN);
    if (s != 0.0) atomicAdd(p.stats + ch, s);
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

__device__ __forceinline__ void epilogue_warps(const ConvParams& p, int warp, int lane, uint32_t tmem_base,
                                               uint64_t* tmem_full, uint64_t* tmem_empty, double* s_slots) {
  const int q = warp & 3;                         // TMEM lane quadrant of this warp
  const int half = (warp - 2) >> 2;               // which of the quadrant's two warps: 16-column groups half, half + 2, ...
  const int tid_e = (warp - 2) * 32 + lane;
  const int m = q * 32 + lane;                    // row of the tile = pixel
  int it = 0;
  int cur_nt = -1;
  // per-lane running column sums of the CTA's tiles (this warp's 16-column groups: at most 4 for BN = 128): even lanes own one
  // column per group.  fp32 over a handful of per-tile partials, in a fixed order; converted to fp64 only when flushed.
  float run_s[4] = {0.f, 0.f, 0.f, 0.f}, run_q[4] = {0.f, 0.f, 0.f, 0.f};
  const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  auto spill_running = [&]() {                    // running sums -> this quadrant's fp64 slots (no other writer of these slots)
    if ((lane & 1) == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c0 = half * 16 + g * 32;
        if (c0 < p.BN) {
          double* slot = s_slots + q * 2 * p.BN + c0 + col;
          slot[0] = (double)run_s[g];
          slot[p.BN] = (double)run_q[g];
        }
        run_s[g] = 0.f; run_q[g] = 0.f;
      }
    }
  };
  const int csz = p.cm * p.cn;
  const int crank = csz > 1 ? (int)cluster_ctarank() : 0;
  const int rank_m = crank / p.cn, rank_n = crank - rank_m * p.cn;
  for (int sup = blockIdx.x / csz; sup < p.total_super; sup += gridDim.x / csz, ++it) {
    const int acc = it & 1;
    const int snt = sup % p.super_n;
    const int nt = snt * p.cn + rank_n, mt = (sup / p.super_n) * p.cm + rank_m;      // mt >= m_tiles: ghost tile of a partial cluster
    if (p.stats && nt != cur_nt) {
      if (cur_nt >= 0) { spill_running(); epilogue_flush_stats(p, s_slots, cur_nt, tid_e); }
      cur_nt = nt;
    }
    int tx = mt % p.tiles_x; int rest = mt / p.tiles_x;
    int ty = rest % p.tiles_y; int img = rest / p.tiles_y;
    int py = m / p.TW, px = m - py * p.TW;
    if (p.halo == 2) { px = m / p.TH; py = m - px * p.TH; }     // groups of 8 accumulator rows run down a column
    const int oy = ty * p.TH + py, ox = tx * p.TW + px;
    const bool valid = (m < p.TH * p.TW) && oy < p.Ho && ox < p.Wo && mt < p.m_tiles;
    float* orow = p.out + (((size_t)img * p.out_ph + oy + p.out_ring) * p.out_pw + ox + p.out_ring) * p.out_ct + p.out_coff + nt * p.BN;
    // accumulating launches (a gradient buffer with a second consumer): the old values of this warp's first 16-column group are
    // requested BEFORE the wait for the accumulator, so their latency hides behind the tile's MMAs
    float4 old0[4];
    const bool pre_old = p.accumulate && !(p.dbg & 1) && valid && half * 16 < p.BN;
    if (pre_old) {
#pragma unroll
      for (int j = 0; j < 4; ++j) old0[j] = *reinterpret_cast<const float4*>(orow + half * 16 + 4 * j);
    }
    mbar_wait(&tmem_full[acc], ((uint32_t)it >> 1) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c0 = half * 16 + g * 32;
      if (c0 >= p.BN) break;
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
      if (valid && !(p.dbg & 2)) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4* dst = reinterpret_cast<float4*>(orow + c0 + j);
          float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.accumulate && !(p.dbg & 1)) {
            const float4 old = (g == 0 && pre_old) ? old0[j / 4] : *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *dst = o;
        }
      }
      if (p.stats && !(p.dbg & 4)) {
        float sq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] = valid ? v[j] : 0.f; sq[j] = v[j] * v[j]; }
        // (order-dependent fp32 atomics made the forward maps differ by ~1e-5 from run to run through the cancellation in
        // E[x^2] - mean^2, tools/diag_determinism2.py: everything here runs in a fixed order; fp64 from the flush on)
        run_s[g] += colsum16(v, lane);
        run_q[g] += colsum16(sq, lane);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
  }
  if (p.stats && cur_nt >= 0) { spill_running(); epilogue_flush_stats(p, s_slots, cur_nt, tid_e); }
}

template <int NPROD>
__device__ __forceinline__ void conv_tc_body(const CUtensorMap& map_a_hi, const CUtensorMap& map_a_lo,
                                             const CUtensorMap& map_b_hi, const CUtensorMap& map_b_lo, const ConvParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  // dynamic smem: [stages * stage_bytes | 4 x 2*BN fp64 statistics slots]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  double* s_stats = reinterpret_cast<double*>(smem + (size_t)p.stages * p.stage_bytes);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (NPROD == 3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
    // a stage is free once EVERY CTA of the cluster has retired its MMAs on it (the receivers of this CTA's multicasts are its
    // cluster row and column; waiting for the whole cluster also keeps a fast neighbour from arriving twice in one phase)
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], p.cm * p.cn); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  if (p.stats) for (int i = threadIdx.x; i < 8 * p.BN; i += kThreads) s_stats[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();                       // everything above touched no global memory: it may overlap the previous kernel
  pdl_wait();
  const int csz = p.cm * p.cn;
  const int crank = csz > 1 ? (int)cluster_ctarank() : 0;
  const int rank_m = crank / p.cn, rank_n = crank - rank_m * p.cn;
  // cluster row = the cn CTAs on the same pixel tile (share A), cluster column = the cm CTAs on the same channel tile (share B)
  const uint16_t mask_row = (uint16_t)(((1u << p.cn) - 1u) << (rank_m * p.cn));
  uint16_t mask_col = 0;
  for (int i = 0; i < p.cm; ++i) mask_col |= (uint16_t)(1u << (i * p.cn + rank_n));
  if (csz > 1) cluster_sync_all();               // every CTA's barriers are initialised before anything arrives on them remotely

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    {
      int stage = 0; uint32_t phase = 0;
      const bool mc_a = p.cn > 1, mc_b = p.cm > 1;
      // loads of the shared operands take turns: K-iteration kit's A tile is fetched (and multicast) by the CTA of the cluster row
      // with rank_n == kit % cn, its B tile by the CTA of the cluster column with rank_m == kit % cm
      auto load_a = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int kit, int c0, int c1, int c2, int c3) {
        if (!mc_a) tma_load_4d_elect(dst, map, bar, c0, c1, c2, c3);
        else if (kit % p.cn == rank_n) tma_load_4d_mc_elect(dst, map, bar, c0, c1, c2, c3, mask_row);
      };
      auto load_b3 = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int kit, int c0, int c1, int c2) {
        if (!mc_b) tma_load_3d_elect(dst, map, bar, c0, c1, c2);
        else if (kit % p.cm == rank_m) tma_load_3d_mc_elect(dst, map, bar, c0, c1, c2, mask_col);
      };
      auto load_b2 = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int kit, int c0, int c1) {
        if (!mc_b) tma_load_2d_elect(dst, map, bar, c0, c1);
        else if (kit % p.cm == rank_m) tma_load_2d_mc_elect(dst, map, bar, c0, c1, mask_col);
      };
      for (int sup = blockIdx.x / csz; sup < p.total_super; sup += gridDim.x / csz) {
        const int snt = sup % p.super_n;
        const int nt = snt * p.cn + rank_n, mt = (sup / p.super_n) * p.cm + rank_m;   // ghost tiles (mt >= m_tiles) load image N: zero fill
        int tx = mt % p.tiles_x; int rest = mt / p.tiles_x;
        int ty = rest % p.tiles_y; int img = rest / p.tiles_y;
        const int x_base = tx * p.TW * p.stride - p.pad + p.org;
        const int y_base = ty * p.TH * p.stride - p.pad + p.org;
        int tap = 0, cc = 0, r = 0, sx = 0;
        for (int kit = 0; kit < p.kiters; ++kit) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          uint64_t* fb = &full_bar[stage];
          mbar_expect_tx_elect(fb, p.tx_bytes);
          if (p.fold == 2) {
            // A: the whole (TH+KH-1) x TW halo tile of 64-element fat-pixel slice cc, loaded ONCE for all kernel rows
            const int xo = tx * p.TW, yo = ty * p.TH;
            load_a(st, &map_a_hi, fb, kit, kit * 64, xo, yo, img);
            for (int rr = 0; rr < p.KH; ++rr)
              load_b3(st + p.a_bytes + rr * p.b_each, &map_b_hi, fb, kit, kit * 64, rr, nt * p.BN);
            if (NPROD == 3) {
              uint8_t* lo = st + p.a_bytes + p.b_bytes;
              load_a(lo, &map_a_lo, fb, kit, kit * 64, xo, yo, img);
              for (int rr = 0; rr < p.KH; ++rr)
                load_b3(lo + p.a_bytes + rr * p.b_each, &map_b_lo, fb, kit, kit * 64, rr, nt * p.BN);
            }
          } else if (p.fold) {
            // A: 64-element slice cc of the "fat pixel" row (KW taps x Cin channels, contiguous in NHWC) of kernel row r
            const int xo = tx * p.TW, yo = ty * p.TH * p.stride + r;
            load_a(st, &map_a_hi, fb, kit, cc * 64, xo, yo, img);
            load_b3(st + p.a_bytes, &map_b_hi, fb, kit, cc * 64, r, nt * p.BN);
            if (NPROD == 3) {
              load_a(st + p.a_bytes + p.b_bytes, &map_a_lo, fb, kit, cc * 64, xo, yo, img);
              load_b3(st + 2 * p.a_bytes + p.b_bytes, &map_b_lo, fb, kit, cc * 64, r, nt * p.BN);
            }
            if (++cc == p.cchunks) { cc = 0; ++r; }
          } else {
            load_a(st, &map_a_hi, fb, kit, cc * p.KC, x_base + sx * p.dil, y_base + r * p.dil, img);
            load_b2(st + p.a_bytes, &map_b_hi, fb, kit, tap * p.Cin + cc * p.KC, nt * p.BN);
            if (NPROD == 3) {
              load_a(st + p.a_bytes + p.b_bytes, &map_a_lo, fb, kit, cc * p.KC, x_base + sx * p.dil, y_base + r * p.dil, img);
              load_b2(st + 2 * p.a_bytes + p.b_bytes, &map_b_lo, fb, kit, tap * p.Cin + cc * p.KC, nt * p.BN);
            }
            if (++cc == p.cchunks) { cc = 0; ++tap; if (++sx == p.KW) { sx = 0; ++r; } }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop converged and one elected lane issues: with the loop inside `if (lane == 0)` the compiler
    // wrapped every tcgen05.mma in an ELECT / 7x R2UR.BROADCAST / BRA.U.ANY sequence and this single thread, at ~128 cycles
    // per MMA, was the critical path of every convolution.
    {
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const int ksteps = p.KC / 16;
      // one (A, B) operand pair, K-steps [0, ks); `zero` = first MMA of the tile (literal accumulate predicates otherwise)
      auto group = [&](uint32_t d, uint64_t a, uint64_t b, int ks, bool zero) {
        if (zero) umma_bf16_elect(d, a, b, idesc, 0); else umma_bf16_elect(d, a, b, idesc, 1);
        if (ks > 1) umma_bf16_elect(d, a + 2, b + 2, idesc, 1);
        if (ks > 2) umma_bf16_elect(d, a + 4, b + 4, idesc, 1);
        if (ks > 3) umma_bf16_elect(d, a + 6, b + 6, idesc, 1);
      };
      int stage = 0; uint32_t phase = 0; int it = 0;
      const uint16_t mask_free = (uint16_t)((1u << csz) - 1u);
      const uint64_t desc0 = make_desc(smem_u32(smem), p.sbo, p.layout_type);
      const uint32_t stage16 = p.stage_bytes >> 4, ab16 = p.a_bytes >> 4, lo16 = (p.a_bytes + p.b_bytes) >> 4;
      for (int sup = blockIdx.x / csz; sup < p.total_super; sup += gridDim.x / csz, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int kit = 0; kit < p.kiters; ++kit) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)stage * p.stage_bytes);
          if (!p.fold && ksteps == 4) {
            // one election for the 4 / 12 MMAs of the stage, descriptors by 64-bit adds on a base descriptor
            const uint64_t a_hi = desc0 + (uint64_t)((uint32_t)stage * stage16);
            const uint64_t a_lo = a_hi + lo16;
            umma_chunk_elect<NPROD>(d_tmem, a_hi, a_lo, a_hi + ab16, a_lo + ab16, idesc, kit != 0);
          } else if (p.fold == 2) {
            // the last 64-element slice of a fat pixel is partly padding (zero weights): skip its dead K-steps
            const int ks2 = min(ksteps, (p.KW * p.Cin - kit * 64 + 15) >> 4);
            const uint32_t lo = st + p.a_bytes + p.b_bytes;
            for (int rr = 0; rr < p.KH; ++rr) {
              // kernel row rr reads the halo tile rr*TW rows further down (a whole number of 1024-byte swizzle atoms)
              const uint64_t a_hi = make_desc(st + rr * p.tap_bytes, p.sbo, p.layout_type);
              const uint64_t b_hi = make_desc(st + p.a_bytes + rr * p.b_each, p.sbo, p.layout_type);
              group(d_tmem, a_hi, b_hi, ks2, (kit | rr) == 0);
              if (NPROD == 3) {
                const uint64_t a_lo = make_desc(lo + rr * p.tap_bytes, p.sbo, p.layout_type);
                const uint64_t b_lo = make_desc(lo + p.a_bytes + rr * p.b_each, p.sbo, p.layout_type);
                group(d_tmem, a_lo, b_hi, ks2, false);
                group(d_tmem, a_hi, b_lo, ks2, false);
              }
            }
          } else {
            // folded rows: stage kit holds slice (kit % cchunks) of a fat pixel; its padding K-steps are skipped
            const int ks1 = p.fold ? min(ksteps, (p.KW * p.Cin - (kit % p.cchunks) * 64 + 15) >> 4) : ksteps;
            const uint64_t a_hi = make_desc(st, p.sbo, p.layout_type);
            const uint64_t b_hi = make_desc(st + p.a_bytes, p.sbo, p.layout_type);
            group(d_tmem, a_hi, b_hi, ks1, kit == 0);
            if (NPROD == 3) {
              const uint64_t a_lo = make_desc(st + p.a_bytes + p.b_bytes, p.sbo, p.layout_type);
              const uint64_t b_lo = make_desc(st + 2 * p.a_bytes + p.b_bytes, p.sbo, p.layout_type);
              group(d_tmem, a_lo, b_hi, ks1, false);
              group(d_tmem, a_hi, b_lo, ks1, false);
            }
          }
          // frees the smem stage when these MMAs retire (in every CTA that multicasts into it)
          if (csz > 1) umma_commit_mc_elect(&empty_bar[stage], mask_free); else umma_commit_elect(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_elect(&tmem_full[acc]);         // accumulator complete
      }
    }
  } else {
    // ===================== epilogue =====================
    epilogue_warps(p, warp, lane, tmem_base, tmem_full, tmem_empty, s_stats);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (csz > 1) cluster_sync_all();               // no CTA leaves while a peer may still multicast into it or arrive on its barriers
}

#define FSNET_CONV_KERNEL(name, cluster_attr)                                                                              \
  template <int NPROD>                                                                                                     \
  __global__ void cluster_attr __launch_bounds__(kThreads, 3)                                                              \
  name(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,                         \
       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const ConvParams p) {   \
    conv_tc_body<NPROD>(map_a_hi, map_a_lo, map_b_hi, map_b_lo, p);                                                        \
  }
FSNET_CONV_KERNEL(conv_tc_kernel, )
FSNET_CONV_KERNEL(conv_tc_kernel_c2, __cluster_dims__(2, 1, 1))
FSNET_CONV_KERNEL(conv_tc_kernel_c4, __cluster_dims__(4, 1, 1))
FSNET_CONV_KERNEL(conv_tc_kernel_c8, __cluster_dims__(8, 1, 1))
#undef FSNET_CONV_KERNEL

// ---------------------------------------------------------------------------------------------------------------------------------
// Halo variant for 3x3 / stride-1 layers with >= 64 input channels.
//
// Why: the delivery of TMA bytes into an SM is capped near 32 B/clk (tools/ubench/mc_bw.cu: 9-10 TB/s over 148 SMs, the same for
// unicast and for cluster multicast), and the per-tap kernel above needs 85-128 B/clk to keep the tensor pipe busy (it re-reads the
// A tile for each of the nine taps).  Here the 18 x 10 pixel halo of a 16 x 8 pixel tile is loaded once per 64-channel chunk and
// plane; tap (r, s) is that same tile seen through an operand descriptor whose start address is shifted by (r * 10 + s) pixels and
// whose 8-row group stride is the halo pitch (1280 B).  The 128-byte swizzle is a function of the shared-memory ADDRESS bits, so
// shifted descriptors read what TMA wrote (checked on the B200 by tools/ubench/umma_shift.cu for all nine shifts).
// A and the per-tap weight tiles ride in two separate rings with their own producer warps (0: weights, 10: halo tiles).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int kHaloThreads = 352;      // warp 0 weight producer, 1 MMA issuer, 2-9 epilogue, 10 halo producer
constexpr int kHaloMaxA = 8, kHaloMaxB = 8;

// KS = K-steps per tap = channels per chunk / 16: 4 (64-channel chunks, 128-byte rows / swizzle), 2 (32, 64 B), 1 (16, 32 B).
template <int NPROD, int KS>
__global__ void __launch_bounds__(kHaloThreads, KS == 4 ? 1 : 2)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kHaloMaxA], a_empty[kHaloMaxA], b_full[kHaloMaxB], b_empty[kHaloMaxB], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;
  constexpr int KC = 16 * KS;                    // channels per chunk
  constexpr uint32_t RB = 2 * KC;                // bytes per pixel row of the halo tile / per weight row
  constexpr uint32_t LAYOUT = KS == 4 ? 2u : (KS == 2 ? 4u : 6u);

  // dynamic smem: [a_stages * a_stage_bytes | b_stages * b_stage_bytes | 4 x 2*BN fp64 statistics slots]
  // (b_resident: b_stages = 9 * cchunks, every (chunk, tap) weight tile is loaded once per CTA and stays)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + (size_t)p.a_stages * p.a_stage_bytes;
  double* s_stats = reinterpret_cast<double*>(smem_b + (size_t)p.b_stages * p.b_stage_bytes);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_bbar = p.b_resident ? 1 : p.b_stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (NPROD == 3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < n_bbar; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  if (p.stats) for (int i = threadIdx.x; i < 8 * p.BN; i += kHaloThreads) s_stats[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (p.dbg & 32) {
    // diagnostics: prologue and teardown only
  } else if (warp == 10) {
    // ===================== halo producer: one box per (tile, channel chunk, plane) =====================
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int tx = mt % p.tiles_x; const int rest = mt / p.tiles_x;
      const int ty = rest % p.tiles_y; const int img = rest / p.tiles_y;
      const int x0 = tx * p.TW - p.pad + p.org, y0 = ty * p.TH - p.pad + p.org;
      const int c1 = p.halo == 2 ? y0 : x0, c2 = p.halo == 2 ? x0 : y0;          // halo 2: the tensor map's dimension 1 is the image row
      for (int cc = 0; cc < p.cchunks; ++cc) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * p.a_stage_bytes;
        if (p.dbg & 8) { if (lane == 0) mbar_arrive(&a_full[stage]); __syncwarp(); }      // diagnostics: no loads
        else {
        mbar_expect_tx_elect(&a_full[stage], p.a_tx);
        tma_load_4d_elect(st, &map_a_hi, &a_full[stage], cc * KC, c1, c2, img);
        if (NPROD == 3) tma_load_4d_elect(st + p.a_plane_bytes, &map_a_lo, &a_full[stage], cc * KC, c1, c2, img);
        }
        if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ===================== weight producer: one [BN x KC] box per (chunk, tap, plane) =====================
    if (p.b_resident) {
      // thin layers: all 9 * cchunks weight tiles of the (single) channel tile fit next to the halo ring: loaded once per CTA
      if (!(p.dbg & 8)) {
        mbar_expect_tx_elect(&b_full[0], p.b_tx * 9u * (uint32_t)p.cchunks);
        for (int cc = 0; cc < p.cchunks; ++cc)
          for (int tap = 0; tap < 9; ++tap) {
            uint8_t* st = smem_b + (size_t)(cc * 9 + tap) * p.b_stage_bytes;
            tma_load_2d_elect(st, &map_b_hi, &b_full[0], tap * p.Cin + cc * KC, 0);
            if (NPROD == 3) tma_load_2d_elect(st + p.b_plane_bytes, &map_b_lo, &b_full[0], tap * p.Cin + cc * KC, 0);
          }
      } else { if (lane == 0) mbar_arrive(&b_full[0]); __syncwarp(); }
    } else {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        for (int cc = 0; cc < p.cchunks; ++cc) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[stage], phase ^ 1);
            uint8_t* st = smem_b + (size_t)stage * p.b_stage_bytes;
            if (p.dbg & 8) { if (lane == 0) mbar_arrive(&b_full[stage]); __syncwarp(); }
            else {
            mbar_expect_tx_elect(&b_full[stage], p.b_tx);
            tma_load_2d_elect(st, &map_b_hi, &b_full[stage], tap * p.Cin + cc * KC, nt * p.BN);
            if (NPROD == 3) tma_load_2d_elect(st + p.b_plane_bytes, &map_b_lo, &b_full[stage], tap * p.Cin + cc * KC, nt * p.BN);
            }
            if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elected lane; see conv_tc_body) =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    // descriptors = base descriptor + byte offset / 16 (the address field holds address >> 4 and shared memory ends below 2^18):
    // a stage or a tap is one 64-bit add.  Halo tile: 8-pixel groups 10 pixels apart; weights: 8 rows of RB bytes.
    const uint64_t a_base = make_desc(smem_u32(smem), 10u * RB, LAYOUT);
    const uint64_t b_base = make_desc(smem_u32(smem_b), 8u * RB, LAYOUT);
    const uint32_t a_stage16 = p.a_stage_bytes >> 4, b_stage16 = p.b_stage_bytes >> 4, a_lo16 = p.a_plane_bytes >> 4, b_lo16 = p.b_plane_bytes >> 4;
    // halo 1: pixel (y, x) of the halo sits at (y * 10 + x) * RB; halo 2: at (x * 10 + y) * RB  (units of 16 B below)
    const uint32_t row16 = p.halo == 2 ? RB / 16u : 10u * RB / 16u, col16 = p.halo == 2 ? 10u * RB / 16u : RB / 16u;
    int as = 0, bs = 0; uint32_t aph = 0, bph = 0; int it = 0;
    uint32_t a_off16 = 0, b_off16 = 0;
    if (p.b_resident) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_empty[acc], (((uint32_t)it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      if (p.b_resident) b_off16 = 0;
      for (int cc = 0; cc < p.cchunks; ++cc) {
        mbar_wait(&a_full[as], aph);
        tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if (!p.b_resident) { mbar_wait(&b_full[bs], bph); tc_fence_after(); }
          const uint64_t a_hi = a_base + (uint64_t)(a_off16 + (uint32_t)(tap / 3) * row16 + (uint32_t)(tap % 3) * col16);
          const uint64_t b_hi = b_base + (uint64_t)b_off16;
          if (!(p.dbg & 16))                                           // diagnostics: 16 = no MMAs
            umma_chunk_elect<NPROD, KS>(d_tmem, a_hi, a_hi + a_lo16, b_hi, b_hi + b_lo16, idesc, (cc | tap) != 0);
          b_off16 += b_stage16;
          if (!p.b_resident) {
            umma_commit_elect(&b_empty[bs]);
            if (++bs == p.b_stages) { bs = 0; bph ^= 1; b_off16 = 0; }
          }
        }
        umma_commit_elect(&a_empty[as]);
        a_off16 += a_stage16;
        if (++as == p.a_stages) { as = 0; aph ^= 1; a_off16 = 0; }
      }
      umma_commit_elect(&tmem_full[acc]);
    }
  } else {
    epilogue_warps(p, warp, lane, tmem_base, tmem_full, tmem_empty, s_stats);
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

static int encode_act_map(CUtensorMap* map, const fsnet_view* v, int plane, int use_ring, int kc, int box_w, int box_h, int stride,
                          const char* who) {
  EncodeTiledFn enc = encode_fn();
  FSNET_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled not available from the driver", who);
  const int pw = v->w + 2 * v->ring, ph = v->h + 2 * v->ring;
  const size_t plane_elems = (size_t)v->n * ph * pw * v->c_total;
  const char* base = (const char*)v->ptr + (plane_elems * plane + v->c_off) * 2;
  if (!use_ring) base += ((size_t)v->ring * pw + v->ring) * v->c_total * 2;
  const int Wm = use_ring ? pw : v->w, Hm = use_ring ? ph : v->h;
  cuuint64_t dim[4] = {(cuuint64_t)v->c, (cuuint64_t)Wm, (cuuint64_t)Hm, (cuuint64_t)v->n};
  cuuint64_t str[3] = {(cuuint64_t)v->c_total * 2, (cuuint64_t)pw * v->c_total * 2, (cuuint64_t)ph * pw * v->c_total * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  FSNET_REQUIRE(box[1] <= 256 && box[2] <= 256, "%s: TMA box too large", who);
  FSNET_REQUIRE(((uintptr_t)base & 15) == 0 && (str[0] & 15) == 0, "%s: activation view is not 16-byte aligned", who);
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FSNET_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed with %d", who, (int)r);
  return FSNET_OK;
}

// ---------------------------------------------------------------------------------------------------
// fsnet_conv -- see include/fsnet_b200.h
// ---------------------------------------------------------------------------------------------------
extern "C" int fsnet_conv(const fsnet_view* in, int use_ring, const void* w_hi, const void* w_lo, int Cout, int KH, int KW,
                          int stride, int pad, int nprod, const float* bias, int relu, const fsnet_view* out, int accumulate,
                          double* stats, void* stream) {
  FSNET_REQUIRE(in && in->ptr && w_hi && out && out->ptr, "fsnet_conv: null pointer");
  FSNET_REQUIRE(nprod == 1 || (nprod == 3 && w_lo), "fsnet_conv: nprod must be 1 or 3 (3 needs the lo planes)");
  const int N = in->n, H = in->h, W = in->w, Cin = in->c;
  // (Cin = 8: only the 7x7 / stride-2 network stem -- 3 or 6 image channels padded to 8 -- on its folded-tap path, see below)
  FSNET_REQUIRE(N > 0 && H > 0 && W > 0 && (Cin % 16 == 0 || Cin == 8) && Cout % 16 == 0, "fsnet_conv: channels must be multiples of 16 (Cin=%d Cout=%d)", Cin, Cout);
  FSNET_REQUIRE(stride == 1 || stride == 2, "fsnet_conv: stride %d unsupported", stride);
  FSNET_REQUIRE(!use_ring || (in->ring >= pad), "fsnet_conv: replicate padding needs a materialised ring >= pad");
  EncodeTiledFn enc = encode_fn();
  FSNET_REQUIRE(enc != nullptr, "fsnet_conv: cuTensorMapEncodeTiled not available from the driver");

  ConvParams p = {};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.dil = 1;
  p.Ho = (H + 2 * pad - KH) / stride + 1;
  p.Wo = (W + 2 * pad - KW) / stride + 1;
  FSNET_REQUIRE(out->n == N && out->h == p.Ho && out->w == p.Wo && out->c == Cout, "fsnet_conv: output view is [%d,%d,%d,%d], expected [%d,%d,%d,%d]",
                out->n, out->h, out->w, out->c, N, p.Ho, p.Wo, Cout);
  FSNET_REQUIRE(out->c_off % 4 == 0 && out->c_total % 4 == 0, "fsnet_conv: output channel slice must be 16-byte aligned");
  p.org = use_ring ? in->ring : 0;
  pick_tile(p.Ho, p.Wo, p.TH, p.TW);
  // halo path (conv_halo_kernel) for 3x3 / stride-1 layers with 64-channel K chunks: 16 x 8 or 8 x 16 pixel tiles, whichever wastes
  // fewer accumulator rows on this map size; small maps (6 x 20) stay on the per-tap path.  FSNET_CONV_HALO=0 switches it off.
  static int halo_env = -1;
  if (halo_env < 0) { const char* e = getenv("FSNET_CONV_HALO"); halo_env = e ? atoi(e) : 1; }
  p.halo = 0;
  // (FSNET_CONV_HALO=4: only the 64-channel-chunk layers, the thin ones stay on the folded-tap path)
  const bool halo_pad_ok = (pad == 1 && (!use_ring || in->ring == 1)) || (pad == 2 && use_ring && in->ring == 2);
  if (halo_env && KH == 3 && KW == 3 && stride == 1 && halo_pad_ok && Cin % 16 == 0 && (Cin % 64 == 0 || halo_env != 4) &&
      in->c_off % 8 == 0) {
    auto eff = [&](int th, int tw) { return ((double)p.Ho / (ceil_div(p.Ho, th) * th)) * ((double)p.Wo / (ceil_div(p.Wo, tw) * tw)); };
    const double ex = eff(16, 8), ey = eff(8, 16);
    const double cur = ((double)p.Ho / (ceil_div(p.Ho, p.TH) * p.TH)) * ((double)p.Wo / (ceil_div(p.Wo, p.TW) * p.TW)) * (p.TH * p.TW / 128.0);
    if ((ex > ey ? ex : ey) >= 0.7 * cur) {
      p.halo = (halo_env == 3 || (halo_env != 2 && ex >= ey)) ? 1 : 2;      // FSNET_CONV_HALO = 2 / 3: force one orientation (diagnostics)
      p.TH = p.halo == 1 ? 16 : 8; p.TW = p.halo == 1 ? 8 : 16;
    }
  }
  p.tiles_x = ceil_div(p.Wo, p.TW); p.tiles_y = ceil_div(p.Ho, p.TH);
  p.BN = Cout <= 128 ? Cout : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16)));
  FSNET_REQUIRE(Cout % p.BN == 0, "fsnet_conv: cannot tile Cout=%d", Cout);
  p.n_tiles = Cout / p.BN;
  p.total_tiles = N * p.tiles_x * p.tiles_y * p.n_tiles;
  // the 6x20 / 12x40 layers have only 12..48 pixel tiles: narrower channel tiles put more of the 148 SMs to work
  // (every CTA re-reads its A tile, which is tiny there; the MMA time of a tile scales with BN)
  static int split_env = -1;
  if (split_env < 0) { const char* e = getenv("FSNET_CONV_NSPLIT"); split_env = e ? atoi(e) : 1; }
  while (split_env && p.BN >= 64 && p.total_tiles * 2 <= 148 && Cout % (p.BN / 2) == 0) {
    p.BN /= 2; p.n_tiles = Cout / p.BN;
    p.total_tiles = N * p.tiles_x * p.tiles_y * p.n_tiles;
  }
  // wave quantisation (opt-in, FSNET_CONV_WAVE=1|2, off by default until measured): 180 tiles on 148 SMs run as two waves with the
  // second one 22 % full; half-width channel tiles (twice the tiles, half the MMA time each) fill the last wave better.
  // 2 also allows quarter-width tiles for a single under-filled wave (the 96-tile layers at 128 channels per tile).
  static int wave_env = -1;
  if (wave_env < 0) { const char* e = getenv("FSNET_CONV_WAVE"); wave_env = e ? atoi(e) : 0; }
  if (wave_env && p.total_tiles > 74) {
    auto fill = [](int tiles) { return (double)tiles / (148.0 * ((tiles + 147) / 148)); };
    int best = 1;
    if (p.BN >= 64 && Cout % (p.BN / 2) == 0 && fill(p.total_tiles * 2) > fill(p.total_tiles) + 0.1) best = 2;
    else if (wave_env >= 2 && p.total_tiles <= 148 && p.BN >= 128 && Cout % (p.BN / 4) == 0 &&
             fill(p.total_tiles * 4) > fill(p.total_tiles) + 0.15) best = 4;
    if (best > 1) { p.BN /= best; p.n_tiles = Cout / p.BN; p.total_tiles = N * p.tiles_x * p.tiles_y * p.n_tiles; }
  }
  p.KC = pick_kc(Cin); p.cchunks = Cin / p.KC; p.kiters = KH * KW * p.cchunks;
  // x-tap folding for thin replicate-padded layers: the KW taps x Cin channels of a kernel row are one contiguous
  // run in NHWC, read as 64-element slices through an overlapping-stride tensor map (pixel stride = Cin elements)
  static int fold_env = -1;
  if (fold_env < 0) { const char* e = getenv("FSNET_CONV_FOLD"); fold_env = e ? atoi(e) : 2; }
  p.fold = !p.halo && fold_env && stride == 1 && use_ring && in->ring == pad && (pad == 1 || pad == 2) && KH == 3 && KW == 3 && in->c_off == 0 &&
           in->c == in->c_total && (Cin < 64 || Cin == 96);
  // the 7x7 / stride-2 stem on 8-channel (3 or 6 real) image planes with a materialised zero ring: the 7 taps x 8 channels of a
  // kernel row are 56 contiguous elements = ONE 64-element slice (K = 7 x 64 for a real K of 147; 16-channel planes needed two
  // slices per row, i.e. twice the TMA bytes and 1.75x the MMAs -- the stem was the longest forward launch, 165 us at cfg2a)
  const bool stem_fold = fold_env && stride == 2 && use_ring && in->ring == pad && pad == 3 && KH == 7 && KW == 7 && in->c_off == 0 &&
                         in->c == in->c_total && (Cin == 16 || Cin == 8);
  if (stem_fold) p.fold = 1;
  FSNET_REQUIRE(Cin % 16 == 0 || stem_fold, "fsnet_conv: 8 input channels are only supported by the folded 7x7/2 stem (zero ring of 3)");

  if (p.fold) {
    p.KC = 64; p.cchunks = ceil_div(KW * Cin, 64); p.kiters = KH * p.cchunks;
    // halo variant: an 8 x 16 output tile whose 10 x 16 fat-pixel halo is loaded once; needs full 16-wide tiles in smem
    // (rows of the box are TW pixels apart, so the kernel-row offset TW*128 B must be a multiple of the 1024 B atom)
    if (fold_env >= 2 && stride == 1 && p.Ho >= 8 && p.Wo >= 16) {
      p.fold = 2; p.TH = 8; p.TW = 16;
      p.tiles_x = ceil_div(p.Wo, p.TW); p.tiles_y = ceil_div(p.Ho, p.TH);
      p.total_tiles = N * p.tiles_x * p.tiles_y * p.n_tiles;
      p.kiters = p.cchunks;
    }
  }
  p.a_bytes = 128u * p.KC * 2;
  p.b_bytes = ((uint32_t)p.BN * p.KC * 2 + 1023u) & ~1023u;
  p.tx_bytes = (nprod == 3 ? 2u : 1u) * ((uint32_t)p.TH * p.TW * p.KC * 2 + (uint32_t)p.BN * p.KC * 2);
  if (p.fold == 2) {
    const uint32_t rows = (uint32_t)(p.TH + KH - 1) * p.TW;
    p.a_bytes = rows * 128u;
    p.b_each = p.b_bytes;
    p.b_bytes = (uint32_t)KH * p.b_each;
    p.tap_bytes = (uint32_t)p.TW * 128u;
    p.tx_bytes = (nprod == 3 ? 2u : 1u) * (rows * 128u + (uint32_t)KH * p.BN * 128u);
  }
  p.stage_bytes = (nprod == 3 ? 2u : 1u) * (p.a_bytes + p.b_bytes);
  const uint32_t stats_bytes = stats ? 64u * (uint32_t)p.BN : 0u;      // 4 quadrants x (sum, sum of squares) x BN fp64 slots
  uint32_t cols = 32;
  while (cols < 2u * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  // CTAs per SM.  A tile of a thin layer is a handful of small MMAs between fixed latencies (TMA round trip, commit, accumulator
  // hand-over, the accumulating epilogue's read of the old gradient): with one CTA per SM the 16..32-channel full-resolution layers
  // spent ~2600 cycles per 128-pixel tile (ncu r2c10: tensor pipe 4 %, 0.6 warp instructions per cycle and SM, DRAM 10 %).
  // Several CTAs per SM -- each with a shallower pipeline out of the same shared memory, each with its own TMEM columns -- overlap
  // those latencies.  FSNET_CONV_OCC: 0 = one CTA per SM (round 1), 1 = thin layers only (default), 2 = every layer that fits.
  static int occ_env = -1;
  if (occ_env < 0) { const char* e = getenv("FSNET_CONV_OCC"); occ_env = e ? atoi(e) : 1; }
  int occ = 1;
  if (occ_env > 0 && (occ_env >= 2 || p.stage_bytes <= 56u * 1024u)) {
    for (int o = 3; o >= 2; --o) {                     // 320 threads x ~66 registers: three CTAs fill the register file
      const uint32_t per_cta = 216u * 1024u / (uint32_t)o;
      if (per_cta < stats_bytes + 2048u) continue;
      const int st = (int)((per_cta - stats_bytes - 2048u) / p.stage_bytes);
      if (st >= 2 && (uint32_t)o * cols <= 512u && p.total_tiles >= 148 * o * 2) { occ = o; break; }
    }
  }
  // Thread-block clusters with TMA multicast.  These kernels are bound by the delivery of TMA bytes from L2 into the SMs (one
  // 64-channel layer re-reads its input 9x and its weights once per 128-pixel tile: ~9.4 TB/s delivered at cfg2a).
  // A cluster of cm x cn CTAs works on cm pixel tiles x cn channel tiles; each A box is fetched once per cluster row and each B
  // box once per cluster column.  Shape = the one that minimises the bytes a CTA has to fetch itself (a/cn + b/cm) without
  // padding the pixel-tile count by more than 10 % ghost tiles.  FSNET_CONV_MC = largest cluster (1 = off).
  static int mc_env = -1;
  if (mc_env < 0) { const char* e = getenv("FSNET_CONV_MC"); mc_env = e ? atoi(e) : 1; }
  p.cm = p.cn = 1;
  p.m_tiles = N * p.tiles_x * p.tiles_y;
  if (mc_env > 1) {
    const double a = (double)p.a_bytes, b = (double)p.b_bytes;
    double best = a + b;
    for (int cn = 1; cn <= 8; cn *= 2) {
      if (cn > p.n_tiles || p.n_tiles % cn) continue;
      for (int cm = 1; cm <= 8; cm *= 2) {
        if (cm * cn > mc_env || cm * cn > 8 || cm * cn == 1 || cm > p.m_tiles) continue;
        const double ghost = (double)(ceil_div(p.m_tiles, cm) * cm) / p.m_tiles;
        if (ghost > 1.1) continue;
        const double cost = (a / cn + b / cm) * ghost;
        if (cost < best * 0.95) { best = cost; p.cm = cm; p.cn = cn; }
      }
    }
  }
  p.super_n = p.n_tiles / p.cn;
  p.total_super = ceil_div(p.m_tiles, p.cm) * p.super_n;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  if (p.halo) {
    const uint32_t planes = nprod == 3 ? 2u : 1u;
    p.fold = 0; p.KC = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16); p.cchunks = Cin / p.KC; p.kiters = 9 * p.cchunks;
    const uint32_t rb = 2u * (uint32_t)p.KC;             // bytes per pixel row (= per weight row) of a chunk
    p.cm = p.cn = 1; p.super_n = p.n_tiles; p.total_super = p.total_tiles;
    p.a_plane_bytes = (180u * rb + 1023u) & ~1023u;      // 18 x 10 pixels, rounded up to the 1024-byte atom of the widest swizzle
    p.a_stage_bytes = planes * p.a_plane_bytes; p.a_tx = planes * 180u * rb;
    p.b_plane_bytes = (uint32_t)p.BN * rb;
    p.b_stage_bytes = planes * p.b_plane_bytes; p.b_tx = p.b_stage_bytes;
    const uint32_t budget = 220u * 1024u - stats_bytes;
    // weights resident in shared memory when all 9 * cchunks tap tiles of the (single) channel tile fit beside a halo ring of >= 3
    // stages: the thin layers (16..96 input channels) then receive nothing but their halo tiles
    const uint32_t all_b = 9u * (uint32_t)p.cchunks * p.b_stage_bytes;
    p.b_resident = p.n_tiles == 1 && all_b + 3u * p.a_stage_bytes <= budget && all_b <= 128u * 1024u;
    if (p.b_resident) {
      p.b_stages = 9 * p.cchunks;
      int ast = (int)((budget - all_b) / p.a_stage_bytes);
      p.a_stages = ast > kHaloMaxA ? kHaloMaxA : ast;
    } else {
      p.a_stages = nprod == 3 ? 2 : 3;
      if (p.a_stages > p.cchunks + 1) p.a_stages = p.cchunks + 1;
      { static int as_env = -1; if (as_env < 0) { const char* e = getenv("FSNET_CONV_HALO_ASTAGES"); as_env = e ? atoi(e) : 0; } if (as_env) p.a_stages = as_env; }
      int bst = (int)((budget - (uint32_t)p.a_stages * p.a_stage_bytes) / p.b_stage_bytes);
      p.b_stages = bst > kHaloMaxB ? kHaloMaxB : bst;
    }
    // thin layers with resident weights: two CTAs per SM (a tile is ~1.2 K cycles of MMAs between ~400 cycles of hand-overs; the
    // second CTA's MMAs fill them).  FSNET_CONV_HALO_OCC=1 keeps one CTA per SM.
    static int hocc_env = -1;
    if (hocc_env < 0) { const char* e = getenv("FSNET_CONV_HALO_OCC"); hocc_env = e ? atoi(e) : 2; }
    int hocc = 1;
    if (hocc_env >= 2 && p.b_resident && p.KC < 64 && p.total_tiles >= 4 * sms && 2u * p.tmem_cols <= 512u) {
      const uint32_t per_cta = 108u * 1024u - stats_bytes;
      if (all_b + 3u * p.a_stage_bytes <= per_cta) {
        hocc = 2;
        int ast = (int)((per_cta - all_b) / p.a_stage_bytes);
        p.a_stages = ast > kHaloMaxA ? kHaloMaxA : ast;
      }
    }
    FSNET_REQUIRE(p.b_stages >= 2 && p.a_stages >= 2, "fsnet_conv: halo tile does not fit shared memory");
    p.stages = p.a_stages;
    p.out = (float*)out->ptr; p.out_pw = out->w + 2 * out->ring; p.out_ph = out->h + 2 * out->ring; p.out_ring = out->ring;
    p.out_ct = out->c_total; p.out_coff = out->c_off; p.accumulate = accumulate;
    p.bias = bias; p.stats = stats; p.relu = relu;
    { static int dbg_env = -1; if (dbg_env < 0) { const char* e = getenv("FSNET_CONV_DBG"); dbg_env = e ? atoi(e) : 0; } p.dbg = dbg_env; }
    CUtensorMap ma[2], mb[2];
    for (uint32_t pl = 0; pl < planes; ++pl) {
      if (p.halo == 1) {
        int rc = encode_act_map(&ma[pl], in, (int)pl, use_ring, p.KC, 10, 18, 1, "fsnet_conv(halo)");
        if (rc) return rc;
      } else {
        const int pw = in->w + 2 * in->ring, ph = in->h + 2 * in->ring;
        const size_t plane_elems = (size_t)in->n * ph * pw * in->c_total;
        const char* base = (const char*)in->ptr + (plane_elems * pl + in->c_off) * 2;
        if (!use_ring) base += ((size_t)in->ring * pw + in->ring) * in->c_total * 2;
        const int Wm = use_ring ? pw : in->w, Hm = use_ring ? ph : in->h;
        cuuint64_t dim[4] = {(cuuint64_t)in->c, (cuuint64_t)Hm, (cuuint64_t)Wm, (cuuint64_t)in->n};
        cuuint64_t str[3] = {(cuuint64_t)pw * in->c_total * 2, (cuuint64_t)in->c_total * 2, (cuuint64_t)ph * pw * in->c_total * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.KC, 10, 18, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        FSNET_REQUIRE(((uintptr_t)base & 15) == 0, "fsnet_conv(halo): activation view is not 16-byte aligned");
        CUresult r = enc(&ma[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv(halo): cuTensorMapEncodeTiled(A, rows fastest) failed with %d", (int)r);
      }
      const int Ktot = 9 * Cin;
      cuuint64_t bdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
      cuuint64_t bstr[1] = {(cuuint64_t)Ktot * 2};
      cuuint32_t bbox[2] = {(cuuint32_t)p.KC, (cuuint32_t)p.BN};
      cuuint32_t bes[2] = {1, 1};
      CUresult r = enc(&mb[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(pl ? w_lo : w_hi), bdim, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv(halo): cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }
    if (planes == 1) { ma[1] = ma[0]; mb[1] = mb[0]; }
    const int grid = p.total_tiles < sms * hocc ? p.total_tiles : sms * hocc;
    const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes + stats_bytes + 1024;
    typedef void (*HaloKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvParams);
    static const HaloKernel kernels[2][3] = {{conv_halo_kernel<1, 1>, conv_halo_kernel<1, 2>, conv_halo_kernel<1, 4>},
                                             {conv_halo_kernel<3, 1>, conv_halo_kernel<3, 2>, conv_halo_kernel<3, 4>}};
    const int ki = p.KC == 64 ? 2 : (p.KC == 32 ? 1 : 0);
    HaloKernel kern = kernels[nprod == 3][ki];
    static bool halo_attr[2][3] = {};
    if (!halo_attr[nprod == 3][ki]) {
      FSNET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      halo_attr[nprod == 3][ki] = true;
    }
    FSNET_LAUNCH_PDL(kern, grid, kHaloThreads, smem, (cudaStream_t)stream, ma[0], ma[1], mb[0], mb[1], p);
    FSNET_LAUNCH_OK();
    return FSNET_OK;
  }
  const uint32_t smem_budget = occ == 1 ? 200u * 1024u : 216u * 1024u / (uint32_t)occ - 2048u;
  int stages = (int)((smem_budget - stats_bytes) / p.stage_bytes);
  p.stages = stages > kMaxStages ? kMaxStages : stages;
  FSNET_REQUIRE(p.stages >= 2, "fsnet_conv: tile does not fit shared memory");
  p.sbo = 8u * p.KC * 2;
  p.layout_type = p.KC == 64 ? 2u : (p.KC == 32 ? 4u : 6u);
  p.out = (float*)out->ptr; p.out_pw = out->w + 2 * out->ring; p.out_ph = out->h + 2 * out->ring; p.out_ring = out->ring;
  p.out_ct = out->c_total; p.out_coff = out->c_off; p.accumulate = accumulate;
  p.bias = bias; p.stats = stats; p.relu = relu;
  { static int dbg_env = -1; if (dbg_env < 0) { const char* e = getenv("FSNET_CONV_DBG"); dbg_env = e ? atoi(e) : 0; } p.dbg = dbg_env; }

  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (p.fold) {
    const int pw = W + 2 * in->ring, ph = H + 2 * in->ring;
    const size_t plane_elems = (size_t)N * ph * pw * Cin;
    cuuint64_t adim[4] = {(cuuint64_t)(64 * p.cchunks), (cuuint64_t)p.Wo, (cuuint64_t)ph, (cuuint64_t)N};
    cuuint64_t astr[3] = {(cuuint64_t)stride * Cin * 2, (cuuint64_t)pw * Cin * 2, (cuuint64_t)ph * pw * Cin * 2};
    cuuint32_t abox[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)(p.fold == 2 ? p.TH + KH - 1 : p.TH * stride), 1};
    cuuint32_t aes[4] = {1, 1, (cuuint32_t)stride, 1};
    cuuint64_t wdim[3] = {(cuuint64_t)(KW * Cin), (cuuint64_t)KH, (cuuint64_t)Cout};
    cuuint64_t wstr[2] = {(cuuint64_t)KW * Cin * 2, (cuuint64_t)KH * KW * Cin * 2};
    cuuint32_t wbox[3] = {64, 1, (cuuint32_t)p.BN};
    cuuint32_t wes[3] = {1, 1, 1};
    for (int pl = 0; pl < (nprod == 3 ? 2 : 1); ++pl) {
      CUresult r = enc(pl ? &ma_lo : &ma_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (char*)in->ptr + plane_elems * pl * 2, adim, astr, abox, aes,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv: cuTensorMapEncodeTiled(folded A) failed with %d", (int)r);
      r = enc(pl ? &mb_lo : &mb_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)(pl ? w_lo : w_hi), wdim, wstr, wbox, wes,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv: cuTensorMapEncodeTiled(folded B) failed with %d", (int)r);
    }
    if (nprod != 3) { ma_lo = ma_hi; mb_lo = mb_hi; }
  }
  int rc = 0;
  if (!p.fold) {
  rc = encode_act_map(&ma_hi, in, 0, use_ring, p.KC, p.TW, p.TH, stride, "fsnet_conv");
  if (rc) return rc;
  ma_lo = ma_hi;
  if (nprod == 3) { rc = encode_act_map(&ma_lo, in, 1, use_ring, p.KC, p.TW, p.TH, stride, "fsnet_conv"); if (rc) return rc; }
  }
  const int Ktot = KH * KW * Cin;
  cuuint64_t bdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t bstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t bbox[2] = {(cuuint32_t)p.KC, (cuuint32_t)p.BN};
  cuuint32_t bes[2] = {1, 1};
  CUresult r = CUDA_SUCCESS;
  if (!p.fold) {
    r = enc(&mb_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_hi, bdim, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    mb_lo = mb_hi;
  }
  if (nprod == 3 && !p.fold) {
    r = enc(&mb_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_lo, bdim, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle_for(p.KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv: cuTensorMapEncodeTiled(B lo) failed with %d", (int)r);
  }

  const size_t smem = (size_t)p.stages * p.stage_bytes + stats_bytes + 1024;
  const int csz = p.cm * p.cn, ci = csz == 8 ? 3 : (csz == 4 ? 2 : (csz == 2 ? 1 : 0));
  typedef void (*ConvKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvParams);
  static const ConvKernel kernels[2][4] = {{conv_tc_kernel<1>, conv_tc_kernel_c2<1>, conv_tc_kernel_c4<1>, conv_tc_kernel_c8<1>},
                                           {conv_tc_kernel<3>, conv_tc_kernel_c2<3>, conv_tc_kernel_c4<3>, conv_tc_kernel_c8<3>}};
  ConvKernel kern = kernels[nprod == 3][ci];
  static bool attr_set[2][4] = {};
  if (!attr_set[nprod == 3][ci]) {
    FSNET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    if (csz == 8) FSNET_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set[nprod == 3][ci] = true;
  }
  int grid = p.total_tiles < sms * occ ? p.total_tiles : sms * occ;
  if (csz > 1) {
    // persistent clusters: as many as the device can hold at once (GPCs whose SM count is no multiple of the cluster size lose a few)
    static int slots[2][4][5] = {};                                  // [nprod][cluster size][CTAs per SM]
    int& slot = slots[nprod == 3][ci][occ];
    if (!slot) {
      slot = sms * occ / csz;
#ifndef FSNET_HOST_PLAN_ONLY
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(sms * occ / csz * csz)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) == cudaSuccess && n > 0) slot = n;
      else cudaGetLastError();
      if (getenv("FSNET_CONV_MC_PRINT")) fprintf(stderr, "fsnet_conv: cluster %d, %d CTAs/SM, %zu B smem: %d clusters resident\n", csz, occ, smem, slot);
#endif
    }
    const int clusters = p.total_super < slot ? p.total_super : slot;
    grid = clusters * csz;
  }
  FSNET_LAUNCH_PDL(kern, grid, kThreads, smem, (cudaStream_t)stream, ma_hi, ma_lo, mb_hi, mb_lo, p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}

// ===================================================================================================
// weight gradient:  dW[co, r, s, ci] = sum_{n, yo, xo} dy[n, yo, xo, co] * x[n, yo*st + r - p, xo*st + s - p, ci]
//
// GEMM with M = Cout (128 rows, duplicated when Cout < 128), N = Cin tile, K = pixels.  Both operands are
// read straight from the NHWC planes, i.e. MN-major (the contiguous dimension is the channel): one TMA box
// per 64/32/16-channel swizzle atom, [channels, PW, PH] -> 64 pixel rows, UMMA descriptors with
// a_major = b_major = MN.  One CTA per (tap, co tile, ci tile, K split); the fp32 tile is reduced into the
// accumulator with atomics.
// ===================================================================================================
namespace fsnet {
namespace {


struct WgradParams {
  int N, Ho, Wo, KH, KW, stride, pad, org;
  int Cout, Cin;                    // padded channel counts of the accumulator
  int co_tiles, ci_tiles, ksplit, taps;
  int BM_real, BN;                  // real rows (channels of dy) in the tile, ci tile width
  int atomA, atomB, nA, nB;         // channels per swizzle atom, atoms actually loaded
  int pix;                          // pixels (K) per pipeline stage: 64, 128 or 256
  int PW, PH, chunks_x, chunks_y, total_chunks, chunks_per_split;
  int stages;
  uint32_t a_atom_bytes, b_atom_bytes, a_bytes, b_bytes, stage_bytes, tx_bytes;
  uint32_t tmem_cols;
  uint32_t a_layout, b_layout;
  float* acc;
  // folded mode (thin 3x3 layers, the 7x7 stem): the N operand is a 64-element slice of the "fat pixel" (KW taps x Cin
  // channels, contiguous in NHWC, read through an overlapping-stride tensor map), one work item per KERNEL ROW:
  // KW x fewer MMAs and dy re-reads than one item per tap.  taps = KH, ci_tiles = slices, row_elems = KW * Cin.
  int fold, row_elems;
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // stride between swizzle atoms along M/N
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride between groups of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

constexpr int kWgradThreads = 192;     // warp 0 TMA producer, 1 MMA issuer, 2-5 epilogue
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  // work item
  int item = blockIdx.x;
  const int ks = item % p.ksplit; item /= p.ksplit;
  const int cit = item % p.ci_tiles; item /= p.ci_tiles;
  const int cot = item % p.co_tiles; const int tap = item / p.co_tiles;
  const int r = tap / p.KW, s = tap - r * p.KW;
  const int chunk_begin = ks * p.chunks_per_split;
  const int chunk_end = min(chunk_begin + p.chunks_per_split, p.total_chunks);
  const int nchunks = max(chunk_end - chunk_begin, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    {   // whole warp converged, one elected lane issues (see conv_tc_kernel)
      int stage = 0; uint32_t phase = 0;
      for (int ch = chunk_begin; ch < chunk_end; ++ch) {
        int cx = ch % p.chunks_x; int rest = ch / p.chunks_x;
        int cy = rest % p.chunks_y; int img = rest / p.chunks_y;
        const int xo = cx * p.PW, yo = cy * p.PH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * p.stage_bytes;
        mbar_expect_tx_elect(&full_bar[stage], p.tx_bytes);
        for (int a = 0; a < p.nA; ++a)
          tma_load_4d_elect(st + a * p.a_atom_bytes, &map_dy, &full_bar[stage], cot * 128 + a * p.atomA, xo, yo, img);
        if (p.fold) {
          tma_load_4d_elect(st + p.a_bytes, &map_x, &full_bar[stage], cit * 64, xo, yo * p.stride + tap, img);
        } else {
          for (int b = 0; b < p.nB; ++b)
            tma_load_4d_elect(st + p.a_bytes + b * p.b_atom_bytes, &map_x, &full_bar[stage], cit * p.BN + b * p.atomB,
                        xo * p.stride + s - p.pad + p.org, yo * p.stride + r - p.pad + p.org, img);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // D=f32, A=B=bf16, both MN-major (bits 15, 16), N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t rowA = (uint32_t)p.atomA * 2, rowB = (uint32_t)p.atomB * 2;
      // when fewer than 128/atomA atoms exist, LBO = 0 makes the missing atoms alias atom 0 (duplicate rows, never stored)
      const uint32_t lboA = (p.nA * p.atomA >= 128) ? p.a_atom_bytes : 0u;
      int stage = 0; uint32_t phase = 0;
      // descriptors = base + byte offset / 16; four K-steps (16 pixels each) per statement with one election (see umma_chunk_elect)
      const uint64_t da0 = make_desc_mn(smem_u32(smem), lboA, 8 * rowA, p.a_layout);
      const uint64_t db0 = make_desc_mn(smem_u32(smem) + p.a_bytes, p.b_atom_bytes, 8 * rowB, p.b_layout);
      const uint64_t stepA = rowA, stepB = rowB;                     // 16 pixel rows = 16 * row bytes = row bytes in units of 16 B
      const uint32_t stage16 = p.stage_bytes >> 4;
      for (int i = 0; i < nchunks; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        uint64_t da = da0 + (uint64_t)((uint32_t)stage * stage16), db = db0 + (uint64_t)((uint32_t)stage * stage16);
        for (int k = 0; k < p.pix / 64; ++k) {
          asm volatile(
              "{\n"
              ".reg .pred pe, pz, pt;\n"
              ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
              "elect.sync _|pe, 0xffffffff;\n"
              "setp.ne.b32 pz, %6, 0;\n"
              "setp.eq.b32 pt, %6, %6;\n"
              "add.u64 a1, %1, %4;\n add.u64 a2, a1, %4;\n add.u64 a3, a2, %4;\n"
              "add.u64 b1, %2, %5;\n add.u64 b2, b1, %5;\n add.u64 b3, b2, %5;\n"
              "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pz;\n"
              "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n"
              "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n"
              "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n"
              "}\n" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "l"(stepA), "l"(stepB), "r"((uint32_t)((i | k) != 0)) : "memory");
          da += 4 * stepA; db += 4 * stepB;
        }
        umma_commit_elect(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(&done_bar);
    }
  } else if (nchunks > 0) {
    const int q = warp & 3;
    const int m = q * 32 + lane;              // row = output channel within the tile
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int co = cot * 128 + m;
    const bool valid = m < p.BM_real && co < p.Cout;
    float* dst = p.acc + ((size_t)co * p.taps + tap) * p.row_elems + cit * p.BN;
    const int n_cols = min(p.BN, p.row_elems - cit * p.BN);       // folded mode: the last slice is partly padding
    for (int c0 = 0; c0 < n_cols; c0 += 16) {
      uint32_t raw[16];
      tmem_ld16(taddr + c0, raw);
      tmem_ld_wait();
      if (valid) {
        if (p.ksplit == 1) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (c0 + j < n_cols)                // (a folded row of 7 taps x 8 channels = 56 columns ends inside a group of 16)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]),
                                                                     __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 4)       // one 128-bit reduction per four accumulators (the L2 atomic units are the limit)
            if (c0 + j < n_cols)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(__uint_as_float(raw[j])),
                           "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])), "f"(__uint_as_float(raw[j + 3])) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

int atom_channels(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }
uint32_t layout_for_atom(int a) { return a == 64 ? 2u : (a == 32 ? 4u : 6u); }

}  // namespace
}  // namespace fsnet

extern "C" int fsnet_conv_wgrad(const fsnet_view* x, int use_ring, const fsnet_view* dy, int KH, int KW, int stride, int pad,
                                float* acc, void* stream) {
  FSNET_REQUIRE(x && dy && x->ptr && dy->ptr && acc, "fsnet_conv_wgrad: null pointer");
  FSNET_REQUIRE((x->c % 16 == 0 || x->c == 8) && dy->c % 16 == 0, "fsnet_conv_wgrad: channels must be multiples of 16");
  FSNET_REQUIRE(stride == 1 || stride == 2, "fsnet_conv_wgrad: stride %d unsupported", stride);
  WgradParams p = {};
  p.N = dy->n; p.Ho = dy->h; p.Wo = dy->w; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
  p.org = use_ring ? x->ring : 0;
  p.Cout = dy->c; p.Cin = x->c; p.taps = KH * KW;
  FSNET_REQUIRE((x->h + 2 * pad - KH) / stride + 1 == p.Ho && (x->w + 2 * pad - KW) / stride + 1 == p.Wo, "fsnet_conv_wgrad: shape mismatch");
  p.atomA = atom_channels(p.Cout); p.atomB = atom_channels(p.Cin % 16 == 0 ? p.Cin : 16);
  const int BMr = p.Cout < 128 ? p.Cout : 128;
  FSNET_REQUIRE(p.Cout % BMr == 0, "fsnet_conv_wgrad: cannot tile Cout=%d", p.Cout);
  p.BM_real = BMr; p.co_tiles = p.Cout / BMr;
  p.nA = BMr / p.atomA;
  FSNET_REQUIRE(p.nA == 1 || p.nA * p.atomA == 128, "fsnet_conv_wgrad: Cout=%d needs partial atom aliasing (unsupported)", p.Cout);
  p.BN = p.Cin <= 128 ? (p.Cin % 16 == 0 ? p.Cin : 16) : (p.Cin % 128 == 0 ? 128 : (p.Cin % 64 == 0 ? 64 : (p.Cin % 32 == 0 ? 32 : 16)));
  FSNET_REQUIRE(p.Cin % p.BN == 0 || p.Cin == 8, "fsnet_conv_wgrad: cannot tile Cin=%d", p.Cin);
  p.ci_tiles = p.Cin / p.BN; p.nB = p.BN / p.atomB;
  p.row_elems = p.Cin;
  static int wfold_env = -1;
  if (wfold_env < 0) { const char* e = getenv("FSNET_WGRAD_FOLD"); wfold_env = e ? atoi(e) : 1; }
  p.fold = wfold_env && use_ring && x->ring == pad && x->c_off == 0 && x->c == x->c_total &&
           ((stride == 1 && KH == 3 && KW == 3 && pad == 1 && p.Cin < 64) || (stride == 2 && KH == 7 && KW == 7 && pad == 3 && (p.Cin == 16 || p.Cin == 8)));
  FSNET_REQUIRE(p.Cin % 16 == 0 || p.fold, "fsnet_conv_wgrad: 8 input channels are only supported by the folded 7x7/2 stem");
  if (p.fold) {
    p.row_elems = KW * p.Cin;
    p.taps = KH;
    p.BN = 64; p.atomB = 64; p.nB = 1;
    p.ci_tiles = ceil_div(p.row_elems, 64);
  }
  // thin layers get more pixels per stage so that one stage moves >= 16 KB
  p.pix = 64;
  while (false && p.pix < 128 && (size_t)(p.pix * 2) * (p.nA * p.atomA + p.BN) * 2 <= 32 * 1024 && (size_t)p.pix * 2 <= (size_t)p.Ho * p.Wo) p.pix *= 2;
  int pw = 1;
  while (pw * 2 <= p.Wo && pw * 2 <= 64) pw *= 2;
  p.PW = pw; p.PH = p.pix / pw;
  p.chunks_x = ceil_div(p.Wo, p.PW); p.chunks_y = ceil_div(p.Ho, p.PH);
  p.total_chunks = p.N * p.chunks_x * p.chunks_y;
  const int base_items = p.taps * p.co_tiles * p.ci_tiles;
  // K split: ~3 CTAs per SM hide the TMA latency of thin layers (32-64 byte rows); once the (tap, tile) grid alone
  // fills half the machine, more splits only add fp32 atomics on large weight tensors
  // every K-split CTA adds a full [Cout, taps, Cin] tile through L2 reductions: with many outputs one wave of CTAs is
  // enough (measured: the reductions, not the MMAs, bound the 128..512-channel layers), with few outputs three
  static int occ_env = -1, big_env = -1;
  if (occ_env < 0) { const char* e = getenv("FSNET_WGRAD_OCC"); occ_env = e ? atoi(e) : 0; }
  if (big_env < 0) { const char* e = getenv("FSNET_WGRAD_BIG"); big_env = e ? atoi(e) : 64; }
  const long outputs = (long)p.Cout * p.taps * p.Cin;
  const int occ = occ_env ? occ_env : (outputs <= (long)big_env * 1024 ? 3 : 1);
  int ksplit = base_items >= 148 ? 1 : (occ == 1 ? (148 / base_items > 0 ? 148 / base_items : 1) : ceil_div(148 * occ, base_items));
  if (ksplit > p.total_chunks) ksplit = p.total_chunks;
  if (ksplit < 1) ksplit = 1;
  p.chunks_per_split = ceil_div(p.total_chunks, ksplit);
  p.ksplit = ceil_div(p.total_chunks, p.chunks_per_split);
  p.a_atom_bytes = (uint32_t)p.pix * p.atomA * 2; p.b_atom_bytes = (uint32_t)p.pix * p.atomB * 2;
  p.a_bytes = ((uint32_t)p.nA * p.a_atom_bytes + 1023u) & ~1023u;
  p.b_bytes = ((uint32_t)p.nB * p.b_atom_bytes + 1023u) & ~1023u;
  p.stage_bytes = p.a_bytes + p.b_bytes;
  p.tx_bytes = (uint32_t)p.nA * p.a_atom_bytes + (uint32_t)p.nB * p.b_atom_bytes;
  // K-split layers run `occ` CTAs per SM: size the pipeline so that they actually fit next to each other (their fixed
  // costs -- barrier / TMEM set-up, pipeline ramp, reduction epilogue -- then overlap)
  static int share_env = -1;
  if (share_env < 0) { const char* e = getenv("FSNET_WGRAD_SHARE"); share_env = e ? atoi(e) : 1; }
  const uint32_t budget = (share_env && occ > 1 && p.ksplit > 1) ? 200u * 1024u / (uint32_t)occ : 200u * 1024u;
  int stages = (int)(budget / p.stage_bytes);
  p.stages = stages > kMaxStages ? kMaxStages : (stages < 2 ? 2 : stages);
  uint32_t cols = 32;
  while (cols < (uint32_t)p.BN) cols <<= 1;
  p.tmem_cols = cols;
  p.a_layout = layout_for_atom(p.atomA); p.b_layout = layout_for_atom(p.atomB);
  p.acc = acc;
  CUtensorMap mdy, mx;
  int rc = encode_act_map(&mdy, dy, 0, 0, p.atomA, p.PW, p.PH, 1, "fsnet_conv_wgrad");
  if (rc) return rc;
  if (p.fold) {
    EncodeTiledFn enc = encode_fn();
    FSNET_REQUIRE(enc != nullptr, "fsnet_conv_wgrad: cuTensorMapEncodeTiled not available from the driver");
    const int pw = x->w + 2 * x->ring, ph = x->h + 2 * x->ring;
    cuuint64_t adim[4] = {(cuuint64_t)(64 * p.ci_tiles), (cuuint64_t)p.Wo, (cuuint64_t)ph, (cuuint64_t)p.N};
    cuuint64_t astr[3] = {(cuuint64_t)stride * p.Cin * 2, (cuuint64_t)pw * p.Cin * 2, (cuuint64_t)ph * pw * p.Cin * 2};
    cuuint32_t abox[4] = {64, (cuuint32_t)p.PW, (cuuint32_t)(p.PH * stride), 1};
    cuuint32_t aes[4] = {1, 1, (cuuint32_t)stride, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x->ptr, adim, astr, abox, aes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FSNET_REQUIRE(r == CUDA_SUCCESS, "fsnet_conv_wgrad: cuTensorMapEncodeTiled(folded x) failed with %d", (int)r);
  } else {
    rc = encode_act_map(&mx, x, 0, use_ring, p.atomB, p.PW, p.PH, stride, "fsnet_conv_wgrad");
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    FSNET_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_set = true;
  }
  const int grid = base_items * p.ksplit;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  FSNET_LAUNCH_PDL(wgrad_tc_kernel, grid, kWgradThreads, smem, (cudaStream_t)stream, mdy, mx, p);
  FSNET_LAUNCH_OK();
  return FSNET_OK;
}
