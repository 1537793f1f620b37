// Micro-test: can one (TH+2) x (TW+2) halo tile in shared memory (128B swizzle, written by TMA) feed all nine taps of a 3x3
// convolution through tcgen05.mma operand descriptors whose start address is shifted by whole pixels (128 B, not a multiple of
// the 1024-byte swizzle atom) and whose 8-row group stride is the halo pitch?  Tile: TH = 16 rows of TW = 8 pixels, 64 channels.
// Variants: pitch 1280 B (dense 10-pixel halo rows, one TMA box) / 2048 B (rows padded to 16 pixels, one TMA box per row);
//           descriptor base_offset field = 0 / (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_shift umma_shift.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int TH = 16, TW = 8, HR = TH + 2, HC = TW + 2, NOUT = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}

// out[variant 0..3][tap 0..8][128][64]
template <int C>
__global__ void __launch_bounds__(128, 1) shift_kernel(const __grid_constant__ CUtensorMap map_dense, const __grid_constant__ CUtensorMap map_row,
                                                       const __grid_constant__ CUtensorMap map_b, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t ld_bar, mma_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t RB = C * 2;                 // bytes per pixel row: 128 / 64 / 32
  constexpr uint32_t LAYOUT = C == 64 ? 2u : (C == 32 ? 4u : 6u);
  uint8_t* a_dense = smem;                       // 18 * 10 * RB <= 23040 B -> 23552
  uint8_t* a_pad = smem + 23552;                 // 18 * 16 * RB <= 36864
  uint8_t* b_sm = smem + 23552 + 36864;          // 64 * RB <= 8192
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&ld_bar, 1); mbar_init(&mma_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the padded variant's unused pixels so stale shared memory cannot hide an addressing error
  for (int i = threadIdx.x; i < 36864 / 4; i += 128) reinterpret_cast<uint32_t*>(a_pad)[i] = 0x7fc07fc0u;   // bf16 NaNs
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&ld_bar, HR * HC * RB * 2 + NOUT * RB);
    tma_3d(a_dense, &map_dense, &ld_bar, 0, 0, 0);
    for (int y = 0; y < HR; ++y) tma_3d(a_pad + y * 16 * RB, &map_row, &ld_bar, 0, 0, y);
    tma_2d(b_sm, &map_b, &ld_bar, 0, 0);
  }
  mbar_wait(&ld_bar, 0);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t ph = 0;
  for (int variant = 0; variant < 4; ++variant) {
    const uint32_t pitch = (variant & 1) ? 16u * RB : 10u * RB;
    const uint8_t* a_base = (variant & 1) ? a_pad : a_dense;
    for (int tap = 0; tap < 9; ++tap) {
      const int r = tap / 3, s = tap % 3;
      if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t start = smem_u32(a_base) + r * pitch + s * RB;
        const uint32_t boff = (variant & 2) ? ((start >> 7) & 7) : 0;
        const uint64_t ad = make_desc(start, pitch, boff, LAYOUT), bd = make_desc(smem_u32(b_sm), 8 * RB, 0, LAYOUT);
        for (int k = 0; k < C / 16; ++k) {
          asm volatile("{\n.reg .pred pa;\nsetp.ne.b32 pa, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n}\n"
                       ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar)) : "memory");
      }
      mbar_wait(&mma_bar, ph); ph ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* o = out + ((size_t)(variant * 9 + tap) * 128 + warp * 32 + lane) * NOUT;
      for (int c0 = 0; c0 < NOUT; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) o[c0 + j] = __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int C>
int run(EncodeTiledFn enc) {
  std::vector<__nv_bfloat16> hx(HR * HC * C), hb(NOUT * C);
  std::vector<float> fx(HR * HC * C), fb(NOUT * C);
  srand(7);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 9 - 4); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *dx, *db; float* dout;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 4 * 9 * 128 * NOUT * 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap md, mr, mb;
  cuuint64_t dim[3] = {C, HC, HR}; cuuint64_t str[2] = {C * 2, HC * C * 2}; cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t boxd[3] = {C, HC, HR}, boxr[3] = {C, HC, 1};
  CUresult r1 = enc(&md, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, dim, str, boxd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&mr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, dim, str, boxr, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t bdim[2] = {C, NOUT}; cuuint64_t bstr[1] = {C * 2}; cuuint32_t bbox[2] = {C, NOUT}; cuuint32_t bes[2] = {1, 1};
  CUresult r3 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, bdim, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE, (C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 || r2 || r3) { printf("encode failed %d %d %d\n", (int)r1, (int)r2, (int)r3); return 1; }
  const int smem = 23552 + 36864 + 8192 + 1024;
  CK(cudaFuncSetAttribute(shift_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  shift_kernel<C><<<1, 128, smem>>>(md, mr, mb, dout);
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(4 * 9 * 128 * NOUT);
  CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
  printf("channels per pixel row %d (row %d B):\n", C, C * 2);
  const char* names[4] = {"pitch 10 px boff=0", "pitch 16 px boff=0", "pitch 10 px boff=addr", "pitch 16 px boff=addr"};
  for (int v = 0; v < 4; ++v) {
    printf("%-20s:", names[v]);
    for (int tap = 0; tap < 9; ++tap) {
      const int r = tap / 3, s = tap % 3;
      double worst = 0; int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < NOUT; ++n) {
        const int py = m / TW, px = m % TW;
        double ref = 0;
        for (int c = 0; c < C; ++c) ref += (double)fx[((py + r) * HC + px + s) * C + c] * fb[n * C + c];
        const double got = ho[((size_t)(v * 9 + tap) * 128 + m) * NOUT + n];
        const double e = fabs(got - ref);
        if (!(e <= 1e-3)) ++bad;
        if (e > worst || e != e) worst = e;
      }
      printf(" (%d,%d):%s", r, s, bad ? "BAD" : "ok");
      if (bad) printf("[%d]", bad);
    }
    printf("\n");
  }
  return 0;
}

int main() {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  return run<64>(enc) | run<32>(enc) | run<16>(enc);
}
