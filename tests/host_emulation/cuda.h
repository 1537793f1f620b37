// Host stand-in for the slice of the CUDA driver API that fsnet_b200/csrc/conv_tc.cu's HOST code uses (tensor-map encoding), for
// the plan check of tests/host_emulation: cuTensorMapEncodeTiled here VALIDATES its arguments against the driver's documented
// constraints (alignment, stride granularity, box limits, swizzle span) instead of encoding anything.
#pragma once
#include <cstdint>
#include <cstring>

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
static const CUresult CUDA_SUCCESS = 0;
static const CUresult CUDA_ERROR_INVALID_VALUE = 1;
struct alignas(64) CUtensorMap { uint64_t opaque[16]; };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

namespace simt_plan {
struct MapRecord {            // what the last encode asked for (the plan test reads it back)
  int rank;
  uint64_t dims[5], strides[4];
  uint32_t box[5], estr[5];
  int swizzle;
};
inline int n_maps = 0;
inline MapRecord last_map;
}  // namespace simt_plan

static inline CUresult cuTensorMapEncodeTiled(CUtensorMap* map, CUtensorMapDataType dtype, cuuint32_t rank, void* addr, const cuuint64_t* dims,
                                              const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                                              CUtensorMapInterleave, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion,
                                              CUtensorMapFloatOOBfill) {
  const uint64_t esize = dtype == CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 ? 2 : dtype == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 1;
  if (!map || !addr || rank < 1 || rank > 5) return CUDA_ERROR_INVALID_VALUE;
  if (((uintptr_t)addr & 15u) != 0) return CUDA_ERROR_INVALID_VALUE;                       // 16-byte aligned base
  for (cuuint32_t i = 0; i < rank; ++i) {
    if (dims[i] == 0 || dims[i] > (1ull << 32)) return CUDA_ERROR_INVALID_VALUE;
    if (box[i] == 0 || box[i] > 256) return CUDA_ERROR_INVALID_VALUE;
    if (estr[i] == 0 || estr[i] > 8) return CUDA_ERROR_INVALID_VALUE;
  }
  for (cuuint32_t i = 0; i + 1 < rank; ++i)
    if ((strides[i] & 15u) != 0 || strides[i] >= (1ull << 40) || strides[i] == 0) return CUDA_ERROR_INVALID_VALUE;
  const uint64_t inner = (uint64_t)box[0] * esize;
  if ((inner & 15u) != 0) return CUDA_ERROR_INVALID_VALUE;                                  // inner box: multiple of 16 bytes
  const uint64_t span = swizzle == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : swizzle == CU_TENSOR_MAP_SWIZZLE_64B ? 64
                        : swizzle == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : (1ull << 40);
  if (inner > span) return CUDA_ERROR_INVALID_VALUE;                                        // inner box must fit the swizzle span
  simt_plan::MapRecord& r = simt_plan::last_map;
  r.rank = (int)rank;
  r.swizzle = (int)swizzle;
  for (cuuint32_t i = 0; i < rank; ++i) { r.dims[i] = dims[i]; r.box[i] = box[i]; r.estr[i] = estr[i]; if (i + 1 < rank) r.strides[i] = strides[i]; }
  simt_plan::n_maps++;
  std::memset(map, 0, sizeof(*map));
  return CUDA_SUCCESS;
}
