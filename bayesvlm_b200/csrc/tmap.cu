#include "tmap.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace bvlm {

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;

void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr) {
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
}

CUtensorMapDataType to_cu_dtype(int dt) {
  switch (dt) {
    case TM_F16: return CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    case TM_BF16: return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    case TM_U8: return CU_TENSOR_MAP_DATA_TYPE_UINT8;
    default: return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  }
}

}  // namespace

int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle128) {
  std::call_once(g_once, resolve_encode);
  if (g_encode == nullptr) return BVLM_EDRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch_bytes & 15u) != 0) return BVLM_EINVAL;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, to_cu_dtype(dtype), 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BVLM_OK : BVLM_EDRIVER;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle128) {
  std::call_once(g_once, resolve_encode);
  if (g_encode == nullptr) return BVLM_EDRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (s1_bytes & 15u) != 0 || (s2_bytes & 15u) != 0)
    return BVLM_EINVAL;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, to_cu_dtype(dtype), 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BVLM_OK : BVLM_EDRIVER;
}

int64_t operand_pitch(int64_t k_elems) {
  static int pad = -1;
  if (pad < 0) {
    const char* e = getenv("BVLM_PITCH_PAD");
    pad = e != nullptr ? atoi(e) : 64;
    if (pad < 0 || (pad % 8) != 0) pad = 64;
  }
  return ((k_elems * 2) % 1024 == 0) ? k_elems + pad : k_elems;
}

int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev < BVLM_MAX_DEVICES ? dev : BVLM_MAX_DEVICES - 1;
}

int device_sm_count() {
  static std::atomic<int> cached[BVLM_MAX_DEVICES];
  const int slot = current_device_slot();
  if (const int c = cached[slot].load(std::memory_order_relaxed); c > 0) return c;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached[slot].store(n, std::memory_order_relaxed);
  return n;
}

}  // namespace bvlm
