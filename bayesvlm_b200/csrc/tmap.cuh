// Host-side construction of TMA tensor maps. cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so that libbvlm.so has no link-time dependency on libcuda (the library must
// load, and export its symbols, on a machine without a driver; compute calls then fail loudly).
#pragma once
#include "common.cuh"

namespace bvlm {

enum TmapDtype : int { TM_F16 = 0, TM_BF16 = 1, TM_F32 = 2, TM_U8 = 3 };

// 2-D row-major tensor [outer, inner] with row pitch `pitch_bytes`; box = [box_outer, box_inner].
int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle128);

// 3-D tensor [d2, d1, d0] (d0 innermost) with byte strides s1 (between d1 steps) and s2 (between d2 steps).
int make_tmap_3d(CUtensorMap* out, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle128);

// Per-device caches (SM count, occupancy, function attributes) are indexed by the CUDA current device: the library may be
// driven on several GPUs of one process (the Python mirror makes the tensors' device current around every call).
constexpr int BVLM_MAX_DEVICES = 64;
int current_device_slot();  // CUDA current device ordinal, clamped to [0, BVLM_MAX_DEVICES)
int device_sm_count();      // SM count of the current device

}  // namespace bvlm
