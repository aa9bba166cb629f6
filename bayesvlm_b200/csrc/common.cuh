// Thin inline-PTX wrappers for the sm_100a primitives the Laplace hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and fences.
// Everything here is device-side plumbing; no torch types.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>

#include "../../include/bvlm.h"

namespace bvlm {

// ---------------------------------------------------------------------------------------------
// status codes of the C-ABI (include/bvlm.h): 0 ok, <0 invalid argument, >0 cudaError_t
// ---------------------------------------------------------------------------------------------
// (the BVLM_* status macros come from include/bvlm.h)

#define BVLM_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return static_cast<int>(_e);       \
  } while (0)

// kernel-launch counter (bench.py reports it as gpu_launches); defined in api.cu
void count_launch(int n = 1);

// Optional per-kernel CUDA-event timing of the tensor-core launches (bench.py's live roofline numerator).
// Disabled by default: timing_begin/timing_end are no-ops unless bvlm_timing_enable(1) was called.
enum KernelTag : int {
  TAG_GEMM_DIAG = 0,
  TAG_SYRK = 1,
  TAG_GGN_ROWSTATS = 2,
  TAG_GGN_WEIGHTS = 3,
  TAG_GGN_MOMENTS = 4,
  TAG_GGN_STACKED = 5,
  TAG_QUADFORM = 6,
  TAG_PREDICTIVE = 7,
  TAG_EPIG_JOINT = 8,
  TAG_EPIG_PREPARE = 9,
  TAG_PRED_PREP = 10,
  TAG_PROBIT = 11,
  TAG_SYRK_PREP = 12,
  TAG_COUNT = 13
};
void timing_begin(int tag, cudaStream_t st);
void timing_end(int tag, cudaStream_t st);

__host__ __device__ constexpr inline int64_t ceil_div_i64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ constexpr inline int64_t round_up_i64(int64_t a, int64_t b) { return ceil_div_i64(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// shared-memory address helper
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware instead of spinning
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap instead of hanging the GPU (≈ 2–3 s at 1.3–1.9 GHz).
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // The poll loop runs next to the epilogue warps of the same scheduler for the whole kernel (a producer / MMA thread of an
  // epilogue-bound kernel waits almost all the time): every instruction in it is an issue slot taken from them.  ncu
  // (profiles/r1e EPIG capture) attributed ~12 % of all issued instructions to the previous clock64()-based timeout
  // check, so the bound is an iteration counter (each failed try_wait already sleeps in hardware for its time hint).
  // BACKOFF: waits that are off the critical path (a producer waiting for a free stage, the MMA thread waiting for a drained
  // accumulator) sleep 128 ns between polls after the first few -- one poll per ~250 cycles instead of one per ~60.
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (BACKOFF && spins >= 4) __nanosleep(128);
    if (++spins == (1u << 26)) {  // >= 64 Mi failed polls of >= ~60 cycles each: seconds, i.e. a protocol bug
      printf("bvlm: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
// plain (non-tensor) bulk copy global -> shared of `bytes` (multiple of 16, both addresses 16-byte aligned), completing on `bar`
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (its CTAs
// become resident, run their prologue) as soon as every CTA of the preceding kernel has executed launch_dependents (or exited);
// it must execute grid_dependency_wait() before touching anything the preceding kernel wrote.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// explicit shared-memory accesses (the compiler cannot prove the address space of pointers carved out of the
// dynamic smem block and would emit generic LD.E / ST.E)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts_v4_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float a) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; kind::f16 covers fp16 and bf16 operands with fp32 accumulate.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout follows the PTX ISA "shared memory matrix descriptor" and
// "instruction descriptor" tables for tcgen05.mma; K-major operand tiles in 128-byte swizzle).
// ---------------------------------------------------------------------------------------------
// K-major, SWIZZLE_128B tile: rows are 128 bytes (64 x 16-bit), 8-row groups are 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                       // layout type: SWIZZLE_128B
  return d;
}

// MN-major, SWIZZLE_128B tile: the operand's M (or N) index is contiguous in memory. Shared memory holds 64-element
// (128-byte) wide column blocks of [K rows x 128 bytes]; inside a block 8-row groups are 1024 bytes apart (SBO), blocks
// are `block_bytes` apart (LBO). Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t block_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address
  d |= static_cast<uint64_t>((block_bytes >> 4) & 0x3FFFu) << 16;    // leading byte offset: next 64-wide MN block
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // stride byte offset: next 8 K rows
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                               // layout type: SWIZZLE_128B
  return d;
}

enum OperandFormat : int { FMT_F16 = 0, FMT_BF16 = 1 };

// kind::f16 instruction descriptor: fp32 accumulate, both operands K-major.
__host__ __device__ inline uint32_t make_idesc_f16(int m, int n, int a_fmt, int b_fmt, int a_mn_major = 0,
                                                   int b_mn_major = 0) {
  uint32_t d = 0;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;  // A major-ness: 0 = K-major, 1 = MN-major
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;  // B major-ness
  d |= 1u << 4;                                  // D format: F32
  d |= static_cast<uint32_t>(a_fmt & 7) << 7;    // A format
  d |= static_cast<uint32_t>(b_fmt & 7) << 10;   // B format
  d |= static_cast<uint32_t>(n >> 3) << 17;      // N >> 3
  d |= static_cast<uint32_t>(m >> 4) << 24;      // M >> 4
  return d;
}

// kind::f8f6f4 instruction descriptor for two E4M3 operands (format code 0), fp32 accumulate, K-major.
__host__ __device__ inline uint32_t make_idesc_e4m3(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                              // D format: F32
  d |= static_cast<uint32_t>(n >> 3) << 17;  // N >> 3
  d |= static_cast<uint32_t>(m >> 4) << 24;  // M >> 4
  return d;
}

// ---------------------------------------------------------------------------------------------
// small math / memory helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fp32 + fp16 -> fp32 in one instruction (sm_100 mixed-precision add, SASS FHADD): the half is converted exactly
__device__ __forceinline__ float add_f32_f16(float acc, __half h) {
  asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc) : "h"(__half_as_ushort(h)));
  return acc;
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

}  // namespace bvlm
