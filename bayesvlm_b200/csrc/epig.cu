// E0 / E1 / E2: EPIG acquisition kernels (bayesvlm/epig.py:275-397, bayesvlm/vlm.py:116-123).
//
// The reference evaluates the joint-entropy term in fp16 with a materialised [N_p, Cl, chunk] tile; its scores are
// dominated by where that arithmetic rounds.  The kernels below reproduce those rounding points:
//   joint = fp16( fp16(pool @ targ) / K )             epig.py:387-388   (fp32 accumulate inside the matmul)
//   xl    = fp16( joint * fp16(log(joint)) )          epig.py:390       (torch.xlogy on Half rounds log() to Half first)
//   Hc    = fp16( fp16(-sum_{c,col in chunk} xl) / N_t )   epig.py:391  (fp32 accumulate, one rounding per chunk)
//   Hjoint += Hc  in fp32                              epig.py:381,393
// while the [N_p*Cl, N_t*Cl] joint matrix only ever exists tile-by-tile in tensor memory.
#include "epilogues.cuh"
#include "gemm2_engine.cuh"
#include "prep.cuh"

using namespace bvlm;

namespace {

constexpr int EPIG_BN = 256;

inline int64_t pad64(int64_t k) { return round_up_i64(k, 64); }

__device__ __forceinline__ float round_f16(float v) { return __half2float(__float2half_rn(v)); }

// x * log(x) on a Half tensor AS TORCH'S CUDA KERNEL EVALUATES IT (xlogy_kernel_cuda: `x * std::log(y)` in device code --
// the operands are widened to float, logf and the product stay fp32, ONE rounding to fp16 at the end); 0 at x == 0.
// (torch's CPU kernel rounds log() to Half before the product -- measured on B200 with torch 2.11: the CUDA result equals
//  the single-rounding form for 15359 of the 15360 fp16 values in (0, 1], the CPU form for 11360; scripts/probe_epig_parity.py.)
__device__ __forceinline__ float xlogx_f16(float x) {
  if (x == 0.f) return 0.f;
  return round_f16(x * logf(x));
}

// ---------------------------------------------------------------------------------------------------------------
// E0 + E1 + operand layout in ONE pass over a block of R sample rows                 (vlm.py:116-123, epig.py:294-311, 374-376)
//   FROM_NOISE : eps [K, N, Cl] fp32 (torch.randn order), mean / var [N, Cl]  ->  probs = softmax(eps * sqrt(var) + mean) -> fp16
//   !FROM_NOISE: probs_in [N, K, Cl] fp16
// outputs (each optional): probs_out [N, K, Cl] fp16, oper [N, Cl, Kp] fp16 (the K-major operand of the joint-entropy GEMM,
// zero padded along K), marg [N] fp16 marginal entropies with torch's CUDA rounding points (mean over K = fp32 sum times
// 1/K -> fp16; xlogy -> fp16; sum over Cl in fp32 -> fp16; negate).
//
// Layout of the work (third version.  v1 staged 15 rows in 100 KB with 4-byte cp.async, two 256-thread blocks per SM:
// 0.87 TB/s.  v2 had every thread read its own sample's noise from global memory: 32 cache lines per warp instruction, the
// L1 tag stage at 64 % and the loads latency bound, 39 us for 66 MB; profiles/r2_*):
//   * a block is 128 threads = 128 consecutive MC samples k and owns R = 4 sample rows (28 KB of shared memory at Cl = 10,
//     K = 100: 7 blocks per SM, the fetch of one block overlaps the arithmetic and the stores of the others);
//   * the block's noise is ONE contiguous run of R*Cl floats per sample: cp.async with consecutive threads on consecutive
//     16- / 8- / 4-byte pieces (coalesced), all of the block's bytes in flight at once, into [k][R*Cl + V] rows whose stride
//     makes the per-sample vector reads below bank-conflict free;
//   * thread k evaluates the softmax of its sample for each row in registers (compile-time class count for Cl <= 16) and
//     writes the fp16 probabilities to the shared tile in the OPERAND layout [row][class][k] (lanes along k: conflict free);
//   * the tile is the block's contiguous slice of `oper` and leaves with 128-bit stores; the same 128-bit reads feed the sum
//     over K of the marginal entropy (8 values per thread, 3 shuffles, fixed order: deterministic).
// The softmax follows torch's softmax_warp_forward for rows of <= 16 classes (sum in the 16-lane butterfly order, expf,
// true division), so the fp16 probabilities are bit-identical to the reference's on the same device.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 128;
constexpr int PREP_ROWS = 4;                    // rows per block (compile-time class counts); 4 | R * Cl: 16-byte aligned runs
constexpr size_t PREP_SMEM_MAX = 100 * 1024;    // a block may take up to this much

struct PrepLayout {
  int R, estride;
  size_t off_part, off_ms, off_eps, bytes;
};

// noise vector width (floats) of the per-sample reads for a compile-time class count
constexpr int prep_vec(int cl) { return cl % 4 == 0 ? 4 : cl % 2 == 0 ? 2 : 1; }

inline PrepLayout prep_layout_for(int64_t R, int64_t K, int64_t Cl, int64_t Kp, bool from_noise) {
  PrepLayout L{};
  L.R = static_cast<int>(R);
  // [k][estride] floats.  Cl <= 16 (R = 4): even class counts use estride = 4 Cl + 4 (rows 16-byte aligned for 16-byte
  // cp.async; the 128-bit reads of a quarter warp are conflict free, the 64-bit reads of Cl = 2 mod 4 two-way), odd ones
  // 4 Cl + 1 (scalar reads, conflict free).  Generic path: odd stride, scalar reads.
  L.estride = Cl <= 16 ? static_cast<int>(R * Cl + (Cl % 2 == 0 ? 4 : 1)) : static_cast<int>((R * Cl) | 1);
  L.off_part = static_cast<size_t>(R) * Cl * Kp * 2;                       // float [R*Cl][Kp/64] partial sums over K
  L.off_ms = L.off_part + static_cast<size_t>(R) * Cl * (Kp / 64) * 4;     // float mean[R*Cl], std[R*Cl]
  L.off_eps = (L.off_ms + static_cast<size_t>(R) * Cl * 8 + 15) & ~static_cast<size_t>(15);  // float [K][estride] noise
  L.bytes = L.off_eps + (from_noise ? static_cast<size_t>(K) * L.estride * 4 : 0);
  return L;
}

inline PrepLayout prep_layout(int64_t K, int64_t Cl, int64_t Kp, bool from_noise) {
  PrepLayout L{};
  if (Cl <= 16) {  // register-resident softmax with a compile-time class count: 4, 2 or 1 rows per block, whatever fits
    for (int r : {PREP_ROWS, 2, 1}) {
      L = prep_layout_for(r, K, Cl, Kp, from_noise);
      if (L.bytes <= PREP_SMEM_MAX) return L;
    }
    L.R = 0;
    return L;
  }
  L = prep_layout_for(1, K, Cl, Kp, from_noise);
  if (L.bytes > PREP_SMEM_MAX) L.R = 0;
  return L;
}

__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// CL floats from shared memory with V-float vector loads (the address is V-float aligned by construction)
template <int CL, int V>
__device__ __forceinline__ void load_vec(float (&e)[CL], const float* __restrict__ ep) {
  if constexpr (V == 4) {
#pragma unroll
    for (int c = 0; c < CL; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(ep + c);
      e[c] = v.x, e[c + 1] = v.y, e[c + 2] = v.z, e[c + 3] = v.w;
    }
  } else if constexpr (V == 2) {
#pragma unroll
    for (int c = 0; c < CL; c += 2) {
      const float2 v = *reinterpret_cast<const float2*>(ep + c);
      e[c] = v.x, e[c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CL; ++c) e[c] = ep[c];
  }
}

// softmax(e * std + mean) of one (sample, row) -> fp16, written into the operand tile (stride Kp along the class) and, if
// requested, to the [N, K, Cl] probabilities.  Arithmetic = torch's: randn * std + mean with two roundings, expf, the sum in
// softmax_warp_forward's butterfly order over 16 lanes (missing classes add exact zeros), true division -- evaluated as
// q = z * r, q' = fma(fma(-q, s, z), r, q) with r = RN(1 / s): the correctly rounded quotient (Markstein) in 3 instructions
// per class instead of the ~12 of the generic division sequence with its slow-path check.
template <int CL, bool WANT_PROBS>
__device__ __forceinline__ void softmax_row_f16(const float* __restrict__ ep, const float* __restrict__ m, const float* __restrict__ sd,
                                                __half* __restrict__ s_dst, int Kp, __half* __restrict__ g_dst) {
  constexpr int V = prep_vec(CL);
  float e[CL], mm[CL], ss[CL];
  load_vec<CL, V>(e, ep);
  load_vec<CL, V>(mm, m);   // (row n of the block starts at n * CL floats in all three arrays: same alignment)
  load_vec<CL, V>(ss, sd);
  float z[16];
  float zmax = -INFINITY;
#pragma unroll
  for (int c = 0; c < CL; ++c) {
    z[c] = __fadd_rn(__fmul_rn(e[c], ss[c]), mm[c]);
    zmax = fmaxf(zmax, z[c]);
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) z[c] = c < CL ? expf(z[c] - zmax) : 0.f;
  const float t0 = z[0] + z[8], t1 = z[1] + z[9], t2 = z[2] + z[10], t3 = z[3] + z[11], t4 = z[4] + z[12],
              t5 = z[5] + z[13], t6 = z[6] + z[14], t7 = z[7] + z[15];
  const float u0 = t0 + t4, u1 = t1 + t5, u2 = t2 + t6, u3 = t3 + t7;
  const float zsum = (u0 + u2) + (u1 + u3);  // in [1, 16]
  const float r = __frcp_rn(zsum);
#pragma unroll
  for (int c = 0; c < CL; ++c) {
    const float q = __fmul_rn(z[c], r);
    const __half h = __float2half_rn(__fmaf_rn(__fmaf_rn(-q, zsum, z[c]), r, q));
    s_dst[c * Kp] = h;
    if constexpr (WANT_PROBS) g_dst[c] = h;
  }
}

// CL > 0: compile-time class count (<= 16), R = 4, 2 or 1 rows per block.  CL == 0: any class count, one row per block, three sweeps.
// One side (pool or target set) of a prepare launch.
struct PrepSide {
  const float* mean;
  const float* var;
  const float* eps;
  const __half* probs_in;
  int64_t N;
  __half* probs_out;
  __half* oper;
  __half* marg;
  int piece;  // floats per cp.async of the noise fetch (4, 2 or 1: alignment of this side's runs)
};

// CL > 0: compile-time class count (<= 16), R = 4, 2 or 1 rows per block.  CL == 0: any class count, one row per block, three sweeps.
// Blocks [0, blocks0) work on side a, the rest on side b (EPIG prepares the target set and a pool chunk in ONE launch).
template <bool FROM_NOISE, int CL>
__global__ void __launch_bounds__(PREP_THREADS)
k_epig_prepare(const PrepSide sa, const PrepSide sb, const unsigned blocks0, int K, int Cl, int Kp, const PrepLayout L) {
  extern __shared__ __align__(16) uint8_t ps[];
  const bool second = blockIdx.x >= blocks0;
  const float* __restrict__ mean = second ? sb.mean : sa.mean;
  const float* __restrict__ var = second ? sb.var : sa.var;
  const float* __restrict__ eps = second ? sb.eps : sa.eps;
  const __half* __restrict__ probs_in = second ? sb.probs_in : sa.probs_in;
  const int64_t N = second ? sb.N : sa.N;
  __half* __restrict__ probs_out = second ? sb.probs_out : sa.probs_out;
  __half* __restrict__ oper = second ? sb.oper : sa.oper;
  __half* __restrict__ marg = second ? sb.marg : sa.marg;
  const int piece = second ? sb.piece : sa.piece;
  __half* s_op = reinterpret_cast<__half*>(ps);                // [R][Cl][Kp]
  float* s_part = reinterpret_cast<float*>(ps + L.off_part);   // [R*Cl][Kp/64]
  float* s_mean = reinterpret_cast<float*>(ps + L.off_ms);     // [R*Cl]
  float* s_std = s_mean + L.R * Cl;
  float* s_eps = reinterpret_cast<float*>(ps + L.off_eps);     // [K][estride]: sample k's noise of the block's rows
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x - (second ? blocks0 : 0u)) * L.R;
  const int rows = static_cast<int>(N - n0 < L.R ? N - n0 : L.R);
  const int seg = rows * Cl;
  const __half hzero = __float2half_rn(0.f);

  if constexpr (FROM_NOISE) {
    const float* base = eps + n0 * Cl;
    const int64_t kstride = N * Cl;
    const uint32_t dst0 = smem_u32(s_eps);
    // piece = floats per cp.async (4, 2 or 1), chosen by the host from the alignment of the runs; the tail block's shorter
    // run may not be a multiple of it
    const int pw = (seg % piece) == 0 ? piece : 1;
    const int pieces = seg / pw, total = pieces * K;
    const float inv_pieces = 1.0f / static_cast<float>(pieces);
    for (int i = tid; i < total; i += PREP_THREADS) {
      // k = i / pieces without an integer division: (i + 0.5) / pieces is at least 0.5 / pieces away from an integer and
      // the fp32 error is < 1e-3 for i < 2^17
      const int k = __float2int_rz((static_cast<float>(i) + 0.5f) * inv_pieces), j = (i - k * pieces) * pw;
      const uint32_t d = dst0 + 4u * static_cast<uint32_t>(k * L.estride + j);
      const float* g = base + k * kstride + j;
      if (pw == 4) cp_async_16(d, g);
      else if (pw == 2) cp_async_8(d, g);
      else cp_async_4(d, g);
    }
    for (int j = tid; j < seg; j += PREP_THREADS) {
      s_mean[j] = mean[n0 * Cl + j];
      s_std[j] = sqrtf(var[n0 * Cl + j]);
    }
    cp_async_wait_all();
    __syncthreads();
  }

  // ---- phase 1: one (row, MC sample) per thread and step, lanes along k
  for (int k = tid; k < Kp; k += PREP_THREADS) {
    if (k >= K) {  // zero padding of the operand along K
      for (int i = 0; i < seg; ++i) s_op[i * Kp + k] = hzero;
      continue;
    }
    if constexpr (FROM_NOISE && CL > 0) {
      const float* ek = s_eps + k * L.estride;
      if (probs_out != nullptr) {  // (E0 alone, vlm.py:116-123: the [N, K, Cl] probabilities are an output)
        for (int n = 0; n < rows; ++n)
          softmax_row_f16<CL, true>(ek + n * CL, s_mean + n * CL, s_std + n * CL, s_op + n * CL * Kp + k, Kp,
                                    probs_out + ((n0 + n) * K + k) * CL);
      } else if (rows == PREP_ROWS) {
#pragma unroll
        for (int n = 0; n < PREP_ROWS; ++n)
          softmax_row_f16<CL, false>(ek + n * CL, s_mean + n * CL, s_std + n * CL, s_op + n * CL * Kp + k, Kp, nullptr);
      } else {
        for (int n = 0; n < rows; ++n)
          softmax_row_f16<CL, false>(ek + n * CL, s_mean + n * CL, s_std + n * CL, s_op + n * CL * Kp + k, Kp, nullptr);
      }
    } else if constexpr (FROM_NOISE) {
      for (int n = 0; n < rows; ++n) {
        const float* e = s_eps + k * L.estride + n * Cl;
        const float* m = s_mean + n * Cl;
        const float* sd = s_std + n * Cl;
        float zmax = -INFINITY;
        for (int c = 0; c < Cl; ++c) zmax = fmaxf(zmax, __fadd_rn(__fmul_rn(e[c], sd[c]), m[c]));
        float zsum = 0.f;
        for (int c = 0; c < Cl; ++c) zsum += expf(__fadd_rn(__fmul_rn(e[c], sd[c]), m[c]) - zmax);
        __half* dst_nk = probs_out != nullptr ? probs_out + ((n0 + n) * K + k) * Cl : nullptr;
        for (int c = 0; c < Cl; ++c) {
          const __half h = __float2half_rn(expf(__fadd_rn(__fmul_rn(e[c], sd[c]), m[c]) - zmax) / zsum);
          s_op[(n * Cl + c) * Kp + k] = h;
          if (dst_nk != nullptr) dst_nk[c] = h;
        }
      }
    } else {
      for (int n = 0; n < rows; ++n) {
        const __half* src = probs_in + ((n0 + n) * K + k) * Cl;
        for (int c = 0; c < Cl; ++c) s_op[(n * Cl + c) * Kp + k] = src[c];
      }
    }
  }
  __syncthreads();

  // ---- phase 2: the tile leaves as the block's contiguous slice of `oper` (128-bit stores); the same reads give the sums
  // over K: 8 values per thread, then the 8 threads of a 64-sample group (Kp is a multiple of 64)
  {
    uint4* dst = oper != nullptr ? reinterpret_cast<uint4*>(oper + n0 * Cl * Kp) : nullptr;  // 16-byte aligned: Kp % 64 == 0
    const uint4* src = reinterpret_cast<const uint4*>(s_op);
    const int total = seg * Kp / 8;  // a multiple of 8
    const int groups = Kp / 64;
    for (int i0 = 0; i0 < total; i0 += PREP_THREADS) {
      const int i = i0 + tid;
      float s = 0.f;
      if (i < total) {
        const uint4 v = src[i];
        if (dst != nullptr) dst[i] = v;
        if (marg != nullptr) {
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
          const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
          const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
          s = ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
        }
      }
      if (marg != nullptr) {  // (uniform branch; every lane takes part in the shuffles)
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (i < total && (lane & 7) == 0) s_part[i >> 3] = s;  // [row * groups + group]
      }
    }
    if (marg != nullptr) {
      __syncthreads();
      // ---- phase 3: marginal entropies, one thread per sample row
      const float inv_k = static_cast<float>(1.0 / static_cast<double>(K));  // torch multiplies by the fp32 reciprocal
      if (tid < rows) {
        float ent = 0.f;
        for (int c = 0; c < Cl; ++c) {
          const float* part = s_part + (tid * Cl + c) * groups;
          float s = part[0];
          for (int g = 1; g < groups; ++g) s += part[g];
          ent += xlogx_f16(round_f16(s * inv_k));
        }
        marg[n0 + tid] = __float2half_rn(-round_f16(ent));
      }
    }
  }
}

template <bool FN, int CL>
int launch_prepare_variant(unsigned grid, unsigned blocks0, const PrepLayout& L, const PrepSide& sa, const PrepSide& sb, int K,
                           int Cl, int Kp, cudaStream_t st) {
  auto kfn = k_epig_prepare<FN, CL>;
  if (L.bytes > 48 * 1024)
    BVLM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(PREP_SMEM_MAX)));
  kfn<<<grid, PREP_THREADS, L.bytes, st>>>(sa, sb, blocks0, K, Cl, Kp, L);
  return BVLM_OK;
}

// cp.async piece (floats): every (sample, block) run must start on a piece boundary in global AND shared memory
inline int prep_piece(const float* eps, int64_t N, int64_t Cl, int R, int estride, bool fixed_cl) {
  if (!fixed_cl || eps == nullptr) return 1;
  const uintptr_t a = reinterpret_cast<uintptr_t>(eps);
  for (int p : {4, 2})
    if ((a % (4 * p)) == 0 && (N * Cl) % p == 0 && (static_cast<int64_t>(R) * Cl) % p == 0 && estride % p == 0) return p;
  return 1;
}

// one or two sides (sb.N == 0: one) sharing K and Cl, all from noise or all from probabilities
int launch_epig_prepare(PrepSide sa, PrepSide sb, int64_t K, int64_t Cl, cudaStream_t st) {
  if (sa.N <= 0 && sb.N <= 0) return BVLM_OK;
  if (sa.N <= 0) {
    sa = sb;
    sb.N = 0;
  }
  if (K <= 0 || Cl <= 0 || K > 4096 || Cl > 4096) return BVLM_EINVAL;
  const bool from_noise = sa.eps != nullptr;
  if (sb.N > 0 && (sb.eps != nullptr) != from_noise) return BVLM_EINVAL;
  const int64_t Kp = pad64(K);
  const PrepLayout L = prep_layout(K, Cl, Kp, from_noise);
  if (L.R <= 0) return BVLM_ENOTSUP;  // one row's tiles do not fit shared memory
  const unsigned blocks0 = static_cast<unsigned>(ceil_div_i64(sa.N, L.R));
  const unsigned grid = blocks0 + static_cast<unsigned>(sb.N > 0 ? ceil_div_i64(sb.N, L.R) : 0);
  const bool fixed_cl = from_noise && Cl <= 16;
  sa.piece = prep_piece(sa.eps, sa.N, Cl, L.R, L.estride, fixed_cl);
  sb.piece = prep_piece(sb.eps, sb.N, Cl, L.R, L.estride, fixed_cl);
  const int Ki = static_cast<int>(K), Cli = static_cast<int>(Cl), Kpi = static_cast<int>(Kp);
  int rc = BVLM_OK;
#define BVLM_PREP_CL(C)                                                                              \
  case C:                                                                                            \
    rc = launch_prepare_variant<true, C>(grid, blocks0, L, sa, sb, Ki, Cli, Kpi, st);                \
    break
  timing_begin(TAG_EPIG_PREPARE, st);
  if (fixed_cl) {
    switch (Cli) {
      BVLM_PREP_CL(1); BVLM_PREP_CL(2); BVLM_PREP_CL(3); BVLM_PREP_CL(4); BVLM_PREP_CL(5); BVLM_PREP_CL(6); BVLM_PREP_CL(7);
      BVLM_PREP_CL(8); BVLM_PREP_CL(9); BVLM_PREP_CL(10); BVLM_PREP_CL(11); BVLM_PREP_CL(12); BVLM_PREP_CL(13);
      BVLM_PREP_CL(14); BVLM_PREP_CL(15); BVLM_PREP_CL(16);
      default: rc = BVLM_EINVAL;
    }
  } else if (from_noise) {
    rc = launch_prepare_variant<true, 0>(grid, blocks0, L, sa, sb, Ki, Cli, Kpi, st);
  } else {
    rc = launch_prepare_variant<false, 0>(grid, blocks0, L, sa, sb, Ki, Cli, Kpi, st);
  }
#undef BVLM_PREP_CL
  timing_end(TAG_EPIG_PREPARE, st);
  if (rc) return rc;
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// E2 epilogue. Tile rows are the FLATTENED (pool row, class) axis of the operand [Np * Cl, Kp] -- every lane of every
// 128-row slab carries a real row whatever Cl is (the first version packed floor(128 / Cl) whole pool rows per slab: 51 %
// of the lanes at Cl = 65, and Cl > 128 was not supported); tile columns are the flattened (target, class) axis.  Row
// panels: each CTA pair walks all column tiles of its 256 flat rows.
// Per element (torch's CUDA rounding points; two elements at a time):
//   h = fp16(acc); j = fp16(h * (1/K)); t = fp16(j * log j)  [fp32 log and product, one rounding]; sum += t  (fp32)
// Per (pool row, column chunk) the sum over the row's Cl classes may span slabs, CTAs and CTA pairs: every slab adds its
// share -- summed in double from the per-thread fp32 partials -- to S[pool row, chunk] with one double atomicAdd, and
// k_epig_joint_finalize applies the remaining roundings  Hjoint[p] = sum_chunks fp16( fp16(-S) * (1/N_t) )  in chunk order.
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct EpiEpigJoint {
  static constexpr size_t scratch_bytes(int warps) { return (warps / 4) * 128 * sizeof(float); }
  struct Params {
    double* S;            // [Np, n_chunks], zeroed by the host
    int64_t rows_total;   // Np * Cl
    int Cl;
    int n_chunks;
    int tiles_per_chunk;  // col_chunk / BN
    float inv_K;          // fp32(1 / K): torch's CUDA `tensor / python_scalar` multiplies by the fp32 reciprocal
  };
  struct State {
    float chunk_acc;  // running fp32 sum of fp16(xlogy) over the current column chunk (this thread's row, its column group)
    int64_t p;        // pool row of this thread's flat row
    int head_rows;    // > 0: this thread sums that many consecutive flat rows of pool row p (the part inside this slab)
    int chunk;        // current column chunk
    int seen;         // tiles consumed of the current column chunk
    bool valid;
  };
  static constexpr bool ALL_CHUNKS = false;
  static constexpr bool UNROLL_CHUNKS = false;
  static constexpr bool DRAIN_FIRST = false;
  __device__ static void kernel_begin(State&, const Params&, const EpiCtx&) {}
  __device__ static void kernel_end(State&, const Params&, const EpiCtx&) {}

  __device__ static void flush(State& st, const Params& p, const EpiCtx& ctx) {
    const int r = ctx.ew * 32 + ctx.lane;
    const uint32_t s = ctx.scratch_u32;
    epi_bar_sync(ctx);
    sts_f32(s + 4u * static_cast<uint32_t>((ctx.wid / 4) * 128 + r), st.valid ? st.chunk_acc : 0.f);
    epi_bar_sync(ctx);
    if (st.head_rows > 0) {
      // The fp16 rounding of the chunk sum decides the score; it is accumulated in double (from here to the finalize
      // kernel) so that only the reference's own fp32 summation order can move it across a rounding boundary, not ours.
      double sum = 0.0;
      for (int h = 0; h < ctx.n_warps / 4; ++h)
        for (int c = 0; c < st.head_rows; ++c) sum += static_cast<double>(lds_f32(s + 4u * static_cast<uint32_t>(h * 128 + r + c)));
      atomicAdd(p.S + st.p * p.n_chunks + st.chunk, sum);
    }
    st.chunk_acc = 0.f;
  }

  __device__ static void item_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int r = ctx.ew * 32 + ctx.lane;
    const int64_t flat = static_cast<int64_t>(tc.row0) + r;  // tc.row0: first flat row of this CTA's 128-row slab
    st.valid = flat < p.rows_total;
    st.p = flat / p.Cl;
    const int c = static_cast<int>(flat - st.p * p.Cl);
    st.head_rows = 0;
    if (st.valid && ctx.wid < 4 && (r == 0 || c == 0)) {  // first row of pool row p inside this slab
      int n = p.Cl - c;
      if (n > GEMM_BM - r) n = GEMM_BM - r;
      if (n > p.rows_total - flat) n = static_cast<int>(p.rows_total - flat);
      st.head_rows = n;
    }
    st.chunk_acc = 0.f;
    st.chunk = tc.n / p.tiles_per_chunk;
    st.seen = tc.n - st.chunk * p.tiles_per_chunk;  // a panel starts at tc.n
  }
  __device__ static void tile_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord&) {
    if (st.seen == p.tiles_per_chunk) {  // (a counter, not tc.n / tiles_per_chunk: no division per tile)
      flush(st, p, ctx);
      ++st.chunk;
      st.seen = 0;
    }
    ++st.seen;
  }
  __device__ static void chunk(State& st, const Params& p, const EpiCtx&, const TileCoord&, float (&v)[32], int) {
    // columns beyond N are TMA zero fill: j = 0 -> 0 * log 0 = NaN, which the NaN-suppressing min below turns into 0
    const __half2 zero2 = __float2half2_rn(0.f);
    const float2 inv_k2 = make_float2(p.inv_K, p.inv_K), ln2 = make_float2(0.6931471805599453f, 0.6931471805599453f);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      // packed fp32x2 multiplies (FMUL2) and mixed-precision adds (fp32 += fp16, FHADD): same roundings, fewer issue slots
      const float2 h = __half22float2(__floats2half2_rn(v[j], v[j + 1]));            // matmul output rounded to fp16
      const float2 jf = __half22float2(__float22half2_rn(__fmul2_rn(h, inv_k2)));    // "/ K" on a Half tensor
      const float2 lg = __fmul2_rn(make_float2(fast_log2(jf.x), fast_log2(jf.y)), ln2);  // log j in fp32
      __half2 t = __float22half2_rn(__fmul2_rn(jf, lg));                              // fp16(j * log j): ONE rounding
      t = __hmin2(t, zero2);                                                          // j log j <= 0; NaN (j = 0) -> 0
      s0 = add_f32_f16(s0, __low2half(t));
      s1 = add_f32_f16(s1, __high2half(t));
    }
    st.chunk_acc += s0 + s1;
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State& st, const Params& p, const EpiCtx& ctx, const TileCoord&) { flush(st, p, ctx); }
};

// Hjoint[p] = sum over chunks (in order, fp32) of fp16( fp16(-S[p, chunk]) * fp32(1 / N_t) )        (epig.py:391-393)
__global__ void k_epig_joint_finalize(const double* __restrict__ S, int64_t Np, int n_chunks, float inv_Nt,
                                      float* __restrict__ Hjoint) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  float hj = 0.f;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const float neg = -round_f16(static_cast<float>(S[p * n_chunks + ch]));  // fp16(sum) then negate
    hj += round_f16(neg * inv_Nt);                                             // "/ N_t" on a Half tensor
  }
  Hjoint[p] = hj;
}

}  // namespace

extern "C" {

int bvlm_epig_operand_k(int64_t K) { return static_cast<int>(pad64(K)); }

int bvlm_epig_prepare_supported(int64_t K, int64_t Cl, int from_noise) {
  if (K <= 0 || Cl <= 0 || K > 4096 || Cl > 4096) return 0;
  return prep_layout(K, Cl, pad64(K), from_noise != 0).R > 0 ? 1 : 0;
}

int bvlm_epig_prepare_from_noise(const float* mean, const float* var, const float* eps, int64_t N, int64_t K, int64_t Cl,
                                 void* probs16, void* oper16, void* marg16, void* stream) {
  if (mean == nullptr || var == nullptr || eps == nullptr) return BVLM_EINVAL;
  if (probs16 == nullptr && oper16 == nullptr && marg16 == nullptr) return BVLM_EINVAL;
  const PrepSide sa{mean, var, eps, nullptr, N, static_cast<__half*>(probs16), static_cast<__half*>(oper16),
                    static_cast<__half*>(marg16), 1};
  return launch_epig_prepare(sa, PrepSide{}, K, Cl, static_cast<cudaStream_t>(stream));
}

int bvlm_epig_prepare_pair_from_noise(const float* mean_a, const float* var_a, const float* eps_a, int64_t Na, void* oper_a,
                                      void* marg_a, const float* mean_b, const float* var_b, const float* eps_b, int64_t Nb,
                                      void* oper_b, void* marg_b, int64_t K, int64_t Cl, void* stream) {
  if (mean_a == nullptr || var_a == nullptr || eps_a == nullptr || mean_b == nullptr || var_b == nullptr || eps_b == nullptr)
    return BVLM_EINVAL;
  if ((oper_a == nullptr && marg_a == nullptr) || (oper_b == nullptr && marg_b == nullptr)) return BVLM_EINVAL;
  const PrepSide sa{mean_a, var_a, eps_a, nullptr, Na, nullptr, static_cast<__half*>(oper_a), static_cast<__half*>(marg_a), 1};
  const PrepSide sb{mean_b, var_b, eps_b, nullptr, Nb, nullptr, static_cast<__half*>(oper_b), static_cast<__half*>(marg_b), 1};
  return launch_epig_prepare(sa, sb, K, Cl, static_cast<cudaStream_t>(stream));
}

int bvlm_epig_prepare_from_probs(const void* probs16, int64_t N, int64_t K, int64_t Cl, void* oper16, void* marg16,
                                 void* stream) {
  if (probs16 == nullptr || (oper16 == nullptr && marg16 == nullptr)) return BVLM_EINVAL;
  const PrepSide sa{nullptr, nullptr, nullptr, static_cast<const __half*>(probs16), N, nullptr, static_cast<__half*>(oper16),
                    static_cast<__half*>(marg16), 1};
  return launch_epig_prepare(sa, PrepSide{}, K, Cl, static_cast<cudaStream_t>(stream));
}

int bvlm_epig_sample_probs(const float* mean, const float* var, const float* eps, int64_t N, int64_t K, int64_t Cl,
                           void* probs16, void* stream) {
  if (probs16 == nullptr) return BVLM_EINVAL;
  return bvlm_epig_prepare_from_noise(mean, var, eps, N, K, Cl, probs16, nullptr, nullptr, stream);
}

int bvlm_epig_marginal_entropy_f16(const void* probs16, int64_t N, int64_t K, int64_t Cl, void* out16, void* stream) {
  if (out16 == nullptr) return BVLM_EINVAL;
  return bvlm_epig_prepare_from_probs(probs16, N, K, Cl, nullptr, out16, stream);
}

size_t bvlm_epig_joint_operands_workspace_bytes(int64_t Np, int64_t Nt, int64_t Cl, int64_t col_chunk) {
  if (Np <= 0 || Nt <= 0 || Cl <= 0 || col_chunk <= 0) return 0;
  const int64_t n_chunks = ceil_div_i64(Nt * Cl, col_chunk);
  return static_cast<size_t>(Np * n_chunks) * sizeof(double) + 256;
}

int bvlm_epig_joint_entropy_operands(const void* poolP, int64_t Np, const void* targP, int64_t Nt, int64_t K, int64_t Cl,
                                     int64_t col_chunk, float* Hjoint, void* ws, size_t ws_bytes, void* stream) {
  if (poolP == nullptr || targP == nullptr || Hjoint == nullptr || ws == nullptr) return BVLM_EINVAL;
  if (Np <= 0 || Nt <= 0 || K <= 0 || Cl <= 0) return BVLM_EINVAL;
  if (col_chunk <= 0 || (col_chunk % EPIG_BN) != 0) return BVLM_ENOTSUP;
  if (Nt * Cl > 0x7fffffff || Np * Cl > 0x7fffffff) return BVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(poolP) & 15) != 0 || (reinterpret_cast<uintptr_t>(targP) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(ws) & 7) != 0)
    return BVLM_EINVAL;
  if (ws_bytes < bvlm_epig_joint_operands_workspace_bytes(Np, Nt, Cl, col_chunk)) return BVLM_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t Kp = pad64(K);
  const int n_chunks = static_cast<int>(ceil_div_i64(Nt * Cl, col_chunk));
  double* S = static_cast<double*>(ws);
  BVLM_CUDA_TRY(cudaMemsetAsync(S, 0, static_cast<size_t>(Np) * n_chunks * sizeof(double), st));
  CUtensorMap tmA, tmB;
  int rc;
  Operand16 opA{poolP, Np * Cl, Kp, FMT_F16};
  Operand16 opB{targP, Nt * Cl, Kp, FMT_F16};
  if ((rc = operand_tmap<GEMM_BM>(&tmA, opA))) return rc;
  if ((rc = operand_tmap<EPIG_BN / 2>(&tmB, opB))) return rc;  // CTA pairs: each CTA loads half of the B tile
  GemmPlan plan = make_plan2<EPIG_BN>(static_cast<int>(Np * Cl), static_cast<int>(Nt * Cl), static_cast<int>(Kp), SCHED_ROW_PANEL, 1,
                                      FMT_F16);
  // Row panels are cut into column ranges so that the items fill whole rounds of the CTA pairs: a 4096-row pool chunk at
  // Cl = 10 is 160 panels = 2.16 rounds of 74 pairs (72 % of the machine busy); 6 ranges each make 960 items = 12.97 rounds.
  // (Ranges need not be chunk aligned: a (pool row, chunk) sum is the atomic sum of whatever items cover it.)
  plan.splits = balanced_panel_splits(plan.m_tiles, plan.n_tiles, device_sm_count() / 2, 16);
  EpiEpigJoint<EPIG_BN>::Params ep{S, Np * Cl, static_cast<int>(Cl), n_chunks, static_cast<int>(col_chunk / EPIG_BN),
                                   static_cast<float>(1.0 / static_cast<double>(K))};
  if ((rc = launch_gemm2<EPIG_BN, 6, 16, EpiEpigJoint<EPIG_BN>>(tmA, tmB, plan, ep, st, TAG_EPIG_JOINT))) return rc;
  k_epig_joint_finalize<<<static_cast<unsigned>(ceil_div_i64(Np, 256)), 256, 0, st>>>(
      S, Np, n_chunks, static_cast<float>(1.0 / static_cast<double>(Nt)), Hjoint);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

size_t bvlm_epig_joint_workspace_bytes(int64_t Np, int64_t Nt, int64_t K, int64_t Cl) {
  const int64_t Kp = pad64(K);
  // operands + the per-(pool row, chunk) sums at the smallest supported chunk (one column tile)
  return static_cast<size_t>(round_up_i64(Np * Cl * Kp * 2, 256) + round_up_i64(Nt * Cl * Kp * 2, 256) + 1024) +
         bvlm_epig_joint_operands_workspace_bytes(Np, Nt, Cl, EPIG_BN);
}

int bvlm_epig_joint_entropy_f16(const void* pool16, int64_t Np, const void* targ16, int64_t Nt, int64_t K, int64_t Cl,
                                int64_t col_chunk, float* Hjoint, void* ws, size_t ws_bytes, void* stream) {
  if (pool16 == nullptr || targ16 == nullptr || Hjoint == nullptr || ws == nullptr) return BVLM_EINVAL;
  if (Np <= 0 || Nt <= 0 || K <= 0 || Cl <= 0) return BVLM_EINVAL;
  if (ws_bytes < bvlm_epig_joint_workspace_bytes(Np, Nt, K, Cl)) return BVLM_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return BVLM_EINVAL;
  const int64_t Kp = pad64(K);
  uint8_t* base = static_cast<uint8_t*>(ws);
  void* poolP = base;
  void* targP = base + round_up_i64(Np * Cl * Kp * 2, 256);
  void* sums = base + round_up_i64(Np * Cl * Kp * 2, 256) + round_up_i64(Nt * Cl * Kp * 2, 256);
  const size_t sums_bytes = ws_bytes - static_cast<size_t>(static_cast<uint8_t*>(sums) - base);
  int rc = bvlm_epig_prepare_from_probs(pool16, Np, K, Cl, poolP, nullptr, stream);
  if (rc) return rc;
  if ((rc = bvlm_epig_prepare_from_probs(targ16, Nt, K, Cl, targP, nullptr, stream))) return rc;
  return bvlm_epig_joint_entropy_operands(poolP, Np, targP, Nt, K, Cl, col_chunk, Hjoint, sums, sums_bytes, stream);
}

}  // extern "C"
