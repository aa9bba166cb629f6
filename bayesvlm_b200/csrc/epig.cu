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
// Layout of the work (second version; the first staged the noise of 15 rows in 100 KB of shared memory with 4-byte cp.async,
// two 256-thread blocks per SM and strictly serial fetch / compute / store phases: 0.87 TB/s, bench r2 `epig.reductions`):
//   * a block is 128 threads = 128 consecutive MC samples k, and owns R (1..8) sample rows; it needs only the R*Cl*Kp fp16
//     operand tile in shared memory (10 KB at Cl = 10, K = 100), so 10+ blocks are resident per SM and the fetch of one block
//     overlaps the arithmetic and the stores of the others;
//   * thread k reads ITS Cl noise values of a row straight from global memory into registers (the R rows of a block are one
//     contiguous run per sample, 64-/128-bit loads when Cl is even / a multiple of 4; the next row is fetched before the current
//     one is evaluated), so no noise is staged anywhere;
//   * the probabilities go to the shared tile in the OPERAND layout [row][class][k] (lanes along k: conflict free); the tile is
//     the block's contiguous slice of `oper` and leaves with 128-bit stores, and the same 128-bit reads feed the sum over K of
//     the marginal entropy (8 values per thread, 3 shuffles, fixed order: deterministic).
// The softmax follows torch's softmax_warp_forward for rows of <= 16 classes (sum in the 16-lane butterfly order, expf,
// true division), so the fp16 probabilities are bit-identical to the reference's on the same device.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 128;
constexpr size_t PREP_SMEM_TARGET = 12 * 1024;  // per block when more than one row fits: >= 10 resident blocks per SM
constexpr size_t PREP_SMEM_MAX = 100 * 1024;    // a single row's tile may take up to this much

struct PrepLayout {
  int R;
  size_t off_part, off_ms, bytes;
};

inline size_t prep_row_bytes(int64_t Cl, int64_t Kp) {
  return static_cast<size_t>(Cl * Kp * 2) + static_cast<size_t>(Cl) * (Kp / 64) * 4 + static_cast<size_t>(Cl) * 8;
}

inline PrepLayout prep_layout(int64_t Cl, int64_t Kp) {
  PrepLayout L{};
  const size_t per_row = prep_row_bytes(Cl, Kp);
  int R = static_cast<int>(PREP_SMEM_TARGET / per_row);
  if (R > 8) R = 8;
  if (R < 1) R = per_row <= PREP_SMEM_MAX ? 1 : 0;
  L.R = R;
  if (R <= 0) return L;
  L.off_part = static_cast<size_t>(R) * Cl * Kp * 2;                       // float [R*Cl][Kp/64] partial sums over K
  L.off_ms = L.off_part + static_cast<size_t>(R) * Cl * (Kp / 64) * 4;     // float mean[R*Cl], std[R*Cl]
  L.bytes = L.off_ms + static_cast<size_t>(R) * Cl * 8;
  return L;
}

// Cl noise values of one (sample, row) into registers; VEC = floats per load (alignment guaranteed by the caller)
template <int VEC>
__device__ __forceinline__ void load_noise16(float (&e)[16], const float* __restrict__ src, int Cl) {
  if constexpr (VEC == 4) {
#pragma unroll
    for (int c = 0; c < 16; c += 4)
      if (c < Cl) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
        e[c] = v.x, e[c + 1] = v.y, e[c + 2] = v.z, e[c + 3] = v.w;
      }
  } else if constexpr (VEC == 2) {
#pragma unroll
    for (int c = 0; c < 16; c += 2)
      if (c < Cl) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src + c));
        e[c] = v.x, e[c + 1] = v.y;
      }
  } else {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < Cl) e[c] = __ldg(src + c);
  }
}

template <bool FROM_NOISE, bool CL16, int VEC>
__global__ void __launch_bounds__(PREP_THREADS)
k_epig_prepare(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ eps,
               const __half* __restrict__ probs_in, int64_t N, int K, int Cl, int Kp, const PrepLayout L,
               __half* __restrict__ probs_out, __half* __restrict__ oper, __half* __restrict__ marg) {
  extern __shared__ __align__(16) uint8_t ps[];
  __half* s_op = reinterpret_cast<__half*>(ps);                // [R][Cl][Kp]
  float* s_part = reinterpret_cast<float*>(ps + L.off_part);   // [R*Cl][Kp/64]
  float* s_mean = reinterpret_cast<float*>(ps + L.off_ms);     // [R*Cl]
  float* s_std = s_mean + L.R * Cl;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * L.R;
  const int rows = static_cast<int>(N - n0 < L.R ? N - n0 : L.R);
  const int seg = rows * Cl;
  const __half hzero = __float2half_rn(0.f);

  if constexpr (FROM_NOISE) {
    for (int j = tid; j < seg; j += PREP_THREADS) {
      s_mean[j] = mean[n0 * Cl + j];
      s_std[j] = sqrtf(var[n0 * Cl + j]);
    }
    __syncthreads();
  }

  // ---- phase 1: one (row, MC sample) per thread and step, lanes along k
  for (int k = tid; k < Kp; k += PREP_THREADS) {
    if (k >= K) {  // zero padding of the operand along K
      for (int i = 0; i < seg; ++i) s_op[i * Kp + k] = hzero;
      continue;
    }
    if constexpr (FROM_NOISE && CL16) {
      const float* src = eps + (static_cast<int64_t>(k) * N + n0) * Cl;
      float e[16], en[16];
      load_noise16<VEC>(e, src, Cl);
      for (int n = 0; n < rows; ++n) {
        if (n + 1 < rows) load_noise16<VEC>(en, src + (n + 1) * Cl, Cl);  // next row in flight during this row's math
        const float* m = s_mean + n * Cl;
        const float* sd = s_std + n * Cl;
        float z[16];
        float zmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          z[c] = c < Cl ? __fadd_rn(__fmul_rn(e[c], sd[c]), m[c]) : -INFINITY;  // torch: randn * std + mean, two roundings
          zmax = fmaxf(zmax, z[c]);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) z[c] = c < Cl ? expf(z[c] - zmax) : 0.f;
        // torch softmax_warp_forward: butterfly sum over the 16 lanes of a row (xor 8, 4, 2, 1)
        const float t0 = z[0] + z[8], t1 = z[1] + z[9], t2 = z[2] + z[10], t3 = z[3] + z[11], t4 = z[4] + z[12],
                    t5 = z[5] + z[13], t6 = z[6] + z[14], t7 = z[7] + z[15];
        const float u0 = t0 + t4, u1 = t1 + t5, u2 = t2 + t6, u3 = t3 + t7;
        const float zsum = (u0 + u2) + (u1 + u3);
        __half* dst_nk = probs_out != nullptr ? probs_out + ((n0 + n) * K + k) * Cl : nullptr;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < Cl) {
            const __half h = __float2half_rn(z[c] / zsum);
            s_op[(n * Cl + c) * Kp + k] = h;
            if (dst_nk != nullptr) dst_nk[c] = h;
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) e[c] = en[c];
      }
    } else if constexpr (FROM_NOISE) {
      for (int n = 0; n < rows; ++n) {
        const float* e = eps + (static_cast<int64_t>(k) * N + n0 + n) * Cl;  // re-read from L1 in the three sweeps
        const float* m = s_mean + n * Cl;
        const float* sd = s_std + n * Cl;
        float zmax = -INFINITY;
        for (int c = 0; c < Cl; ++c) zmax = fmaxf(zmax, __fadd_rn(__fmul_rn(__ldg(e + c), sd[c]), m[c]));
        float zsum = 0.f;
        for (int c = 0; c < Cl; ++c) zsum += expf(__fadd_rn(__fmul_rn(__ldg(e + c), sd[c]), m[c]) - zmax);
        __half* dst_nk = probs_out != nullptr ? probs_out + ((n0 + n) * K + k) * Cl : nullptr;
        for (int c = 0; c < Cl; ++c) {
          const __half h = __float2half_rn(expf(__fadd_rn(__fmul_rn(__ldg(e + c), sd[c]), m[c]) - zmax) / zsum);
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

int launch_epig_prepare(const float* mean, const float* var, const float* eps, const __half* probs_in, int64_t N, int64_t K,
                        int64_t Cl, __half* probs_out, __half* oper, __half* marg, cudaStream_t st) {
  if (N <= 0) return BVLM_OK;
  if (K <= 0 || Cl <= 0 || K > 4096 || Cl > 4096) return BVLM_EINVAL;
  const bool from_noise = eps != nullptr;
  const int64_t Kp = pad64(K);
  const PrepLayout L = prep_layout(Cl, Kp);
  if (L.R <= 0) return BVLM_ENOTSUP;  // one row's Cl x Kp tile does not fit shared memory
  const unsigned grid = static_cast<unsigned>(ceil_div_i64(N, L.R));
  const bool cl16 = Cl <= 16;
  // vector width of the noise loads: every (sample, row) run starts at a multiple of Cl floats
  const bool base16 = (reinterpret_cast<uintptr_t>(eps) & 15) == 0;
  const int vec = !from_noise || !cl16 ? 1 : (base16 && Cl % 4 == 0) ? 4 : (base16 && Cl % 2 == 0) ? 2 : 1;
#define BVLM_PREP_LAUNCH(FN, C16, VEC)                                                                                  \
  do {                                                                                                                  \
    auto kfn = k_epig_prepare<FN, C16, VEC>;                                                                            \
    if (L.bytes > 48 * 1024)                                                                                            \
      BVLM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(PREP_SMEM_MAX))); \
    kfn<<<grid, PREP_THREADS, L.bytes, st>>>(mean, var, eps, probs_in, N, static_cast<int>(K), static_cast<int>(Cl),    \
                                             static_cast<int>(Kp), L, probs_out, oper, marg);                           \
  } while (0)
  timing_begin(TAG_EPIG_PREPARE, st);
  if (from_noise) {
    if (cl16 && vec == 4) BVLM_PREP_LAUNCH(true, true, 4);
    else if (cl16 && vec == 2) BVLM_PREP_LAUNCH(true, true, 2);
    else if (cl16) BVLM_PREP_LAUNCH(true, true, 1);
    else BVLM_PREP_LAUNCH(true, false, 1);
  } else {
    BVLM_PREP_LAUNCH(false, false, 1);
  }
#undef BVLM_PREP_LAUNCH
  timing_end(TAG_EPIG_PREPARE, st);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// E2 epilogue. Tile rows are (pool row, class) pairs packed ppt = floor(128/Cl) pool rows per 128-row CTA slab; tile
// columns are the flattened (target, class) axis. Row panels: each CTA pair walks all column tiles of its pool rows.
// Per element (torch's CUDA rounding points; two elements at a time):
//   h = fp16(acc); j = fp16(h * (1/K)); t = fp16(j * log j)  [fp32 log and product, one rounding]; sum += t  (fp32)
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct EpiEpigJoint {
  static constexpr size_t scratch_bytes(int warps) { return (warps / 4) * 128 * sizeof(float); }
  struct Params {
    float* Hjoint;      // [Np]
    int64_t Np;
    int Cl;
    int ppt;            // pool rows per 128-row slab
    int tiles_per_chunk;  // col_chunk / BN
    float inv_K;        // fp32(1 / K): torch's CUDA `tensor / python_scalar` multiplies by the fp32 reciprocal
    float inv_Nt;       // fp32(1 / N_t), same rule
  };
  struct State {
    float chunk_acc;  // running fp32 sum of fp16(xlogy) over the current column chunk (this thread's row, its column half)
    float hj;         // accumulated joint entropy of the pool row (held by the class-0 thread of the lower column half)
    bool valid;
    bool leader;
    int64_t p;
    int cur_chunk;
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
    if (st.leader) {
      // The fp16 rounding of this sum decides the score; it is accumulated in double (Cl * column groups addends of up to
      // ~1e3, once per chunk and pool row) so that only the reference's own fp32 summation order can move it across a
      // rounding boundary, not ours as well.
      double sum = 0.0;
      for (int h = 0; h < ctx.n_warps / 4; ++h)
        for (int c = 0; c < p.Cl; ++c) sum += static_cast<double>(lds_f32(s + 4u * static_cast<uint32_t>(h * 128 + r + c)));
      const float neg = -round_f16(static_cast<float>(sum));  // fp16(sum) then negate
      st.hj += round_f16(neg * p.inv_Nt);              // "/ N_t" on a Half tensor
    }
    st.chunk_acc = 0.f;
  }

  __device__ static void item_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int r = ctx.ew * 32 + ctx.lane;
    const int pl = r / p.Cl;
    st.p = static_cast<int64_t>(tc.row0 / GEMM_BM) * p.ppt + pl;
    st.valid = (pl < p.ppt) && (st.p < p.Np);
    st.leader = st.valid && (r - pl * p.Cl == 0) && ctx.wid < 4;
    st.chunk_acc = 0.f;
    st.hj = 0.f;
    st.cur_chunk = 0;
  }
  __device__ static void tile_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int ch = tc.n / p.tiles_per_chunk;
    if (ch != st.cur_chunk) {
      flush(st, p, ctx);
      st.cur_chunk = ch;
    }
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
  __device__ static void item_end(State& st, const Params& p, const EpiCtx& ctx, const TileCoord&) {
    flush(st, p, ctx);
    if (st.leader) p.Hjoint[st.p] = st.hj;
  }
};

}  // namespace

extern "C" {

int bvlm_epig_operand_k(int64_t K) { return static_cast<int>(pad64(K)); }

int bvlm_epig_prepare_from_noise(const float* mean, const float* var, const float* eps, int64_t N, int64_t K, int64_t Cl,
                                 void* probs16, void* oper16, void* marg16, void* stream) {
  if (mean == nullptr || var == nullptr || eps == nullptr) return BVLM_EINVAL;
  if (probs16 == nullptr && oper16 == nullptr && marg16 == nullptr) return BVLM_EINVAL;
  return launch_epig_prepare(mean, var, eps, nullptr, N, K, Cl, static_cast<__half*>(probs16), static_cast<__half*>(oper16),
                             static_cast<__half*>(marg16), static_cast<cudaStream_t>(stream));
}

int bvlm_epig_prepare_from_probs(const void* probs16, int64_t N, int64_t K, int64_t Cl, void* oper16, void* marg16,
                                 void* stream) {
  if (probs16 == nullptr || (oper16 == nullptr && marg16 == nullptr)) return BVLM_EINVAL;
  return launch_epig_prepare(nullptr, nullptr, nullptr, static_cast<const __half*>(probs16), N, K, Cl, nullptr,
                             static_cast<__half*>(oper16), static_cast<__half*>(marg16), static_cast<cudaStream_t>(stream));
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

int bvlm_epig_joint_entropy_operands(const void* poolP, int64_t Np, const void* targP, int64_t Nt, int64_t K, int64_t Cl,
                                     int64_t col_chunk, float* Hjoint, void* stream) {
  if (poolP == nullptr || targP == nullptr || Hjoint == nullptr) return BVLM_EINVAL;
  if (Np <= 0 || Nt <= 0 || K <= 0 || Cl <= 0) return BVLM_EINVAL;
  if (Cl > 128 || col_chunk <= 0 || (col_chunk % EPIG_BN) != 0) return BVLM_ENOTSUP;
  if (Nt * Cl > 0x7fffffff || Np > 0x7fffffff) return BVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(poolP) & 15) != 0 || (reinterpret_cast<uintptr_t>(targP) & 15) != 0) return BVLM_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t Kp = pad64(K);
  const int ppt = static_cast<int>(128 / Cl);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_3d(&tmA, poolP, TM_F16, static_cast<uint64_t>(Kp), static_cast<uint64_t>(Cl),
                        static_cast<uint64_t>(Np), static_cast<uint64_t>(Kp) * 2, static_cast<uint64_t>(Cl * Kp) * 2,
                        GEMM_BK, static_cast<uint32_t>(Cl), static_cast<uint32_t>(ppt), 1);
  if (rc) return rc;
  Operand16 opB{targP, Nt * Cl, Kp, FMT_F16};
  if ((rc = operand_tmap<EPIG_BN / 2>(&tmB, opB))) return rc;  // CTA pairs: each CTA loads half of the B tile
  const int m_tiles = static_cast<int>(ceil_div_i64(Np, 2 * ppt));  // a CTA pair covers 2 * ppt pool rows
  GemmPlan plan = make_plan2<EPIG_BN>(m_tiles * GEMM2_BM, static_cast<int>(Nt * Cl), static_cast<int>(Kp), SCHED_ROW_PANEL, 1,
                                      FMT_F16);
  plan.m_tiles = m_tiles;
  plan.a_is_3d = 1;
  plan.a_outer_step = ppt;
  plan.a_tx_bytes = static_cast<uint32_t>(ppt * Cl * GEMM_BK * 2);
  EpiEpigJoint<EPIG_BN>::Params ep{Hjoint, Np, static_cast<int>(Cl), ppt, static_cast<int>(col_chunk / EPIG_BN),
                                   static_cast<float>(1.0 / static_cast<double>(K)),
                                   static_cast<float>(1.0 / static_cast<double>(Nt))};
  return launch_gemm2<EPIG_BN, 6, 16, EpiEpigJoint<EPIG_BN>>(tmA, tmB, plan, ep, st, TAG_EPIG_JOINT);
}

size_t bvlm_epig_joint_workspace_bytes(int64_t Np, int64_t Nt, int64_t K, int64_t Cl) {
  const int64_t Kp = pad64(K);
  return static_cast<size_t>(round_up_i64(Np * Cl * Kp * 2, 256) + round_up_i64(Nt * Cl * Kp * 2, 256) + 1024);
}

int bvlm_epig_joint_entropy_f16(const void* pool16, int64_t Np, const void* targ16, int64_t Nt, int64_t K, int64_t Cl,
                                int64_t col_chunk, float* Hjoint, void* ws, size_t ws_bytes, void* stream) {
  if (pool16 == nullptr || targ16 == nullptr || Hjoint == nullptr || ws == nullptr) return BVLM_EINVAL;
  if (Np <= 0 || Nt <= 0 || K <= 0 || Cl <= 0) return BVLM_EINVAL;
  if (ws_bytes < bvlm_epig_joint_workspace_bytes(Np, Nt, K, Cl)) return BVLM_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return BVLM_EINVAL;
  const int64_t Kp = pad64(K);
  void* poolP = ws;
  void* targP = static_cast<uint8_t*>(ws) + round_up_i64(Np * Cl * Kp * 2, 256);
  int rc = bvlm_epig_prepare_from_probs(pool16, Np, K, Cl, poolP, nullptr, stream);
  if (rc) return rc;
  if ((rc = bvlm_epig_prepare_from_probs(targ16, Nt, K, Cl, targP, nullptr, stream))) return rc;
  return bvlm_epig_joint_entropy_operands(poolP, Np, targP, Nt, K, Cl, col_chunk, Hjoint, stream);
}

}  // extern "C"
