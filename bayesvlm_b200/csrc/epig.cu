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
constexpr int EPIG_STAGES = 4;

inline int64_t pad64(int64_t k) { return round_up_i64(k, 64); }

__device__ __forceinline__ float round_f16(float v) { return __half2float(__float2half_rn(v)); }

// x * log(x) with torch's Half semantics (both the log and the product are rounded to fp16); 0 at x == 0.
__device__ __forceinline__ float xlogx_f16(float x) {
  if (x == 0.f) return 0.f;
  const float l = round_f16(logf(x));
  return round_f16(x * l);
}

// ---------------------------------------------------------------------------------------------------------------
// E0: probs[n,k,:] = softmax(mean[n,:] + eps[k,n,:] * sqrt(var[n,:]))  -> fp16           (vlm.py:116-123)
// one thread per (n,k); consecutive threads walk k so the [N,K,Cl] writes are contiguous.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_epig_sample(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ eps, int64_t N,
              int64_t K, int64_t Cl, __half* __restrict__ probs) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  const int64_t n = idx / K;
  const int64_t k = idx - n * K;
  const float* m = mean + n * Cl;
  const float* v = var + n * Cl;
  const float* e = eps + (k * N + n) * Cl;
  float zmax = -INFINITY, zsum = 0.f;
  for (int64_t c = 0; c < Cl; ++c) {
    const float z = fmaf(e[c], sqrtf(v[c]), m[c]);
    if (z > zmax) {
      zsum = zsum * expf(zmax - z);
      zmax = z;
    }
    zsum += expf(z - zmax);
  }
  __half* o = probs + idx * Cl;
  for (int64_t c = 0; c < Cl; ++c) {
    const float z = fmaf(e[c], sqrtf(v[c]), m[c]);
    o[c] = __float2half_rn(expf(z - zmax) / zsum);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// E1: marginal entropy of fp16 probabilities, one warp per sample                      (epig.py:294-311, 275-292)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_epig_marginal(const __half* __restrict__ probs, int64_t N, int64_t K, int64_t Cl, __half* __restrict__ out) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const __half* p = probs + n * K * Cl;
  float ent = 0.f;
  for (int64_t c = lane; c < Cl; c += 32) {
    float s = 0.f;
    for (int64_t k = 0; k < K; ++k) s += __half2float(p[k * Cl + c]);
    const float pbar = round_f16(s / static_cast<float>(K));  // torch.mean on Half: fp32 sum, divide, round
    ent += xlogx_f16(pbar);
  }
  ent = warp_sum(ent);
  if (lane == 0) out[n] = __float2half_rn(-round_f16(ent));
}

// ---------------------------------------------------------------------------------------------------------------
// [N, K, Cl] fp16 -> [N, Cl, Kp] fp16 (K-major GEMM operand, zero padded along K)     (epig.py:374-376 permutes)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_epig_permute(const __half* __restrict__ in, int64_t N, int64_t K, int64_t Cl, int64_t Kp, __half* __restrict__ out) {
  extern __shared__ __half s_tile[];  // [K][Cl]
  const int64_t n = blockIdx.x;
  if (n >= N) return;
  const __half* src = in + n * K * Cl;
  for (int64_t i = threadIdx.x; i < K * Cl; i += blockDim.x) s_tile[i] = src[i];
  __syncthreads();
  __half* dst = out + n * Cl * Kp;
  for (int64_t i = threadIdx.x; i < Cl * Kp; i += blockDim.x) {
    const int64_t c = i / Kp;
    const int64_t k = i - c * Kp;
    dst[i] = k < K ? s_tile[k * Cl + c] : __float2half_rn(0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// E2 epilogue. Tile rows are (pool row, class) pairs packed ppt = floor(128/Cl) pool rows per 128-row CTA slab; tile
// columns are the flattened (target, class) axis. Row panels: each CTA pair walks all column tiles of its pool rows.
// Per element (two at a time in half2 registers where the arithmetic is fp16 anyway):
//   h = fp16(acc); j = fp16(h / K); l = fp16(log j); t = fp16(j * l)  [= HMUL2, exact product rounded once]; sum += t
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
    float inv_K;        // 1 / number of MC samples
    float Nt;           // number of target points (divisor)
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
      float sum = 0.f;
      for (int h = 0; h < ctx.n_warps / 4; ++h)
        for (int c = 0; c < p.Cl; ++c) sum += lds_f32(s + 4u * static_cast<uint32_t>(h * 128 + r + c));
      const float neg = -round_f16(sum);               // fp16(sum) then negate
      st.hj += round_f16(neg / p.Nt);                  // "/ N_t" on a Half tensor
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
    // columns beyond N are TMA zero fill: joint = 0 -> 0 * max(log 0, -65504) = -0, no masking needed
    const __half2 lo_clamp = __floats2half2_rn(-65504.f, -65504.f);
    const float2 inv_k2 = make_float2(p.inv_K, p.inv_K), ln2 = make_float2(0.6931471805599453f, 0.6931471805599453f);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      // packed fp32x2 multiplies (FMUL2) and mixed-precision adds (fp32 += fp16, FHADD): same roundings, fewer issue slots
      const float2 h = __half22float2(__floats2half2_rn(v[j], v[j + 1]));            // matmul output rounded to fp16
      const __half2 jt = __float22half2_rn(__fmul2_rn(h, inv_k2));                    // "/ K" on a Half tensor
      const float2 jf = __half22float2(jt);
      __half2 lg = __float22half2_rn(__fmul2_rn(make_float2(fast_log2(jf.x), fast_log2(jf.y)), ln2));
      lg = __hmax2(lg, lo_clamp);                                                     // log 0 = -inf would make 0 * inf
      const __half2 t = __hmul2(jt, lg);                                              // fp16(j * fp16(log j))
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

int bvlm_epig_sample_probs(const float* mean, const float* var, const float* eps, int64_t N, int64_t K, int64_t Cl,
                           void* probs16, void* stream) {
  if (mean == nullptr || var == nullptr || eps == nullptr || probs16 == nullptr) return BVLM_EINVAL;
  if (N <= 0 || K <= 0 || Cl <= 0) return BVLM_OK;
  const int64_t total = N * K;
  k_epig_sample<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mean, var, eps, N, K, Cl, static_cast<__half*>(probs16));
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int bvlm_epig_marginal_entropy_f16(const void* probs16, int64_t N, int64_t K, int64_t Cl, void* out16, void* stream) {
  if (probs16 == nullptr || out16 == nullptr) return BVLM_EINVAL;
  if (N <= 0) return BVLM_OK;
  k_epig_marginal<<<static_cast<unsigned>((N + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(probs16), N, K, Cl, static_cast<__half*>(out16));
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

size_t bvlm_epig_joint_workspace_bytes(int64_t Np, int64_t Nt, int64_t K, int64_t Cl) {
  const int64_t Kp = pad64(K);
  return static_cast<size_t>(round_up_i64(Np * Cl * Kp * 2, 256) + round_up_i64(Nt * Cl * Kp * 2, 256) + 1024);
}

int bvlm_epig_joint_entropy_f16(const void* pool16, int64_t Np, const void* targ16, int64_t Nt, int64_t K, int64_t Cl,
                                int64_t col_chunk, float* Hjoint, void* ws, size_t ws_bytes, void* stream) {
  if (pool16 == nullptr || targ16 == nullptr || Hjoint == nullptr || ws == nullptr) return BVLM_EINVAL;
  if (Np <= 0 || Nt <= 0 || K <= 0 || Cl <= 0) return BVLM_EINVAL;
  if (Cl > 128 || col_chunk <= 0 || (col_chunk % EPIG_BN) != 0) return BVLM_ENOTSUP;
  if (K * Cl * 2 > 48 * 1024) return BVLM_ENOTSUP;
  if (Nt * Cl > 0x7fffffff || Np > 0x7fffffff) return BVLM_EINVAL;
  if (ws_bytes < bvlm_epig_joint_workspace_bytes(Np, Nt, K, Cl)) return BVLM_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return BVLM_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t Kp = pad64(K);
  __half* poolP = static_cast<__half*>(ws);
  __half* targP = reinterpret_cast<__half*>(static_cast<uint8_t*>(ws) + round_up_i64(Np * Cl * Kp * 2, 256));
  const size_t sh = static_cast<size_t>(K * Cl * 2);
  k_epig_permute<<<static_cast<unsigned>(Np), 256, sh, st>>>(static_cast<const __half*>(pool16), Np, K, Cl, Kp, poolP);
  count_launch();
  k_epig_permute<<<static_cast<unsigned>(Nt), 256, sh, st>>>(static_cast<const __half*>(targ16), Nt, K, Cl, Kp, targP);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());

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
                                   1.0f / static_cast<float>(K), static_cast<float>(Nt)};
  return launch_gemm2<EPIG_BN, 6, 16, EpiEpigJoint<EPIG_BN>>(tmA, tmB, plan, ep, st, TAG_EPIG_JOINT);
}

}  // extern "C"
