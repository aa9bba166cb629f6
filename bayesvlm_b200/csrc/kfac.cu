// K1 / K2 / K3: KFAC factor kernels.
//   K1  A += X^T X                       scripts/hessian_estimation.py:99-104
//   K2  InfoNCE GGN w.r.t. embeddings    bayesvlm/hessians.py:10-48
//   K3  SigLIP  GGN w.r.t. embeddings    bayesvlm/hessians.py:50-117
//
// K2/K3 never form the reference's [B,D,D] / [B,C,D] / [chunk,D,D] temporaries.  With L = Xh Yh^T (cosines),
// w_b = 1/|x_b|^2, J_b = (I - xh xh^T)/|x_b| and the per-source curvature S_b in embedding space,
//     H = s^2 sum_b w_b (I - xh xh^T) S_b (I - xh xh^T) = s^2 sum_b w_b [ S_b - xh u^T - u xh^T + a xh xh^T ],
//     u = S_b xh,  a = xh^T S_b xh.
// SigLIP:   S_b = Yh^T diag(lambda_b) Yh,  lambda = sigma(z)(1-sigma(z)).
// InfoNCE:  S_b = Yh^T (diag p_b - p_b p_b^T) Yh is the covariance of the targets under softmax p_b.  For a peaked
//   softmax the two terms cancel catastrophically (fp32 already loses digits, 16-bit operands lose everything), so
//   every row is centred on its pivot target g = yh[argmax_c L_bc]:  with rho = 1 - p*, e = sum_{c!=piv} p_c (yh_c - g)
//     S_b = sum_{c != piv} p_c yh_c yh_c^T  -  sym[ (e + (1-sqrt p*) g) (e + (1+sqrt p*) g)^T ]
//   in which every term is O(rho) (rho itself comes from the online softmax of pass 1 without forming 1 - p*).
// Pipeline per class batch (all GEMMs on the tcgen05 engine):
//   pass 1 (InfoNCE)  row panels of Xh Yh^T  -> row max, pivot, rest = sum_{c!=piv} exp(l - max)
//   pass 2            column panels of Xh Yh^T -> omega (fp16), omega*(d | L) (fp16), q_c = sum_b w_b omega_bc
//   pass 3            [n ; r] = [omega ; omega*d] Yh
//   finalise          per-source vectors e, u, a -> stacked operands
//   pass 4            Hinc = [Yh sqrt(q) | L_A | L_B]^T [Yh sqrt(q) | R_A | R_B]   (K = C + 2B, split-K)
//   H (+)= s^2 (Hinc + Hinc^T)/2
#include <cstdlib>

#include "epilogues.cuh"
#include "gemm2_engine.cuh"
#include "prep.cuh"

using namespace bvlm;

namespace {

constexpr int GGN_BN = 256;
constexpr int GGN_STAGES = 6;    // CTA-pair engine: 32 KB per stage
constexpr int GGN_W_STAGES = 4;  // the weights pass is store / epilogue bound: 4 operand stages + 64 KB of output slabs (2 per warp)
constexpr int GGN_ROWSTAT_SPLITS_MAX = 16;
constexpr int SYRK_BN = 128;
constexpr int SYRK_STAGES = 6;
constexpr float GGN_OPSCALE = 256.f;  // unit vectors -> fp16
constexpr float GGN_G = 16.f;         // scale of the stacked final-GEMM operands

inline int64_t pad64(int64_t k) { return round_up_i64(k, 64); }
inline int64_t pad128(int64_t k) { return round_up_i64(k, 128); }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return (off + 255) & ~static_cast<size_t>(255); }
};

struct GgnLayout {
  int64_t Dp, Cp, Bp, Bs, Ktot, Kst;
  size_t bytes;
  // Yh16 doubles as the unscaled (B) side of pass 4: [Ktot rows, Kst] = [yh (C rows) | R_A (B rows) | R_B (B rows)], of which
  // the tensor map of pass 4 reads the first Dp columns.  L16 [Ktot, Dp] is the scaled (A) side [q yh | L_A | L_B].
  __half *Xh16, *Yh16, *W16, *L16;
  float *inv_nx, *inv_ny, *w_raw, *w, *scalars, *rowmax2, *rest, *q, *MR, *Hinc;
  int* pivot;
  float4* rowinfo;
};

GgnLayout ggn_layout(int64_t B, int64_t C, int64_t D, int siglip, int prec, void* ws) {
  GgnLayout g{};
  g.Dp = pad64(D);
  g.Cp = pad64(C);
  g.Bp = pad64(B);
  g.Bs = pad128(B);
  g.Ktot = g.Cp + (siglip ? 1 : 2) * g.Bp;
  g.Kst = g.Dp * (prec == 3 ? 2 : 1);  // stored operand width: [hi | lo] for the split-precision logits
  Carver cv(ws);
  g.Xh16 = cv.take<__half>(static_cast<size_t>(B) * g.Kst);
  g.Yh16 = cv.take<__half>(static_cast<size_t>(g.Ktot) * g.Kst);
  g.W16 = cv.take<__half>(static_cast<size_t>(2 * g.Bs) * g.Cp);  // [omega ; omega*d] stacked (SigLIP uses the 2nd half)
  g.L16 = cv.take<__half>(static_cast<size_t>(g.Ktot) * g.Dp);
  g.inv_nx = cv.take<float>(static_cast<size_t>(B));
  g.inv_ny = cv.take<float>(static_cast<size_t>(C));
  g.w_raw = cv.take<float>(static_cast<size_t>(B));
  g.w = cv.take<float>(static_cast<size_t>(B));
  g.scalars = cv.take<float>(64);
  g.rowmax2 = cv.take<float>(static_cast<size_t>(B) * (GGN_ROWSTAT_SPLITS_MAX + 1));  // per-column-range partials + merged
  g.rest = cv.take<float>(static_cast<size_t>(B) * (GGN_ROWSTAT_SPLITS_MAX + 1));
  g.pivot = cv.take<int>(static_cast<size_t>(B) * (GGN_ROWSTAT_SPLITS_MAX + 1));
  g.rowinfo = cv.take<float4>(static_cast<size_t>(B));
  g.q = cv.take<float>(static_cast<size_t>(C));
  g.MR = cv.take<float>(static_cast<size_t>(2 * g.Bs) * D);
  g.Hinc = cv.take<float>(static_cast<size_t>(D) * D);
  g.bytes = cv.used() + 256;
  return g;
}

int ggn_impl(const float* X, int64_t B, int64_t ldx, const float* Y, int64_t C, int64_t ldy, int64_t D, float logit_scale,
             float logit_bias, int siglip, int prec, float* H, int64_t ldh, int accumulate, void* ws, size_t ws_bytes,
             cudaStream_t st) {
  if (X == nullptr || Y == nullptr || H == nullptr || ws == nullptr) return BVLM_EINVAL;
  if (prec != BVLM_PREC_X1 && prec != BVLM_PREC_X3) return BVLM_EINVAL;
  if (B <= 0 || C <= 0 || D <= 0 || ldx < D || ldy < D || ldh < D) return BVLM_EINVAL;
  if (B > (1 << 30) || C > (1 << 30)) return BVLM_EINVAL;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return BVLM_EINVAL;
  GgnLayout g = ggn_layout(B, C, D, siglip, prec, ws);
  if (ws_bytes < g.bytes) return BVLM_EWORKSPACE;
  int rc;
  const float s = expf(logit_scale);
  const float op2 = GGN_OPSCALE * GGN_OPSCALE;
  constexpr float kLog2e = 1.4426950408889634f;

  BVLM_CUDA_TRY(cudaMemsetAsync(g.scalars, 0, 64 * sizeof(float), st));
  BVLM_CUDA_TRY(cudaMemsetAsync(g.Hinc, 0, static_cast<size_t>(D) * D * sizeof(float), st));
  float* w_sum = g.scalars;

  // ---- operand preparation
  if ((rc = launch_ggn_row_prep(X, B, D, ldx, GGN_OPSCALE, prec, 0, g.Xh16, g.Dp, g.inv_nx, g.w_raw, w_sum, st))) return rc;
  if ((rc = launch_ggn_row_prep(Y, C, D, ldy, GGN_OPSCALE, prec, 1, g.Yh16, g.Dp, g.inv_ny, nullptr, nullptr, st))) return rc;
  if ((rc = launch_normalize_weights(g.w_raw, w_sum, B, g.w, st))) return rc;

  CUtensorMap tmX, tmY;
  // logit GEMM: X1 = one fp16 pass; X3 = hi.hi + lo.hi + hi.lo over operands stored as [hi | lo]
  const int64_t Kst = g.Kst;
  Operand16 opX{g.Xh16, B, Kst, FMT_F16};
  Operand16 opY{g.Yh16, C, Kst, FMT_F16};
  auto logit_plan = [&](int mode) {
    return prec == 3 ? make_split_plan2<GGN_BN>(static_cast<int>(B), static_cast<int>(C), static_cast<int>(g.Dp), mode, FMT_F16)
                     : make_plan2<GGN_BN>(static_cast<int>(B), static_cast<int>(C), static_cast<int>(g.Dp), mode, 1, FMT_F16);
  };
  if ((rc = operand_tmap<GEMM_BM>(&tmX, opX))) return rc;
  if ((rc = operand_tmap<GGN_BN / 2>(&tmY, opY))) return rc;  // CTA pairs: each CTA loads half of the B tile

  // ---- pass 1 (InfoNCE only): row max, pivot, rest.  Row panels are cut into column ranges so that every CTA pair
  //      gets an even share; the partial online-softmax states are merged by a tiny kernel.
  const int pairs = device_sm_count() / 2;
  float* rowmax2 = g.rowmax2;
  float* rest = g.rest;
  int* pivot = g.pivot;
  if (!siglip) {
    GemmPlan p1 = logit_plan(SCHED_ROW_PANEL);
    // whole rounds of the CTA pairs (B = 32768: 128 panels x 4 ranges = 512 items = 6.92 rounds, 98.8 % of the machine busy;
    // the first version's 5 ranges made 8.65 rounds = 96 %)
    const int S = balanced_panel_splits(p1.m_tiles, p1.n_tiles, pairs, GGN_ROWSTAT_SPLITS_MAX);
    p1.splits = S;
    EpiRowLse<GGN_BN>::Params e1{g.rowmax2, g.rest, g.pivot, s * kLog2e / op2, S};
    if ((rc = launch_gemm2<GGN_BN, GGN_STAGES, 8, EpiRowLse<GGN_BN>>(tmX, tmY, p1, e1, st, TAG_GGN_ROWSTATS))) return rc;
    if (S > 1) {
      if ((rc = launch_merge_rowstats(g.rowmax2, g.rest, g.pivot, B, S, st))) return rc;
      rowmax2 += static_cast<size_t>(B) * S;
      rest += static_cast<size_t>(B) * S;
      pivot += static_cast<size_t>(B) * S;
    }
  }

  if ((rc = launch_ggn_rowinfo(rowmax2, rest, pivot, g.w, B, siglip, g.rowinfo, st))) return rc;

  // ---- pass 2: curvature weights omega (fp16), omega*(d|L) (fp16), q
  __half* W16 = g.W16;
  __half* WL16 = g.W16 + static_cast<size_t>(g.Bs) * g.Cp;
  {
    GemmPlan p2 = logit_plan(SCHED_COL_PANEL);
    // cut every column panel into M ranges so that all SMs get an even share; q is then accumulated atomically
    const int ps = balanced_panel_splits(p2.n_tiles, p2.m_tiles, pairs, 16);
    p2.splits = ps;
    // both CTAs of a pair (and every M range) contribute to the same q columns: always accumulate atomically
    BVLM_CUDA_TRY(cudaMemsetAsync(g.q, 0, static_cast<size_t>(C) * sizeof(float), st));
    CUtensorMap tmW, tmWL;
    if ((rc = make_tmap_2d(&tmW, W16, TM_F16, static_cast<uint64_t>(g.Cp), static_cast<uint64_t>(B),
                           static_cast<uint64_t>(g.Cp) * 2, 64, 32, 1)))
      return rc;
    if ((rc = make_tmap_2d(&tmWL, WL16, TM_F16, static_cast<uint64_t>(g.Cp), static_cast<uint64_t>(B),
                           static_cast<uint64_t>(g.Cp) * 2, 64, 32, 1)))
      return rc;
    if (siglip) {
      EpiGgnWeights<GGN_BN, true>::Params e2{tmW, tmWL, g.rowinfo, g.q, s / op2, 1.0f / op2, logit_bias};
      if ((rc = launch_gemm2<GGN_BN, GGN_W_STAGES, 8, EpiGgnWeights<GGN_BN, true>>(tmX, tmY, p2, e2, st, TAG_GGN_WEIGHTS)))
        return rc;
    } else {
      EpiGgnWeights<GGN_BN, false>::Params e2{tmW, tmWL, g.rowinfo, g.q, s * kLog2e / op2, 1.0f / op2, 0.f};
      // (4-CTA clusters with the class tile multicast into two MMA pairs were measured for passes 1 and 2: slower, 0.82 vs
      //  0.77 ms and 1.54 vs 1.38 ms -- the bytes delivered per SM do not change)
      if ((rc = launch_gemm2<GGN_BN, GGN_W_STAGES, 8, EpiGgnWeights<GGN_BN, false>>(tmX, tmY, p2, e2, st, TAG_GGN_WEIGHTS)))
        return rc;
    }
  }

  if (g.Cp != C) {  // the K padding of pass 3 must be exact zeros (pass 2 leaves finite garbage there)
    if (!siglip)
      BVLM_CUDA_TRY(cudaMemset2DAsync(W16 + C, static_cast<size_t>(g.Cp) * 2, 0, static_cast<size_t>(g.Cp - C) * 2,
                                      static_cast<size_t>(B), st));
    BVLM_CUDA_TRY(cudaMemset2DAsync(WL16 + C, static_cast<size_t>(g.Cp) * 2, 0, static_cast<size_t>(g.Cp - C) * 2,
                                    static_cast<size_t>(B), st));
  }

  // ---- gamma = max_c q_c and the derived scalars
  if ((rc = launch_ggn_scalars(g.q, C, g.scalars, 1.0f / static_cast<float>(B), st))) return rc;
  const float* inv_gamma = g.scalars + 4;

  // ---- pass 3: InfoNCE [n ; r] = [omega ; omega*d] Yh (M = Bs + B stacked rows) | SigLIP r = (omega*L) Yh ; K = C.
  //      The B operand is Yh16 itself ([C, Dp] row-major = MN-major): no transposed copy of the targets.
  float* Nn = g.MR;
  float* Rr = siglip ? g.MR : g.MR + static_cast<size_t>(g.Bs) * D;
  {
    CUtensorMap tmW, tmYmn;
    Operand16 opW{siglip ? WL16 : W16, siglip ? B : g.Bs + B, g.Cp, FMT_F16};
    if ((rc = operand_tmap<GEMM_BM>(&tmW, opW))) return rc;
    if ((rc = operand_tmap_mn(&tmYmn, g.Yh16, C, g.Dp, Kst, FMT_F16))) return rc;  // hi segment; K rows beyond C read as zero
    GemmPlan p3 = make_plan2<GGN_BN>(static_cast<int>(opW.rows), static_cast<int>(D), static_cast<int>(g.Cp), SCHED_TILES, 1,
                                     FMT_F16);
    p3.idesc = make_idesc_f16(GEMM2_BM, GGN_BN, FMT_F16, FMT_F16, 0, 1);
    EpiStoreF32<GGN_BN>::Params e3{g.MR, D, 1.0f, 0, 0, nullptr, nullptr};
    if ((rc = launch_gemm2<GGN_BN, GGN_STAGES, 4, EpiStoreF32<GGN_BN>, false, true>(tmW, tmYmn, p3, e3, st, TAG_GGN_MOMENTS)))
      return rc;
  }

  // ---- stacked MN-major operands of pass 4 (row-major [K, Dp] fp16, written directly -- no transposes):
  //        scaled side   L16  = [ (q/gamma) yh * g^2/opscale | L_A | L_B ]
  //        unscaled side Yh16 = [ yh * opscale               | R_A | R_B ]      (first Dp columns of the Kst-wide rows)
  const float unscale_n = 1.0f / (GGN_WSCALE * GGN_OPSCALE);
  const float unscale_r =  // InfoNCE: d is in log2-logit units and stored with GGN_WDSCALE
      siglip ? unscale_n : 1.0f / (GGN_WDSCALE * GGN_OPSCALE * s * kLog2e);
  __half* R16 = g.Yh16;
  const int64_t rowA = g.Cp, rowB = siglip ? g.Cp : g.Cp + g.Bp;
  auto zero_rows = [&](__half* base, int64_t pitch, int64_t r0, int64_t r1) -> int {
    if (r1 > r0) BVLM_CUDA_TRY(cudaMemsetAsync(base + r0 * pitch, 0, static_cast<size_t>(r1 - r0) * pitch * 2, st));
    return BVLM_OK;
  };
  if ((rc = zero_rows(g.L16, g.Dp, C, g.Cp))) return rc;
  if ((rc = zero_rows(R16, Kst, C, g.Cp))) return rc;
  if (!siglip) {
    if ((rc = zero_rows(g.L16, g.Dp, rowA + B, rowA + g.Bp))) return rc;
    if ((rc = zero_rows(R16, Kst, rowA + B, rowA + g.Bp))) return rc;
  }
  if ((rc = zero_rows(g.L16, g.Dp, rowB + B, rowB + g.Bp))) return rc;
  if ((rc = zero_rows(R16, Kst, rowB + B, rowB + g.Bp))) return rc;
  if ((rc = launch_ggn_scale_targets(Y, C, D, g.Dp, ldy, g.inv_ny, g.q, inv_gamma, GGN_G * GGN_G / GGN_OPSCALE, g.L16, g.Dp, st)))
    return rc;
  if ((rc = launch_ggn_row_finalize(X, B, D, g.Dp, ldx, g.inv_nx, g.w, Y, ldy, g.inv_ny, pivot, rest, inv_gamma, Nn, Rr, D,
                                    unscale_n, unscale_r, siglip, GGN_G, g.L16 + rowA * g.Dp, R16 + rowA * Kst,
                                    g.L16 + rowB * g.Dp, R16 + rowB * Kst, g.Dp, Kst, st)))
    return rc;

  // ---- pass 4: Hinc = L16^T R16 (both MN-major), split along K, accumulated with red.global.add
  {
    const int64_t K = g.Ktot;
    CUtensorMap tmL, tmR;
    if ((rc = operand_tmap_mn(&tmL, g.L16, K, g.Dp, g.Dp, FMT_F16))) return rc;
    if ((rc = operand_tmap_mn(&tmR, R16, K, g.Dp, Kst, FMT_F16))) return rc;
    const int tiles = static_cast<int>(ceil_div_i64(D, GEMM2_BM) * ceil_div_i64(D, GGN_BN));
    int splits = pairs / tiles;
    if (splits < 1) splits = 1;
    GemmPlan p4 = make_plan2<GGN_BN>(static_cast<int>(D), static_cast<int>(D), static_cast<int>(K), SCHED_TILES, splits,
                                     FMT_F16);
    p4.idesc = make_idesc_f16(GEMM2_BM, GGN_BN, FMT_F16, FMT_F16, 1, 1);
    EpiStoreF32<GGN_BN>::Params e4{g.Hinc, D, 1.0f, 1, 0, nullptr, nullptr};
    if ((D * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(g.Hinc) & 15) == 0) {  // split-K partials as TMA bulk reductions
      if ((rc = make_tmap_2d(&e4.tm_c, g.Hinc, TM_F32, static_cast<uint64_t>(D), static_cast<uint64_t>(D),
                             static_cast<uint64_t>(D) * 4, 32, 32, 1)))
        return rc;
      e4.use_tma = 1;
    }
    if ((rc = launch_gemm2<GGN_BN, GGN_STAGES, 4, EpiStoreF32<GGN_BN>, true, true>(tmL, tmR, p4, e4, st, TAG_GGN_STACKED)))
      return rc;
  }
  // ---- H (+)= s^2 wbar gamma / g^2 * (Hinc + Hinc^T)/2
  return launch_sym_add(g.Hinc, D, D, H, ldh, s * s / (GGN_G * GGN_G), g.scalars + 3, accumulate, st);
}

}  // namespace

extern "C" {

size_t bvlm_syrk_workspace_bytes(int64_t n, int64_t d, int append_one, int precision) {
  (void)precision;
  const int64_t dA = d + (append_one ? 1 : 0);
  return static_cast<size_t>(round_up_i64(n * pad64(dA) * 2, 256)) + 3 * static_cast<size_t>(round_up_i64(dA * 4, 256)) + 512;
}

int bvlm_syrk_f32acc(const float* X, int64_t n, int64_t d, int64_t ldx, int append_one, int precision, float* C,
                     int64_t ldc, float alpha, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  if (X == nullptr || C == nullptr || ws == nullptr || n <= 0 || d <= 0 || ldx < d) return BVLM_EINVAL;
  if (precision != BVLM_PREC_X1) return BVLM_ENOTSUP;
  const int64_t dA = d + (append_one ? 1 : 0);
  if (ldc < dA) return BVLM_EINVAL;
  if (ws_bytes < bvlm_syrk_workspace_bytes(n, d, append_one, precision)) return BVLM_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return BVLM_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t dP = pad64(dA);
  Carver cv(ws);
  __half* X16 = cv.take<__half>(static_cast<size_t>(n) * dP);
  float* scale = cv.take<float>(static_cast<size_t>(dA));
  float* unscale = cv.take<float>(static_cast<size_t>(dA));
  unsigned int* amax = cv.take<unsigned int>(static_cast<size_t>(dA));
  int rc;
  if (!accumulate) {
    BVLM_CUDA_TRY(cudaMemset2DAsync(C, static_cast<size_t>(ldc) * 4, 0, static_cast<size_t>(dA) * 4, static_cast<size_t>(dA), st));
  }
  // per-feature power-of-two scaling keeps every feature inside fp16's normal range whatever its magnitude
  if ((rc = launch_col_pow2_scale(X, n, d, ldx, append_one, amax, scale, unscale, st))) return rc;
  // [X 1] as a row-major fp16 copy: it IS the MN-major operand of [X 1]^T [X 1] (rows = samples = K) -- no transpose
  if ((rc = launch_scale_cols_f16(X, n, d, ldx, scale, append_one, X16, dP, st))) return rc;
  CUtensorMap tm;
  if ((rc = operand_tmap_mn(&tm, X16, n, dA, dP, FMT_F16))) return rc;
  constexpr int BN = 256;  // square 256 x 256 tiles on CTA pairs; only the lower triangle of tiles is computed
  const int kp = static_cast<int>(pad64(n));
  GemmPlan plan = make_plan2<BN>(static_cast<int>(dA), static_cast<int>(dA), kp, SCHED_TRI_TILES, 1, FMT_F16);
  const int tri = plan.m_tiles * (plan.m_tiles + 1) / 2;
  int splits = (device_sm_count() / 2) / tri;
  if (splits < 1) splits = 1;
  // Bound the length of one tensor-core accumulation chain: the fp32 accumulator truncates, which biases a sum of n squares
  // by about -5.4e-9 per row of the chain (measured, scripts/syrk_bias.py: -5.5e-4 at n = 2^20 rows in 12 chains).  Chains of at
  // most 128 K blocks (8192 rows) keep the bias below 5e-5 -- the size of the fp16 operand rounding itself; every extra
  // partial tile costs one red.global.add pass (200 MB at n = 2^20, d = 768).
  constexpr int SYRK_MAX_CHAIN_KB = 128;
  const int min_splits = (plan.kb_total + SYRK_MAX_CHAIN_KB - 1) / SYRK_MAX_CHAIN_KB;
  if (splits < min_splits) {
    // more chains than CTA pairs: pick the count (from min_splits up) whose tri * splits items fill whole rounds of the pairs --
    // 131072 rows at d = 768: 16 chains make 96 items = 1.3 rounds of 74 pairs (65 % of the machine busy), 24 make 1.95 (97 %)
    const int pairs = device_sm_count() / 2;
    double best = 0.0;
    int best_s = min_splits;
    for (int sp = min_splits; sp <= 2 * min_splits; ++sp) {
      const long long items = static_cast<long long>(tri) * sp, rounds = (items + pairs - 1) / pairs;
      const double eff = static_cast<double>(items) / static_cast<double>(rounds * pairs);
      if (eff > best + 0.02) {
        best = eff;
        best_s = sp;
      }
    }
    splits = best_s;
  }
  if (splits > plan.kb_total) splits = plan.kb_total;
  plan.splits = splits;
  plan.idesc = make_idesc_f16(GEMM2_BM, BN, FMT_F16, FMT_F16, 1, 1);
  // lower triangle accumulated with red.global.add, then mirrored: C stays exactly symmetric
  EpiStoreF32<BN>::Params ep{C, ldc, alpha, 1, 1, nullptr, unscale};
  if ((ldc * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {  // split-K partials as TMA bulk reductions
    if ((rc = make_tmap_2d(&ep.tm_c, C, TM_F32, static_cast<uint64_t>(dA), static_cast<uint64_t>(dA),
                           static_cast<uint64_t>(ldc) * 4, 32, 32, 1)))
      return rc;
    ep.use_tma = 1;
  }
  if ((rc = launch_gemm2<BN, 6, 4, EpiStoreF32<BN>, true, true>(tm, tm, plan, ep, st, TAG_SYRK))) return rc;
  return launch_symmetrize_scale(C, dA, ldc, 1.0f, st);
}

size_t bvlm_ggn_workspace_bytes(int64_t B, int64_t C, int64_t D, int precision) {
  if (B <= 0 || C <= 0 || D <= 0 || (precision != BVLM_PREC_X1 && precision != BVLM_PREC_X3)) return 0;
  return ggn_layout(B, C, D, 0, precision, nullptr).bytes;
}

int bvlm_ggn_infonce(const float* X, int64_t B, int64_t ldx, const float* Y, int64_t C, int64_t ldy, int64_t D,
                     float logit_scale, int precision, float* H, int64_t ldh, int accumulate, void* ws, size_t ws_bytes,
                     void* stream) {
  return ggn_impl(X, B, ldx, Y, C, ldy, D, logit_scale, 0.f, 0, precision, H, ldh, accumulate, ws, ws_bytes,
                  static_cast<cudaStream_t>(stream));
}

int bvlm_ggn_siglip(const float* X, int64_t B, int64_t ldx, const float* Y, int64_t C, int64_t ldy, int64_t D,
                    float logit_scale, float logit_bias, int precision, float* H, int64_t ldh, int accumulate, void* ws,
                    size_t ws_bytes, void* stream) {
  return ggn_impl(X, B, ldx, Y, C, ldy, D, logit_scale, logit_bias, 1, precision, H, ldh, accumulate, ws, ws_bytes,
                  static_cast<cudaStream_t>(stream));
}

}  // extern "C"
