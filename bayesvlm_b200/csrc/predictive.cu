// P1 / P2 / P3: quadratic forms, fused Kronecker-Laplace predictive GEMM, probit softmax.
// Reference: bayesvlm/vlm.py:630-684 (CLIP._compute_probabilistic_logits_smith), scripts/zeroshot.py:119-120.
#include "../../include/bvlm.h"
#include <cstdlib>

#include "epilogues.cuh"
#include "gemm2_engine.cuh"
#include "prep.cuh"

using namespace bvlm;

namespace {

constexpr int PRED_BN = 256;
constexpr int PRED_STAGES = 4;
constexpr float PRED_OPSCALE = 256.f;  // unit-energy embeddings are scaled into the fp16 sweet spot

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

inline int64_t pad64(int64_t k) { return round_up_i64(k, 64); }
inline int64_t pad128(int64_t k) { return round_up_i64(k, 128); }
constexpr float PRED_OPSCALE_F8 = 128.f * 32.f;  // target side of the fp16 + fp8 mode: unit-energy * 128 (fp8) * 2^5 (fp16)

// out[i] = | W16 act16_i |^2  through the row-panel GEMM with the sum-of-squares epilogue.
int quadform_impl(const float* act, int64_t n, int64_t d, int64_t ld, int append_one, const void* W16, int64_t dA,
                  int64_t k_pad, float w_scale, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                  bool converted = false, const EmbedPrepArgs* side = nullptr, bool pdl = false) {
  if (n <= 0) return BVLM_OK;
  if (dA != d + (append_one ? 1 : 0) || k_pad != pad64(dA)) return BVLM_EINVAL;
  if (ws_bytes < bvlm_quadform_workspace_bytes(n, d, append_one)) return BVLM_EWORKSPACE;
  Carver cv(ws);
  __half* act16 = cv.take<__half>(static_cast<size_t>(n) * k_pad);
  float* row_unscale = cv.take<float>(static_cast<size_t>(n));
  int rc = BVLM_OK;
  if (!converted) {  // (the predictive converts the activations in its fused row-prep kernel)
    rc = launch_rows_to_16(act, n, d, ld, append_one, FMT_F16, 1, 1.0f, act16, k_pad, row_unscale, st);
    if (rc) return rc;
  }
  CUtensorMap tmA, tmB;
  Operand16 opA{act16, n, k_pad, FMT_F16};
  Operand16 opB{W16, dA, k_pad, FMT_F16};
  if ((rc = operand_tmap<GEMM_BM>(&tmA, opA))) return rc;
  if ((rc = operand_tmap<PRED_BN / 2>(&tmB, opB))) return rc;  // CTA pairs: each CTA loads half of the B tile
  // row panels on CTA pairs; W is lower triangular, so the K loop of column tile n stops at its diagonal block
  GemmPlan plan = make_plan2<PRED_BN>(static_cast<int>(n), static_cast<int>(dA), static_cast<int>(k_pad), SCHED_ROW_PANEL, 1,
                                      FMT_F16);
  plan.tri_k = 1;
  EpiRowSumSq<PRED_BN>::Params ep{out, row_unscale, 1.0f / (w_scale * w_scale)};
  if (side != nullptr) {  // the epilogue warps also convert the embedding rows of their panel (EpiQuadformPrep)
    const int row_bytes = static_cast<int>(side->D * 4);
    int slot_shift = 0;
    while ((2 << slot_shift) <= std::min(PREP_MAX_SLOTS, PREP_RING_BYTES / row_bytes)) ++slot_shift;
    const int slots = 1 << slot_shift;
    const bool exact = side->D % 128 == 0 && side->seg_pad == side->D && (side->nsplit != 2 || side->seg8 == side->D);
#define BVLM_QPREP(EVX)                                                                                       \
    do {                                                                                                      \
      EpiQuadformPrep<PRED_BN, EVX>::Params ep2{ep, *side, row_bytes, slots, slot_shift};                     \
      return launch_gemm2<PRED_BN, 4, 8, EpiQuadformPrep<PRED_BN, EVX>>(tmA, tmB, plan, ep2, st, TAG_QUADFORM, nullptr, nullptr, pdl); \
    } while (0)
    if (exact && side->D == 512) BVLM_QPREP(4);
    if (exact && side->D == 768) BVLM_QPREP(6);
    if (exact && side->D == 1024) BVLM_QPREP(8);
    BVLM_QPREP(0);
#undef BVLM_QPREP
  }
  return launch_gemm2<PRED_BN, 6, 4, EpiRowSumSq<PRED_BN>>(tmA, tmB, plan, ep, st, TAG_QUADFORM, nullptr, nullptr, pdl);
}

}  // namespace

extern "C" {

int64_t bvlm_padded_k(int64_t k) { return pad64(k); }

int64_t bvlm_padded_cols(int64_t c) { return round_up_i64(c, PRED_BN); }

int bvlm_factor_prepare(const float* W, int64_t dA, int64_t ldw, float w_scale, void* W16, int64_t k_pad, void* stream) {
  if (W == nullptr || W16 == nullptr || dA <= 0 || k_pad != pad64(dA)) return BVLM_EINVAL;
  return launch_rows_to_16(W, dA, dA, ldw, 0, FMT_F16, 0, w_scale, W16, k_pad, nullptr, static_cast<cudaStream_t>(stream));
}

size_t bvlm_quadform_workspace_bytes(int64_t n, int64_t d, int append_one) {
  const int64_t k_pad = pad64(d + (append_one ? 1 : 0));
  size_t b = 0;
  b += round_up_i64(static_cast<int64_t>(n) * k_pad * 2, 256) + 256;
  b += round_up_i64(static_cast<int64_t>(n) * 4, 256) + 256;
  return b;
}

int bvlm_quadform(const float* act, int64_t n, int64_t d, int64_t ld, int append_one, const void* W16, int64_t dA,
                  int64_t k_pad, float w_scale, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (act == nullptr || W16 == nullptr || out == nullptr || ws == nullptr) return BVLM_EINVAL;
  return quadform_impl(act, n, d, ld, append_one, W16, dA, k_pad, w_scale, out, ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

size_t bvlm_predictive_target_workspace_bytes(int64_t C, int64_t D, int64_t d_act, int append_one) {
  (void)D;
  return bvlm_quadform_workspace_bytes(C, d_act, append_one) + round_up_i64(C * 4, 256) + 512;
}

int64_t bvlm_predictive_t8_cols(int64_t D) { return 2 * pad128(D); }

int bvlm_predictive_target_prepare(const float* T, int64_t C, int64_t D, int64_t ldt, const float* Tact, int64_t d_act,
                                   int64_t ldact, int append_one, const void* Wt16, int64_t dA, int64_t k_pad,
                                   float w_scale, const float* beta, float sum_delta, float kappa, int precision,
                                   void* T16, void* T8, float* colA, float* colB, void* ws, size_t ws_bytes, void* stream) {
  if (T == nullptr || Tact == nullptr || Wt16 == nullptr || beta == nullptr || T16 == nullptr || colA == nullptr ||
      colB == nullptr || ws == nullptr)
    return BVLM_EINVAL;
  if (precision != BVLM_PREC_X1 && precision != BVLM_PREC_X3 && precision != BVLM_PREC_X2F8) return BVLM_EINVAL;
  if (precision == BVLM_PREC_X2F8 && T8 == nullptr) return BVLM_EINVAL;
  if (ws_bytes < bvlm_predictive_target_workspace_bytes(C, D, d_act, append_one)) return BVLM_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver cv(ws);
  float* gamma = cv.take<float>(static_cast<size_t>(C));
  const size_t used = cv.used();
  int rc = quadform_impl(Tact, C, d_act, ldact, append_one, Wt16, dA, k_pad, w_scale, gamma,
                         static_cast<uint8_t*>(ws) + used, ws_bytes - used, st);
  if (rc) return rc;
  // the epilogue reads colA / colB in whole column tiles: zero the padded tail
  const int64_t cpad = bvlm_padded_cols(C);
  if (cpad > C) {
    BVLM_CUDA_TRY(cudaMemsetAsync(colA + C, 0, static_cast<size_t>(cpad - C) * sizeof(float), st));
    BVLM_CUDA_TRY(cudaMemsetAsync(colB + C, 0, static_cast<size_t>(cpad - C) * sizeof(float), st));
  }
  // side 1: out0 = gamma/E, out1 = (gamma*kappa + sum_d beta_d t_d^2)/E ; E = |t|^2 + gamma * sum(delta)
  return launch_predictive_row_prep(T, C, D, ldt, gamma, beta, sum_delta, kappa, 0.f, /*side=*/1, precision,
                                    precision == BVLM_PREC_X2F8 ? PRED_OPSCALE_F8 : PRED_OPSCALE, static_cast<__half*>(T16),
                                    pad64(D), 0, static_cast<uint8_t*>(T8), pad128(D), colA, colB, st);
}

size_t bvlm_predictive_workspace_bytes(int64_t N, int64_t D, int64_t d_act, int append_one, int precision) {
  size_t b = bvlm_quadform_workspace_bytes(N, d_act, append_one);
  b += 7 * (round_up_i64(N * 4, 256) + 256);
  b += round_up_i64(N * operand_pitch(pad64(D) * (precision == 3 ? 2 : 1)) * 2, 256) + 256;
  if (precision == BVLM_PREC_X2F8) b += round_up_i64(N * 2 * pad128(D), 256) + 256;
  return b;
}

int bvlm_predictive(const float* E, int64_t N, int64_t D, int64_t lde, const float* Eact, int64_t d_act, int64_t ldact,
                    int append_one, const void* Wi16, int64_t dA, int64_t k_pad, float w_scale, const float* delta,
                    float sum_beta, float logit_scale, const float* logit_scale_dev, const void* T16, const void* T8,
                    const float* colA, const float* colB, int64_t C, int precision, float* mean, float* var, float* probs,
                    int64_t ldo, void* ws, size_t ws_bytes, void* stream) {
  if (E == nullptr || Eact == nullptr || Wi16 == nullptr || delta == nullptr || T16 == nullptr || colA == nullptr ||
      colB == nullptr || mean == nullptr || var == nullptr || ws == nullptr)
    return BVLM_EINVAL;
  if (precision != BVLM_PREC_X1 && precision != BVLM_PREC_X3 && precision != BVLM_PREC_X2F8) return BVLM_EINVAL;
  if (precision == BVLM_PREC_X2F8 && T8 == nullptr) return BVLM_EINVAL;
  if (N <= 0 || C <= 0) return BVLM_OK;
  if (ldo < C || N > 0x7fffffff || C > 0x7fffffff) return BVLM_EINVAL;
  if (ws_bytes < bvlm_predictive_workspace_bytes(N, D, d_act, append_one, precision)) return BVLM_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t seg = pad64(D);
  const int64_t kp = seg * (precision == 3 ? 2 : 1);  // stored operand width: [hi | lo] for the split mode
  Carver cv(ws);
  float* alpha = cv.take<float>(static_cast<size_t>(N));
  float* n2 = cv.take<float>(static_cast<size_t>(N));
  float* pd = cv.take<float>(static_cast<size_t>(N));
  float* esc = cv.take<float>(static_cast<size_t>(N));
  const int64_t e_pitch = operand_pitch(kp);
  __half* E16 = cv.take<__half>(static_cast<size_t>(N) * e_pitch);
  const int64_t seg8 = pad128(D);
  uint8_t* A8 = precision == BVLM_PREC_X2F8 ? cv.take<uint8_t>(static_cast<size_t>(N) * 2 * seg8) : nullptr;
  const size_t used = cv.used();
  // One HBM-bound pass over the image rows converts BOTH operands (activations -> fp16 for the quadratic forms,
  // embeddings -> fp16 [+ fp16 lo | + fp8 compensation terms]); neither depends on the quadratic forms, whose 1/sqrt(E_i)
  // normalisation is applied by the epilogue of the mean GEMM.
  int rc;
  if (dA != d_act + (append_one ? 1 : 0) || k_pad != pad64(dA)) return BVLM_EINVAL;
  uint8_t* qws = static_cast<uint8_t*>(ws) + used;
  Carver qcv(qws);  // same carve as quadform_impl
  __half* act16 = qcv.take<__half>(static_cast<size_t>(N) * k_pad);
  float* act_unscale = qcv.take<float>(static_cast<size_t>(N));
  // Embeddings: converted by the epilogue warps of the quadratic-form GEMM (tensor-bound, its epilogue nearly idle) when the
  // rows are 16-byte aligned and fit its register-resident row routine; by the row-prep kernel otherwise.
  static const bool fuse_env = [] {
    const char* e = getenv("BVLM_PRED_FUSE_PREP");
    return e == nullptr || atoi(e) != 0;
  }();
  const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool fuse = fuse_env && al16(E) && al16(delta) && (lde % 4) == 0 && (D % 4) == 0 && D * 4 <= PREP_RING_BYTES &&
                    (e_pitch % 4) == 0 && seg <= 1024 && (precision != BVLM_PREC_X2F8 || seg8 <= 1024);
  EmbedPrepArgs side{E, N, D, lde, delta, precision, E16, seg, e_pitch, A8, seg8, n2, pd, esc};
  // The three kernels of the step are chained by programmatic dependent launches: the CTAs of the next kernel become resident and
  // run their prologue (barrier init, tensor-memory allocation, tensor-map prefetch, cluster sync) on SMs the previous kernel
  // has already left, and wait with griddepcontrol.wait before they touch its results (0.2948 -> 0.2913 ms per step).
  static const bool pdl = [] {
    const char* e = getenv("BVLM_PRED_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  rc = launch_predictive_embed_prep(fuse ? nullptr : E, N, D, lde, delta, precision, E16, seg, e_pitch, A8, seg8, n2, pd, esc,
                                    Eact, d_act, ldact, append_one, act16, k_pad, act_unscale, st);
  if (rc) return rc;
  rc = quadform_impl(Eact, N, d_act, ldact, append_one, Wi16, dA, k_pad, w_scale, alpha, qws, ws_bytes - used, st,
                     /*converted=*/true, fuse ? &side : nullptr, pdl);
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  Operand16 opA{E16, N, kp, FMT_F16, e_pitch};
  Operand16 opB{T16, C, kp, FMT_F16};
  if ((rc = operand_tmap<GEMM_BM>(&tmA, opA))) return rc;
  static const int variant = [] {
    // 1: 4 epilogue warps + 5 stages (default), 2: 8 epilogue warps, 3: as 1 on 4-CTA clusters with the class tile multicast
    const char* e = getenv("BVLM_PRED_VARIANT");
    return e != nullptr ? atoi(e) : 1;
  }();
  GemmPlan plan = precision == 3
                      ? make_split_plan2<PRED_BN>(static_cast<int>(N), static_cast<int>(C), static_cast<int>(seg), SCHED_TILES, FMT_F16)
                      : make_plan2<PRED_BN>(static_cast<int>(N), static_cast<int>(C), static_cast<int>(kp), SCHED_TILES, 1, FMT_F16);
  if ((rc = operand_tmap<PRED_BN / 2>(&tmB, opB))) return rc;  // each CTA of a pair loads half of the B tile
  CUtensorMap tmA8, tmB8;
  if (precision == BVLM_PREC_X2F8) {
    if ((rc = make_tmap_2d(&tmA8, A8, TM_U8, static_cast<uint64_t>(2 * seg8), static_cast<uint64_t>(N),
                           static_cast<uint64_t>(2 * seg8), 128, GEMM_BM, 1)))
      return rc;
    if ((rc = make_tmap_2d(&tmB8, T8, TM_U8, static_cast<uint64_t>(2 * seg8), static_cast<uint64_t>(C),
                           static_cast<uint64_t>(2 * seg8), 128, PRED_BN / 2, 1)))
      return rc;
    plan.kb_alt = plan.kb_total;                           // fp16 hi.hi blocks first ...
    plan.kb_total += static_cast<int>(2 * seg8 / 128);     // ... then [lo8 | e8] x [t8 | tlo8] in E4M3
    plan.idesc_alt = make_idesc_e4m3(GEMM2_BM, PRED_BN);
  }
  const CUtensorMap* pA8 = precision == BVLM_PREC_X2F8 ? &tmA8 : nullptr;
  const CUtensorMap* pB8 = precision == BVLM_PREC_X2F8 ? &tmB8 : nullptr;
  EpiPredictive<PRED_BN>::Params ep{};
  ep.mean = mean;
  ep.var = var;
  ep.ld = ldo;
  ep.alpha = alpha;
  ep.n2 = n2;
  ep.pd = pd;
  ep.esc = esc;
  ep.sum_beta = sum_beta;
  ep.ls = logit_scale;
  ep.ls_dev = logit_scale_dev;
  ep.mean_unscale = precision == BVLM_PREC_X2F8 ? 1.0f / (128.f * 1024.f) : 1.0f / PRED_OPSCALE;
  ep.a = colA;
  ep.b = colB;
  // TMA stores need 16-byte aligned rows; tiny class counts (e.g. C = 10) fall back to direct stores
  ep.use_tma = ((ldo * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(mean) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(var) & 15) == 0) ? 1 : 0;
  if (ep.use_tma) {
    int box_rows = 32;
#ifdef BVLM_DIAG
    if (const char* de = getenv("BVLM_DEBUG_EPI"); de != nullptr && atoi(de) == 5) box_rows = 16;
#endif
    if ((rc = make_tmap_2d(&ep.tm_mean, mean, TM_F32, static_cast<uint64_t>(C), static_cast<uint64_t>(N),
                           static_cast<uint64_t>(ldo) * 4, 32, box_rows, 1)))
      return rc;
    if ((rc = make_tmap_2d(&ep.tm_var, var, TM_F32, static_cast<uint64_t>(C), static_cast<uint64_t>(N),
                           static_cast<uint64_t>(ldo) * 4, 32, box_rows, 1)))
      return rc;
  }
#ifdef BVLM_DIAG  // diagnostic builds only (python -m bayesvlm_b200.build --diag): never in the shipped library
  if (getenv("BVLM_DEBUG_NOSTORE") != nullptr) ep.use_tma = 2;  // main loop without output traffic
  if (const char* de = getenv("BVLM_DEBUG_EPI"); de != nullptr && ep.use_tma)  // 1 / 2 / 3, see EpiPredictive::chunk; 4: direct stores
    ep.use_tma = atoi(de) == 4 ? 0 : 1 + atoi(de);
  if (getenv("BVLM_DEBUG_SHORTK") != nullptr) {                 // one K block per tile -> the kernel is its epilogue
    plan.kb_total = 1;
    plan.kb_alt = 0x7fffffff;
    plan.seg_kb = 0;
  }
#endif
  if (variant == 3) {
    plan_use_pairs(plan, 2);
    rc = launch_gemm2<PRED_BN, 5, 4, EpiPredictive<PRED_BN>, false, false, 2>(tmA, tmB, plan, ep, st, TAG_PREDICTIVE, pA8, pB8);
  } else if (variant == 1)
    rc = launch_gemm2<PRED_BN, 5, 4, EpiPredictive<PRED_BN>>(tmA, tmB, plan, ep, st, TAG_PREDICTIVE, pA8, pB8, pdl);
  else
    rc = launch_gemm2<PRED_BN, 5, 8, EpiPredictive<PRED_BN>>(tmA, tmB, plan, ep, st, TAG_PREDICTIVE, pA8, pB8);
  if (rc) return rc;
  if (probs != nullptr) rc = launch_probit_softmax(mean, var, N, C, ldo, probs, st);
  return rc;
}

int bvlm_probit_softmax(const float* mean, const float* var, int64_t N, int64_t C, int64_t ld, float* probs,
                        void* stream) {
  if (mean == nullptr || var == nullptr || probs == nullptr || ld < C) return BVLM_EINVAL;
  return launch_probit_softmax(mean, var, N, C, ld, probs, static_cast<cudaStream_t>(stream));
}

int bvlm_mc_softmax_accumulate(const float* mean, const float* var, const float* eps, int64_t N, int64_t C, int64_t G,
                               float* acc_probs, float* acc_entropy, void* stream) {
  if (mean == nullptr || var == nullptr || eps == nullptr || (acc_probs == nullptr && acc_entropy == nullptr)) return BVLM_EINVAL;
  if (G > 0x7fffffff) return BVLM_EINVAL;
  return launch_mc_softmax(mean, var, eps, N, C, static_cast<int>(G), acc_probs, acc_entropy, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
