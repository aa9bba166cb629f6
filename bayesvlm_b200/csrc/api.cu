// Library-level entry points: version / status strings, device check, launch counter and the diagnostic
// plain-GEMM path the tests use to validate the tcgen05 engine in isolation.
#include <atomic>
#include <mutex>
#include <vector>

#include <cstdlib>

#include "epilogues.cuh"
#include "gemm2_engine.cuh"
#include "prep.cuh"

namespace bvlm {
namespace {
std::atomic<int64_t> g_launches{0};
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {
struct TimedLaunch {
  int tag;
  cudaEvent_t beg, end;
};
std::atomic<int> g_timing_on{0};
std::mutex g_timing_mu;
std::vector<TimedLaunch> g_timed;
cudaEvent_t g_pending_beg = nullptr;
}  // namespace

void timing_begin(int tag, cudaStream_t st) {
  (void)tag;
  if (!g_timing_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (cudaEventCreate(&g_pending_beg) != cudaSuccess) {
    g_pending_beg = nullptr;
    return;
  }
  cudaEventRecord(g_pending_beg, st);
}
void timing_end(int tag, cudaStream_t st) {
  if (!g_timing_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (g_pending_beg == nullptr) return;
  TimedLaunch t{tag, g_pending_beg, nullptr};
  g_pending_beg = nullptr;
  if (cudaEventCreate(&t.end) != cudaSuccess) {
    cudaEventDestroy(t.beg);
    return;
  }
  cudaEventRecord(t.end, st);
  g_timed.push_back(t);
}
}  // namespace bvlm

using namespace bvlm;

extern "C" {

const char* bvlm_version(void) { return "bvlm 0.1.0 (sm_100a; tcgen05+TMA)"; }

const char* bvlm_status_string(int status) {
  switch (status) {
    case BVLM_OK: return "ok";
    case BVLM_EINVAL: return "invalid argument";
    case BVLM_ENOTSUP: return "unsupported configuration";
    case BVLM_EDRIVER: return "CUDA driver entry point unavailable (cuTensorMapEncodeTiled)";
    case BVLM_EWORKSPACE: return "workspace too small";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown status";
}

int bvlm_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (major != 10) return BVLM_ENOTSUP;
  CUtensorMap probe;
  // encoding a tensor map only inspects the address (alignment): no device memory is allocated by the library
  return make_tmap_2d(&probe, reinterpret_cast<const void*>(static_cast<uintptr_t>(1) << 21), TM_F16, 64, 128, 128, 64, 128, 1);
}

int64_t bvlm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int bvlm_timing_enable(int on) {
  g_timing_on.store(on ? 1 : 0, std::memory_order_relaxed);
  return BVLM_OK;
}

int bvlm_timing_tag_count(void) { return TAG_COUNT; }

const char* bvlm_timing_tag_name(int tag) {
  static const char* names[TAG_COUNT] = {"gemm_diag",    "syrk",     "ggn_rowstats", "ggn_weights", "ggn_moments",
                                         "ggn_stacked",  "quadform", "predictive",   "epig_joint",  "epig_prepare",
                                         "pred_prep",    "probit",   "syrk_prep"};
  return (tag >= 0 && tag < TAG_COUNT) ? names[tag] : "?";
}

int bvlm_timing_collect(int64_t* launches, double* total_ms, int n_tags) {
  if (launches == nullptr || total_ms == nullptr || n_tags < TAG_COUNT) return BVLM_EINVAL;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int i = 0; i < n_tags; ++i) {
    launches[i] = 0;
    total_ms[i] = 0.0;
  }
  int rc = BVLM_OK;
  for (auto& t : g_timed) {
    cudaError_t e = cudaEventSynchronize(t.end);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t.beg, t.end);
    if (e == cudaSuccess) {
      launches[t.tag] += 1;
      total_ms[t.tag] += ms;
    } else {
      rc = static_cast<int>(e);
    }
    cudaEventDestroy(t.beg);
    cudaEventDestroy(t.end);
  }
  g_timed.clear();
  return rc;
}

int bvlm_convert_rows_16(const float* in, int64_t R, int64_t d, int64_t ld, int fmt, void* out, int64_t k_pad,
                         void* stream) {
  if (in == nullptr || out == nullptr || (fmt != FMT_F16 && fmt != FMT_BF16)) return BVLM_EINVAL;
  return launch_rows_to_16(in, R, d, ld, 0, fmt, 0, 1.0f, out, k_pad, nullptr, static_cast<cudaStream_t>(stream));
}

int bvlm_gemm_tn_f32(const void* A16, int64_t M, const void* B16, int64_t N, int64_t k_pad, int fmt, float alpha,
                     float* D, int64_t ldd, int split_k, void* stream) {
  if (A16 == nullptr || B16 == nullptr || D == nullptr || M <= 0 || N <= 0 || k_pad <= 0 || (k_pad % 64) != 0)
    return BVLM_EINVAL;
  if (fmt != FMT_F16 && fmt != FMT_BF16) return BVLM_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int BN = 256;
  CUtensorMap tmA, tmB;
  Operand16 opA{A16, M, k_pad, fmt};
  Operand16 opB{B16, N, k_pad, fmt};
  int rc;
  if ((rc = operand_tmap<GEMM_BM>(&tmA, opA))) return rc;
  if ((rc = operand_tmap<BN>(&tmB, opB))) return rc;
  static const int engine = [] {
    const char* e = getenv("BVLM_DIAG_ENGINE");  // 1: one-CTA engine, 2: CTA-pair engine (4 epilogue warps), 3: pairs + 8 warps
    return e != nullptr ? atoi(e) : 1;
  }();
  if (engine >= 2) {
    if ((rc = operand_tmap<BN / 2>(&tmB, opB))) return rc;
    GemmPlan plan2 = make_plan2<BN>(static_cast<int>(M), static_cast<int>(N), static_cast<int>(k_pad), SCHED_TILES,
                                    split_k < 1 ? 1 : split_k, fmt);
    EpiStoreF32<BN>::Params ep2{D, ldd, alpha, plan2.splits > 1 ? 1 : 0, 0, nullptr, nullptr};
    if (plan2.splits > 1) {
      BVLM_CUDA_TRY(cudaMemset2DAsync(D, static_cast<size_t>(ldd) * 4, 0, static_cast<size_t>(N) * 4, static_cast<size_t>(M), st));
    }
    if (engine == 2) return launch_gemm2<BN, 6, 4, EpiStoreF32<BN>>(tmA, tmB, plan2, ep2, st, TAG_GEMM_DIAG);
    return launch_gemm2<BN, 5, 8, EpiStoreF32<BN>>(tmA, tmB, plan2, ep2, st, TAG_GEMM_DIAG);
  }
  GemmPlan plan = make_plan<BN>(static_cast<int>(M), static_cast<int>(N), static_cast<int>(k_pad), SCHED_TILES,
                                split_k < 1 ? 1 : split_k, fmt, fmt);
  EpiStoreF32<BN>::Params ep{D, ldd, alpha, plan.splits > 1 ? 1 : 0, 0, nullptr, nullptr};
  if (plan.splits > 1) {
    BVLM_CUDA_TRY(cudaMemset2DAsync(D, static_cast<size_t>(ldd) * 4, 0, static_cast<size_t>(N) * 4, static_cast<size_t>(M), st));
  }
  return launch_gemm<BN, 4, EpiStoreF32<BN>>(tmA, tmB, plan, ep, st, TAG_GEMM_DIAG);
}


/* Diagnostics: D[M,N] = alpha * A^T B with MN-major operands A16 [K, lda] (M valid columns), B16 [K, ldb] (N valid
 * columns), K a multiple of 64, through the CTA-pair engine. mode bit 0: A is MN-major (else A16 is [M, K] K-major),
 * bit 1: B is MN-major (else [N, K] K-major). */
int bvlm_gemm_mn_f32(const void* A16, int64_t M, int64_t lda, const void* B16, int64_t N, int64_t ldb, int64_t K, int mode,
                     float alpha, float* D, int64_t ldd, void* stream) {
  if (A16 == nullptr || B16 == nullptr || D == nullptr || M <= 0 || N <= 0 || K <= 0 || (K % 64) != 0) return BVLM_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int BN = 256;
  const bool amn = (mode & 1) != 0, bmn = (mode & 2) != 0;
  CUtensorMap tmA, tmB;
  int rc;
  if (amn) rc = operand_tmap_mn(&tmA, A16, K, M, lda, FMT_F16);
  else {
    Operand16 op{A16, M, K, FMT_F16, lda};
    rc = operand_tmap<GEMM_BM>(&tmA, op);
  }
  if (rc) return rc;
  if (bmn) rc = operand_tmap_mn(&tmB, B16, K, N, ldb, FMT_F16);
  else {
    Operand16 op{B16, N, K, FMT_F16, ldb};
    rc = operand_tmap<BN / 2>(&tmB, op);
  }
  if (rc) return rc;
  GemmPlan plan = make_plan2<BN>(static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), SCHED_TILES, 1, FMT_F16);
  plan.idesc = make_idesc_f16(GEMM2_BM, BN, FMT_F16, FMT_F16, amn, bmn);
  EpiStoreF32<BN>::Params ep{D, ldd, alpha, 0, 0, nullptr, nullptr};
  if (amn && bmn) return launch_gemm2<BN, 6, 4, EpiStoreF32<BN>, true, true>(tmA, tmB, plan, ep, st, TAG_GEMM_DIAG);
  if (amn) return launch_gemm2<BN, 6, 4, EpiStoreF32<BN>, true, false>(tmA, tmB, plan, ep, st, TAG_GEMM_DIAG);
  if (bmn) return launch_gemm2<BN, 6, 4, EpiStoreF32<BN>, false, true>(tmA, tmB, plan, ep, st, TAG_GEMM_DIAG);
  return launch_gemm2<BN, 6, 4, EpiStoreF32<BN>>(tmA, tmB, plan, ep, st, TAG_GEMM_DIAG);
}

}  // extern "C"
