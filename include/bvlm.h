/*
 * bvlm.h -- C ABI of libbvlm.so: the B200-native (sm_100a) kernels behind BayesVLM's post-hoc Laplace hot path.
 *
 * The reference (MridulPandey17/BayesVLM) is pure PyTorch: its "FFI" for this path is the Python module API of
 * bayesvlm/hessians.py, bayesvlm/vlm.py, bayesvlm/epig.py and scripts/hessian_estimation.py::kfac_ggn.  Each entry
 * point below names the reference call site (file:line in the reference tree) whose ATen ops it replaces; the
 * Python mirror of that API (bayesvlm_b200/) binds these symbols with ctypes.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says host; tensors are row-major, fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work (asynchronous, no host sync);
 *   - `ws` / `ws_bytes` is caller-owned scratch of at least bvlm_*_workspace_bytes(...) bytes, 256-byte aligned;
 *     the library never allocates device memory and never keeps a pointer after the call returns;
 *   - return value: 0 = ok, negative = invalid argument / unsupported / driver entry point missing,
 *     positive = cudaError_t of the failing runtime call.  Nothing throws.  There is no CPU fallback.
 */
#ifndef BVLM_H_
#define BVLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVLM_OK 0
#define BVLM_EINVAL (-1)
#define BVLM_ENOTSUP (-2)
#define BVLM_EDRIVER (-3)
#define BVLM_EWORKSPACE (-4)

/* precision of the mean-logit / SYRK contractions: operands are rounded to 16 bit, accumulation is fp32.
 * BVLM_PREC_X1: single fp16 pass (unit-norm embeddings as they are; raw activations with per-feature power-of-two scaling).
 * BVLM_PREC_X3: hi/lo split, three tensor-core passes (a_hi b_hi + a_lo b_hi + a_hi b_lo), ~2^-22 relative. */
#define BVLM_PREC_X1 1
#define BVLM_PREC_X2F8 2 /* predictive mean only: fp16 hi.hi + the two error-compensation terms lo.hi, hi.lo in FP8 (E4M3)
                            at twice the fp16 tensor rate; ~2^-15 relative, i.e. 2/3 of the cost of BVLM_PREC_X3 */
#define BVLM_PREC_X3 3

const char* bvlm_version(void);
const char* bvlm_status_string(int status);
/* 0 when the current CUDA device is compute capability 10.x and the TMA driver entry point resolves. */
int bvlm_device_check(void);

/* ---------------------------------------------------------------------------------------------------------
 * K1 -- KFAC A factor:  C += alpha * [X 1?]^T [X 1?]      (scripts/hessian_estimation.py:99-104)
 * X [n, d] fp32 row-major (pitch ldx); append_one adds the SigLIP bias column; C [dA, dA] fp32 (pitch ldc),
 * dA = d + append_one.  accumulate = 0 zeroes C first.  Both triangles of C are written.
 * --------------------------------------------------------------------------------------------------------- */
size_t bvlm_syrk_workspace_bytes(int64_t n, int64_t d, int append_one, int precision);
int bvlm_syrk_f32acc(const float* X, int64_t n, int64_t d, int64_t ldx, int append_one, int precision, float* C,
                     int64_t ldc, float alpha, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K2 / K3 -- per-class-batch GGN of the contrastive loss w.r.t. the source embeddings (the KFAC B factor):
 *   H (+)= sum_b J_b^T Hess_b J_b     bayesvlm/hessians.py:10-48 (InfoNCE), :50-117 (SigLIP)
 * X [B, D] sources, Y [C, D] targets (un-normalised), logit_scale in log space (exp applied inside).
 * H [D, D] fp32 (pitch ldh); accumulate = 0 overwrites H.  The increment is exactly symmetric.
 * precision: BVLM_PREC_X1 evaluates the logits s*<xh,yh> from fp16 operands (logit error ~ s*2^-11/sqrt(D) per pair,
 * which averages out over a class batch of thousands of sources); BVLM_PREC_X3 uses the hi/lo split (~fp32 logits)
 * and is what small source batches need (e.g. the single-sample update of bayesvlm/epig.py:242-246).
 * --------------------------------------------------------------------------------------------------------- */
size_t bvlm_ggn_workspace_bytes(int64_t B, int64_t C, int64_t D, int precision);
int bvlm_ggn_infonce(const float* X, int64_t B, int64_t ldx, const float* Y, int64_t C, int64_t ldy, int64_t D,
                     float logit_scale, int precision, float* H, int64_t ldh, int accumulate, void* ws, size_t ws_bytes,
                     void* stream);
int bvlm_ggn_siglip(const float* X, int64_t B, int64_t ldx, const float* Y, int64_t C, int64_t ldy, int64_t D,
                    float logit_scale, float logit_bias, int precision, float* H, int64_t ldh, int accumulate, void* ws,
                    size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * P1 -- quadratic forms  out[i] = a_i^T A_inv a_i   (bayesvlm/vlm.py:662-663, after the optional ones column :650-654)
 * evaluated as |W a_i|^2 with A_inv = W^T W, W lower triangular (host code derives W once per covariance).
 * bvlm_factor_prepare converts W [dA, dA] fp32 to the fp16 operand W16 [dA, k_pad] (k_pad = dA rounded up to 64),
 * multiplied by w_scale (a power of two chosen by the caller so that max|W| * w_scale ~ 2^9).
 * --------------------------------------------------------------------------------------------------------- */
int64_t bvlm_padded_k(int64_t k);
/* colA / colB of the predictive hold bvlm_padded_cols(C) floats (C rounded up to the 256-wide column tile). */
int64_t bvlm_padded_cols(int64_t c);
int bvlm_factor_prepare(const float* W, int64_t dA, int64_t ldw, float w_scale, void* W16, int64_t k_pad, void* stream);
size_t bvlm_quadform_workspace_bytes(int64_t n, int64_t d, int append_one);
int bvlm_quadform(const float* act, int64_t n, int64_t d, int64_t ld, int append_one, const void* W16, int64_t dA,
                  int64_t k_pad, float w_scale, float* out, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * P2 -- Kronecker-Laplace predictive  (bayesvlm/vlm.py:630-684, CLIP._compute_probabilistic_logits_smith)
 *
 * Target (text/class) side, once per (target set, covariance):
 *   gamma_j = t_act_j^T A_txt_inv t_act_j ; E_j = |t_j|^2 + gamma_j * sum(delta)
 *   T16 [C, (prec == 3 ? 2 : 1) * d_pad] packed unit-energy embeddings ([hi | lo] in the split mode), colA_j = gamma_j / E_j, colB_j = (gamma_j kappa + q_j) / E_j
 *   (colA / colB: bvlm_padded_cols(C) floats each, zero beyond C)
 * with beta = diag(B_img_inv), delta = diag(B_txt_inv), kappa = beta.delta, q_j = sum_d beta_d t_jd^2.
 *
 * Source (image) side, per call:
 *   mean [N, C] = exp(logit_scale) * <e_i / sqrt(E_i), t_j / sqrt(E_j)>
 *   var  [N, C] = exp(2 logit_scale) * [ (e_i^2 + alpha_i beta) . (gamma_j delta) + alpha_i beta . t_j^2 ] / (E_i E_j)
 *   probs [N, C] (optional, may be NULL) = softmax_j( mean / sqrt(1 + pi/8 var) )   scripts/zeroshot.py:119-120
 * (logit_bias is NOT added to the probabilistic mean -- vlm.py:681-684.)
 * logit_scale is in log space; when logit_scale_dev (a DEVICE fp32 scalar, e.g. the module parameter itself) is not NULL it
 * is read by the kernel instead of the host value, so in-place parameter updates are always seen and no host sync is needed.
 * --------------------------------------------------------------------------------------------------------- */
size_t bvlm_predictive_target_workspace_bytes(int64_t C, int64_t D, int64_t d_act, int append_one);
/* T8: [C, bvlm_predictive_t8_cols(D)] bytes of E4M3 operands, only for BVLM_PREC_X2F8 (NULL otherwise). */
int64_t bvlm_predictive_t8_cols(int64_t D);
int bvlm_predictive_target_prepare(const float* T, int64_t C, int64_t D, int64_t ldt, const float* Tact, int64_t d_act,
                                   int64_t ldact, int append_one, const void* Wt16, int64_t dA, int64_t k_pad,
                                   float w_scale, const float* beta, float sum_delta, float kappa, int precision,
                                   void* T16, void* T8, float* colA, float* colB, void* ws, size_t ws_bytes, void* stream);
size_t bvlm_predictive_workspace_bytes(int64_t N, int64_t D, int64_t d_act, int append_one, int precision);
int bvlm_predictive(const float* E, int64_t N, int64_t D, int64_t lde, const float* Eact, int64_t d_act, int64_t ldact,
                    int append_one, const void* Wi16, int64_t dA, int64_t k_pad, float w_scale, const float* delta,
                    float sum_beta, float logit_scale, const float* logit_scale_dev, const void* T16, const void* T8,
                    const float* colA, const float* colB, int64_t C, int precision, float* mean, float* var, float* probs,
                    int64_t ldo, void* ws, size_t ws_bytes, void* stream);

/* P3 -- standalone canonical probit softmax (scripts/zeroshot.py:119-120). */
int bvlm_probit_softmax(const float* mean, const float* var, int64_t N, int64_t C, int64_t ld, float* probs,
                        void* stream);

/* T2 -- Monte-Carlo methods of ProbabilisticLogits with a diagonal logit covariance (bayesvlm/vlm.py:68-103 softmax,
 * :142-159 expected_aleatoric_entropy): for G noise draws eps [G, N, C] (one torch.randn call each)
 *   acc_probs [N, C] += softmax(mean + eps_g * sqrt(var)),  acc_entropy [N] += -sum_j p log p   (either may be NULL),
 * added draw by draw onto the values already in the accumulators (the reference's summation order).  C <= 1024, contiguous rows;
 * returns BVLM_ENOTSUP otherwise (the caller keeps the device expression). */
int bvlm_mc_softmax_accumulate(const float* mean, const float* var, const float* eps, int64_t N, int64_t C, int64_t G,
                               float* acc_probs, float* acc_entropy, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * E0 -- MC class probabilities  (bayesvlm/vlm.py:116-123):  probs[n,k,:] = softmax(mean[n,:] + eps[k,n,:] sqrt(var[n,:]))
 * eps [K, N, Cl] fp32 comes from torch.randn under the caller's torch.manual_seed (RNG parity with the reference);
 * probs [N, K, Cl] fp16.
 * E1 -- marginal entropy  (bayesvlm/epig.py:294-311, 275-292) on fp16 probabilities with the rounding points of torch's
 * CUDA kernels (mean over K = fp32 sum * fp32(1/K) -> fp16; xlogy = fp16(x * logf(x)), one rounding; sum -> fp16); out [N] fp16.
 * E2 -- joint-entropy term of EPIG (bayesvlm/epig.py:376-393):
 *   Hjoint[p] = sum_chunks fp16( fp16(-sum_{c,col in chunk} fp16(j log j)) * fp32(1/N_t) ),  j = fp16(fp16(pool @ targ) * fp32(1/K))
 * pool [Np, K, Cl], targ [Nt, K, Cl] fp16; col_chunk = chunk_size of the reference (columns of the flattened (t,c) axis).
 * NOTE: torch's CPU kernels round differently (xlogy rounds log() to Half first); these entry points follow the CUDA
 * kernels, i.e. what the reference computes when it runs on the GPU this library replaces it on.
 *
 * bvlm_epig_prepare_from_noise / _from_probs: E0 + E1 + the permute of epig.py:374-376 in ONE pass over the samples.
 * Every output is optional (NULL): probs16 [N, K, Cl]; oper16 [N, Cl, bvlm_epig_operand_k(K)] = the K-major, zero-padded
 * operand bvlm_epig_joint_entropy_operands consumes; marg16 [N] marginal entropies.
 * bvlm_epig_prepare_supported: 1 when a sample row's tiles fit the kernel's shared memory for this (K, Cl), else 0 (the
 * prepare entry points then return BVLM_ENOTSUP and the caller keeps the generic device expression).
 * --------------------------------------------------------------------------------------------------------- */
int bvlm_epig_operand_k(int64_t K);
int bvlm_epig_prepare_supported(int64_t K, int64_t Cl, int from_noise);
int bvlm_epig_prepare_from_noise(const float* mean, const float* var, const float* eps, int64_t N, int64_t K, int64_t Cl,
                                 void* probs16, void* oper16, void* marg16, void* stream);
int bvlm_epig_prepare_from_probs(const void* probs16, int64_t N, int64_t K, int64_t Cl, void* oper16, void* marg16,
                                 void* stream);
/* two sample sets sharing K and Cl (EPIG: the target set and one pool chunk, epig.py:326-333) in ONE launch */
int bvlm_epig_prepare_pair_from_noise(const float* mean_a, const float* var_a, const float* eps_a, int64_t Na, void* oper_a,
                                      void* marg_a, const float* mean_b, const float* var_b, const float* eps_b, int64_t Nb,
                                      void* oper_b, void* marg_b, int64_t K, int64_t Cl, void* stream);
/* joint-entropy term from the permuted operands; ws: caller-owned scratch of bvlm_epig_joint_operands_workspace_bytes
 * (per-(pool row, column chunk) sums in double), 8-byte aligned.  Any Cl (a pool row's classes may span CTAs). */
size_t bvlm_epig_joint_operands_workspace_bytes(int64_t Np, int64_t Nt, int64_t Cl, int64_t col_chunk);
int bvlm_epig_joint_entropy_operands(const void* poolP, int64_t Np, const void* targP, int64_t Nt, int64_t K, int64_t Cl,
                                     int64_t col_chunk, float* Hjoint, void* ws, size_t ws_bytes, void* stream);
int bvlm_epig_sample_probs(const float* mean, const float* var, const float* eps, int64_t N, int64_t K, int64_t Cl,
                           void* probs16, void* stream);
int bvlm_epig_marginal_entropy_f16(const void* probs16, int64_t N, int64_t K, int64_t Cl, void* out16, void* stream);
size_t bvlm_epig_joint_workspace_bytes(int64_t Np, int64_t Nt, int64_t K, int64_t Cl);
int bvlm_epig_joint_entropy_f16(const void* pool16, int64_t Np, const void* targ16, int64_t Nt, int64_t K, int64_t Cl,
                                int64_t col_chunk, float* Hjoint, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Diagnostics (used by tests/ and bench.py only): plain D = alpha * A B^T through the same tcgen05 engine.
 * A [M, k_pad], B [N, k_pad] 16-bit K-major operands (fmt 0 = fp16, 1 = bf16), D fp32 [M, N] (pitch ldd).
 * --------------------------------------------------------------------------------------------------------- */
int bvlm_gemm_tn_f32(const void* A16, int64_t M, const void* B16, int64_t N, int64_t k_pad, int fmt, float alpha,
                     float* D, int64_t ldd, int split_k, void* stream);
/* D[M,N] = alpha * op(A) op(B)^T through the CTA-pair engine with MN-major operands: mode bit 0 -> A16 is [K, lda]
 * (M valid columns) instead of [M, lda] (K valid columns), bit 1 -> the same for B16. fp16, K a multiple of 64. */
int bvlm_gemm_mn_f32(const void* A16, int64_t M, int64_t lda, const void* B16, int64_t N, int64_t ldb, int64_t K, int mode,
                     float alpha, float* D, int64_t ldd, void* stream);
int bvlm_convert_rows_16(const float* in, int64_t R, int64_t d, int64_t ld, int fmt, void* out, int64_t k_pad,
                         void* stream);
/* number of kernel launches issued through this library since load (for bench.py's gpu_launches claim). */
int64_t bvlm_launch_count(void);
/* Optional CUDA-event timing of every tensor-core (tcgen05) kernel launch, on the stream it is launched on.
 * bvlm_timing_enable(1) starts recording; bvlm_timing_collect synchronises the recorded events, sums launches and
 * milliseconds per kernel tag (arrays of at least bvlm_timing_tag_count() entries) and clears the record. */
int bvlm_timing_enable(int on);
int bvlm_timing_tag_count(void);
const char* bvlm_timing_tag_name(int tag);
int bvlm_timing_collect(int64_t* launches, double* total_ms, int n_tags);

#ifdef __cplusplus
}
#endif
#endif /* BVLM_H_ */
