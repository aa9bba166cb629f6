// Operand preparation kernels: fp32 row-major activations / embeddings -> 16-bit (or FP8) GEMM operands (converted,
// normalised / power-of-two scaled, optionally hi/lo split), plus the small per-row statistics the epilogues need.
// All are HBM-bound, one warp per row. Nothing is transposed: products of the X^T X kind use MN-major operands.
#pragma once
#include "common.cuh"

namespace bvlm {

// fp32 [R, d] (row pitch ld) -> 16-bit [R, k_pad]; optional ones column (SigLIP bias, vlm.py:650-654);
// optional per-row power-of-two scaling with row_unscale[r] = 2^(-2 e_r) (undoes the scaling of a squared norm).
int launch_rows_to_16(const float* in, int64_t R, int64_t d, int64_t ld, int append_one, int fmt, int row_pow2_scale,
                      float gmult, void* out, int64_t k_pad, float* row_unscale, cudaStream_t st);

// Predictive row statistics + operand packing (vlm.py:659-668).
//   E_r   = |x_r|^2 + quad_r * sum_diag_self
//   xhat  = x_r / sqrt(E_r) * opscale  -> fp16 [R, seg_pad] (nsplit 1) or [R, 2*seg_pad] = [hi | lo] (nsplit 3)
//   side 0 (source/image):  out0 = s2 * (sum_d x_d^2 * diag_other_d) / E_r ,  out1 = s2 * quad_r / E_r
//   side 1 (target/text) :  out0 = quad_r / E_r ,  out1 = (quad_r * kappa + sum_d diag_other_d * x_d^2) / E_r
int launch_predictive_row_prep(const float* x, int64_t R, int64_t D, int64_t ld, const float* quad,
                               const float* diag_other, float sum_diag_self, float kappa, float s2, int side,
                               int nsplit, float opscale, __half* packed, int64_t seg_pad, int64_t out_pitch,
                               uint8_t* packed8, int64_t seg8, float* out0, float* out1, cudaStream_t st);

// Source side of the predictive, independent of the quadratic forms (so it can overlap them): packed = fp16 [hi | lo] of
// x_r * 2^k_r (exact power-of-two row scale), n2 = |x_r|^2, pd = sum_d x_rd^2 diag_other_d, unscale = 2^-k_r.
// nsplit 2 (fp16 + fp8 error compensation): packed = fp16(x 2^k 32) [R, seg_pad]; packed8 [R, 2*seg8] E4M3 =
// [32 (v - fp16(v)) | v / 32] for the source side (embed prep) and [v / 32 | 32 (v - fp16(v))] for the target side.
int launch_predictive_embed_prep(const float* x, int64_t R, int64_t D, int64_t ld, const float* diag_other, int nsplit,
                                 __half* packed, int64_t seg_pad, int64_t out_pitch, uint8_t* packed8, int64_t seg8, float* n2,
                                 float* pd, float* unscale, const float* act, int64_t d_act, int64_t ld_act, int append_one,
                                 __half* act16, int64_t act_kpad, float* act_unscale, cudaStream_t st);
// (act != NULL additionally converts the row's activations: act16 [R, act_kpad] fp16 with an exact per-row power-of-two
//  scale, act_unscale[r] = 2^(-2 e_r) -- the same operand launch_rows_to_16(row_pow2_scale = 1) produces.)
// GGN row prep (hessians.py:15-21): xhat = x/|x| * opscale -> fp16 [R, d_pad] (nsplit 1) or [R, 2*d_pad] = [hi | lo]
// (nsplit 3; `side` is ignored); inv_norm[r] = 1/|x_r|;
// w_raw[r] = 1/|x_r|^2 ; *w_sum += sum_r w_raw[r] (atomic).
int launch_ggn_row_prep(const float* x, int64_t R, int64_t D, int64_t ld, float opscale, int nsplit, int side, __half* xhat,
                        int64_t d_pad, float* inv_norm, float* w_raw, float* w_sum, cudaStream_t st);

// w[r] = w_raw[r] * R / *w_sum   (mean-one weights keep the fp16 operands of the final GEMM in range)
int launch_normalize_weights(const float* w_raw, const float* w_sum, int64_t R, float* w, cudaStream_t st);

// Per-source finalisation of the pivot-centred GGN (collapsed form of hessians.py:30-46 / 103-113; see kfac.cu).
// gamma = max_c q_c normalises the stacked operands of pass 4 into fp16's range whatever the curvature scale is.
// Outputs are fp16 ROW-MAJOR [B, Dp] blocks (zero padded beyond D) written straight into the stacked MN-major operands:
//   InfoNCE (conditional, rho-free inputs from pass 3): nbar = Nraw*unscale_n, rbar = Rraw*unscale_r, gv = yh[pivot],
//            rho = rest/(1+rest), p* = 1/(1+rest), ebar = nbar - gv, tau = ebar.xh, ubar = rbar - tau (gv + rho ebar),
//            abar = ubar.xh, kappa = g sqrt(w rho / gamma)
//            L_A = -kappa (ebar + gv/(1+sqrt p*)),  R_A = kappa (rho ebar + (1+sqrt p*) gv),
//            L_B = -2 kappa xh,                     R_B = kappa (ubar - abar/2 xh)
//   SigLIP : u = Rraw*unscale_r, a = u.xh, kappa = g sqrt(w / gamma), L_B = -2 kappa xh, R_B = kappa (u - a/2 xh)
int launch_ggn_row_finalize(const float* x, int64_t B, int64_t D, int64_t Dp, int64_t ldx, const float* inv_norm, const float* w,
                            const float* y, int64_t ldy, const float* inv_norm_y, const int* pivot, const float* rest,
                            const float* inv_gamma, const float* Nraw, const float* Rraw, int64_t ldm, float unscale_n,
                            float unscale_r, int siglip, float g, __half* LA, __half* RA, __half* LB, __half* RB, int64_t ldl,
                            int64_t ldr, cudaStream_t st);

// out[c, :] = fp16( yh_c * q_c / gamma * mult ) -- the scaled side of the Yh^T diag(q) Yh segment of pass 4
int launch_ggn_scale_targets(const float* y, int64_t C, int64_t D, int64_t Dp, int64_t ldy, const float* inv_norm_y,
                             const float* q, const float* inv_gamma, float mult, __half* out, int64_t ldo, cudaStream_t st);

// Pass-1 partials [B, S] (row max, rest, pivot per column range) -> merged stats written at offset B*S of each array
// (the arrays hold B*(S+1) entries).
int launch_merge_rowstats(float* m, float* rest, int* piv, int64_t B, int S, cudaStream_t st);

// rowinfo[b] = {m2, log2(rest) - 12, w * rho / 4096, pivot} (InfoNCE) or {0, 0, w / 4096, 0} (SigLIP): pass-2 row constants
int launch_ggn_rowinfo(const float* m2, const float* rest, const int* piv, const float* w, int64_t B, int siglip, float4* out,
                       cudaStream_t st);

// scalars[2] = gamma = max_c q_c (scalars[2] must be zero on entry), then
// scalars[1] = wbar = scalars[0] * inv_count, scalars[3] = wbar * gamma, scalars[4] = 1/gamma (0 when gamma == 0)
int launch_ggn_scalars(const float* q, int64_t C, float* scalars, float inv_count, cudaStream_t st);

// out (+)= alpha * (*alpha_dev) * (S + S^T) / 2
int launch_sym_add(const float* S, int64_t d, int64_t lds, float* out, int64_t ldo, float alpha, const float* alpha_dev,
                   int accumulate, cudaStream_t st);

// out[r, j] = fp16(x[r, j] * scale[j]) row-major with pitch ldo (ones column at j == d when append_one, zero padded)
int launch_scale_cols_f16(const float* x, int64_t n, int64_t d, int64_t ld, const float* scale, int append_one, __half* out,
                          int64_t ldo, cudaStream_t st);

// per-feature power-of-two scale: scale[j] = 2^e_j with max_r |x_rj| 2^e_j in [512,1024); unscale[j] = 2^-e_j
int launch_col_pow2_scale(const float* x, int64_t n, int64_t d, int64_t ld, int append_one, unsigned int* amax_bits,
                          float* scale, float* unscale, cudaStream_t st);

// Canonical probit softmax (scripts/zeroshot.py:119-120): probs = softmax_j(mean / sqrt(1 + pi/8 var)).
int launch_mc_softmax(const float* mean, const float* var, const float* eps, int64_t N, int64_t C, int G, float* acc_p,
                      float* acc_h, cudaStream_t st);
int launch_probit_softmax(const float* mean, const float* var, int64_t N, int64_t C, int64_t ld, float* probs,
                          cudaStream_t st);

// Lower -> upper mirror of a square fp32 matrix, and scale:  A[i,j] = A[j,i] = scale * A[max,min].
int launch_symmetrize_scale(float* A, int64_t d, int64_t ld, float scale, cudaStream_t st);

}  // namespace bvlm
