#include "prep.cuh"

#include <cuda_fp8.h>

#include "rowprep_device.cuh"

namespace bvlm {

namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int ROW_BLOCK = WARPS_PER_BLOCK * 32;

__device__ __forceinline__ uint16_t to_16(float v, int fmt) {
  if (fmt == FMT_BF16) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&b);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

inline unsigned row_grid(int64_t R) { return static_cast<unsigned>((R + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_BLOCK) k_rows_to_16(const float* __restrict__ in, int64_t R, int64_t d, int64_t ld,
                                                          int append_one, int fmt, int row_pow2_scale, float gmult,
                                                          uint16_t* __restrict__ out, int64_t k_pad,
                                                          float* __restrict__ row_unscale) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* x = in + row * ld;
  float sc = gmult;
  if (row_pow2_scale) {
    float amax = append_one ? 1.f : 0.f;
    for (int64_t j = lane; j < d; j += 32) amax = fmaxf(amax, fabsf(x[j]));
    amax = warp_max(amax);
    int e = 0;
    if (amax > 0.f && isfinite(amax)) {
      int ex;
      frexpf(amax, &ex);  // amax = f * 2^ex, f in [0.5, 1)
      e = 10 - ex;        // scaled absmax lands in [512, 1024)
      e = e < -30 ? -30 : (e > 30 ? 30 : e);
    }
    sc = ldexpf(gmult, e);
    if (lane == 0 && row_unscale != nullptr) row_unscale[row] = ldexpf(1.f, -2 * e);
  } else if (lane == 0 && row_unscale != nullptr) {
    row_unscale[row] = 1.f;
  }
  uint16_t* o = out + row * k_pad;
  for (int64_t j = 2 * lane; j < k_pad; j += 64) {
    float v0 = 0.f, v1 = 0.f;
    if (j < d) v0 = x[j] * sc;
    else if (j == d && append_one) v0 = sc;
    if (j + 1 < d) v1 = x[j + 1] * sc;
    else if (j + 1 == d && append_one) v1 = sc;
    const uint32_t pk = static_cast<uint32_t>(to_16(v0, fmt)) | (static_cast<uint32_t>(to_16(v1, fmt)) << 16);
    *reinterpret_cast<uint32_t*>(o + j) = pk;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_BLOCK)
k_predictive_row_prep(const float* __restrict__ x, int64_t R, int64_t D, int64_t ld, const float* __restrict__ quad,
                      const float* __restrict__ diag_other, float sum_diag_self, float kappa, float s2, int side,
                      int nsplit, float opscale, __half* __restrict__ packed, int64_t seg_pad, int64_t out_pitch,
                      uint8_t* __restrict__ packed8, int64_t seg8, float* __restrict__ out0, float* __restrict__ out1) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + row * ld;
  float n2 = 0.f, pd = 0.f;
  for (int64_t j = lane; j < D; j += 32) {
    const float v = xr[j];
    const float v2 = v * v;
    n2 += v2;
    pd = fmaf(v2, diag_other[j], pd);
  }
  n2 = warp_sum(n2);
  pd = warp_sum(pd);
  const float qd = quad[row];
  const float E = n2 + qd * sum_diag_self;
  const float rinv = 1.0f / sqrtf(E);
  const float mul = rinv * opscale;
  // split operands are stored once as [hi | lo]; out_pitch >= that width (0: tight)
  const int64_t pitch = out_pitch > 0 ? out_pitch : seg_pad * (nsplit == 3 ? 2 : 1);
  __half* o = packed + row * pitch;
  for (int64_t j = 2 * lane; j < seg_pad; j += 64) {
    const float v0 = j < D ? xr[j] * mul : 0.f;
    const float v1 = j + 1 < D ? xr[j + 1] * mul : 0.f;
    const __half2 hi = __floats2half2_rn(v0, v1);
    *reinterpret_cast<__half2*>(o + j) = hi;
    if (nsplit == 3) {
      const float2 hf = __half22float2(hi);
      const __half2 lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      *reinterpret_cast<__half2*>(o + seg_pad + j) = lo;
    }
  }
  if (nsplit == 2) {  // fp16 + fp8 error compensation: [x8 | lo8] (B side); see k_predictive_embed_prep
    uint8_t* o8 = packed8 + row * 2 * seg8;
    for (int64_t j = 2 * lane; j < seg8; j += 64) {
      const float v0 = j < D ? xr[j] * mul : 0.f;
      const float v1 = j + 1 < D ? xr[j + 1] * mul : 0.f;
      const float2 hf = __half22float2(__floats2half2_rn(v0, v1));
      *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + j) =
          __nv_cvt_float2_to_fp8x2(make_float2(v0 * 0.03125f, v1 * 0.03125f), __NV_SATFINITE, __NV_E4M3);
      *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + seg8 + j) =
          __nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * 32.f, (v1 - hf.y) * 32.f), __NV_SATFINITE, __NV_E4M3);
    }
  }
  if (lane == 0) {
    if (side == 0) {
      out0[row] = s2 * pd / E;
      out1[row] = s2 * qd / E;
    } else {
      out0[row] = qd / E;
      out1[row] = (qd * kappa + pd) / E;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Source-side operand of the predictive that does NOT depend on the quadratic forms: e_i * 2^k (row power-of-two scale,
// exact) as fp16 [hi | lo], plus |e_i|^2, sum_d e_id^2 delta_d and 2^-k.  The 1/sqrt(E_i) normalisation is applied by the
// GEMM epilogue (a per-row factor), so this pass can overlap the quadratic-form GEMM on another stream.
__global__ void __launch_bounds__(ROW_BLOCK)
k_predictive_embed_prep(const float* __restrict__ x, int64_t R, int64_t D, int64_t ld, const float* __restrict__ diag_other,
                        int nsplit, __half* __restrict__ packed, int64_t seg_pad, int64_t out_pitch,
                        uint8_t* __restrict__ packed8, int64_t seg8, float* __restrict__ n2_out,
                        float* __restrict__ pd_out, float* __restrict__ unscale_out,
                        // optional fused activation conversion (the quadratic-form operand of the same row)
                        const float* __restrict__ act, int64_t d_act, int64_t ld_act, int append_one,
                        __half* __restrict__ act16, int64_t act_kpad, float* __restrict__ act_unscale) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  if (act != nullptr) {  // fp16 activations with an exact per-row power-of-two scale (as k_rows_to_16)
    const float* ar = act + row * ld_act;
    float am = append_one ? 1.f : 0.f;
    for (int64_t j = lane; j < d_act; j += 32) am = fmaxf(am, fabsf(ar[j]));
    am = warp_max(am);
    int ea = 0;
    if (am > 0.f && isfinite(am)) {
      int ex;
      frexpf(am, &ex);
      ea = 10 - ex;
      ea = ea < -30 ? -30 : (ea > 30 ? 30 : ea);
    }
    const float sa = ldexpf(1.f, ea);
    if (lane == 0) act_unscale[row] = ldexpf(1.f, -2 * ea);
    __half* oa = act16 + row * act_kpad;
    for (int64_t j = 2 * lane; j < act_kpad; j += 64) {
      float v0 = 0.f, v1 = 0.f;
      if (j < d_act) v0 = ar[j] * sa;
      else if (j == d_act && append_one) v0 = sa;
      if (j + 1 < d_act) v1 = ar[j + 1] * sa;
      else if (j + 1 == d_act && append_one) v1 = sa;
      *reinterpret_cast<__half2*>(oa + j) = __floats2half2_rn(v0, v1);
    }
  }
  if (x == nullptr) return;  // embeddings converted elsewhere (epilogue warps of the quadratic-form GEMM)
  const float* xr = x + row * ld;
  float n2 = 0.f, pd = 0.f, amax = 0.f;
  for (int64_t j = lane; j < D; j += 32) {
    const float v = xr[j];
    const float v2 = v * v;
    n2 += v2;
    pd = fmaf(v2, diag_other[j], pd);
    amax = fmaxf(amax, fabsf(v));
  }
  n2 = warp_sum(n2);
  pd = warp_sum(pd);
  amax = warp_max(amax);
  int e = 0;
  if (amax > 0.f && isfinite(amax)) {
    int ex;
    frexpf(amax, &ex);
    e = (nsplit == 2 ? 8 : 9) - ex;  // scaled absmax lands in [256, 512) ([128, 256) when the values also go to fp8)
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
  }
  // fp16 + fp8 mode: the fp16 operand carries an extra 2^5 so that hi.hi, lo8.t8 and e8.tlo8 share the scale 2^10
  const float sc = ldexpf(1.f, e) * (nsplit == 2 ? 32.f : 1.f);
  const int64_t pitch = out_pitch > 0 ? out_pitch : seg_pad * (nsplit == 3 ? 2 : 1);
  __half* o = packed + row * pitch;
  for (int64_t j = 2 * lane; j < seg_pad; j += 64) {
    const float v0 = j < D ? xr[j] * sc : 0.f;
    const float v1 = j + 1 < D ? xr[j + 1] * sc : 0.f;
    const __half2 hi = __floats2half2_rn(v0, v1);
    *reinterpret_cast<__half2*>(o + j) = hi;
    if (nsplit == 3) {
      const float2 hf = __half22float2(hi);
      *reinterpret_cast<__half2*>(o + seg_pad + j) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    }
  }
  if (nsplit == 2) {  // A side of the FP8 phase: [lo8 | x8], paired with the target's [t8 | tlo8]
    uint8_t* o8 = packed8 + row * 2 * seg8;
    for (int64_t j = 2 * lane; j < seg8; j += 64) {
      const float v0 = j < D ? xr[j] * sc : 0.f;
      const float v1 = j + 1 < D ? xr[j + 1] * sc : 0.f;
      const float2 hf = __half22float2(__floats2half2_rn(v0, v1));
      *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + j) =
          __nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * 32.f, (v1 - hf.y) * 32.f), __NV_SATFINITE, __NV_E4M3);
      *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + seg8 + j) =
          __nv_cvt_float2_to_fp8x2(make_float2(v0 * 0.03125f, v1 * 0.03125f), __NV_SATFINITE, __NV_E4M3);
    }
  }
  if (lane == 0) {
    n2_out[row] = n2;
    pd_out[row] = pd;
    unscale_out[row] = ldexpf(1.f, -e);  // (the extra 2^5 of the fp16 + fp8 mode is folded into the caller's mean_scale)
  }
}

// Register-resident variant of k_predictive_embed_prep for 16-byte aligned rows: every lane keeps its float4 slices of
// the row in registers (one global read per element, 512 contiguous bytes per warp instruction), EV / AV = float4 slices
// per lane for the embedding / activation row (rows up to 128*EV / 128*AV floats).
template <int EV, int AV>
__global__ void __launch_bounds__(ROW_BLOCK)
k_predictive_prep_vec(const float* __restrict__ x, int64_t R, int64_t D, int64_t ld, const float* __restrict__ diag_other,
                      int nsplit, __half* __restrict__ packed, int64_t seg_pad, int64_t out_pitch,
                      uint8_t* __restrict__ packed8, int64_t seg8, float* __restrict__ n2_out, float* __restrict__ pd_out,
                      float* __restrict__ unscale_out, const float* __restrict__ act, int64_t d_act, int64_t ld_act,
                      int append_one, __half* __restrict__ act16, int64_t act_kpad, float* __restrict__ act_unscale) {
  // the quadratic-form GEMM is launched programmatically behind this kernel: its CTAs may become resident under this kernel's tail.
  // (Launching THIS kernel programmatically behind the previous step's GEMM as well was measured: its resident, waiting blocks
  //  cost the GEMM more than the overlap gains -- 0.2966 vs 0.2913 ms per step.)
  grid_launch_dependents();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  // ---------------- activations -> fp16, exact per-row power-of-two scale
  {
    const float* ar = act + row * ld_act;
    float4 a[AV];
    float am = append_one ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < AV; ++i) {
      const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * 4;
      a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 3 < d_act) a[i] = __ldg(reinterpret_cast<const float4*>(ar + c));
      else if (c < d_act) {
        a[i].x = ar[c];
        if (c + 1 < d_act) a[i].y = ar[c + 1];
        if (c + 2 < d_act) a[i].z = ar[c + 2];
      }
      am = fmaxf(am, fmaxf(fmaxf(fabsf(a[i].x), fabsf(a[i].y)), fmaxf(fabsf(a[i].z), fabsf(a[i].w))));
    }
    am = warp_max(am);
    int ea = 0;
    if (am > 0.f && isfinite(am)) {
      int ex;
      frexpf(am, &ex);
      ea = 10 - ex;
      ea = ea < -30 ? -30 : (ea > 30 ? 30 : ea);
    }
    const float sa = ldexpf(1.f, ea);
    if (lane == 0) act_unscale[row] = ldexpf(1.f, -2 * ea);
    __half* oa = act16 + row * act_kpad;
#pragma unroll
    for (int i = 0; i < AV; ++i) {
      const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * 4;
      if (c >= act_kpad) continue;
      float v[4] = {a[i].x * sa, a[i].y * sa, a[i].z * sa, a[i].w * sa};
      if (append_one && d_act >= c && d_act < c + 4) v[d_act - c] = sa;
      const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(oa + c) = pk;
    }
  }
  // ---------------- embeddings -> fp16 (+ lo fp16 | + fp8 compensation terms), row statistics
  // (skipped when the quadratic-form GEMM converts them in its epilogue warps: x == nullptr)
  if (x == nullptr) return;
  EmbedPrepArgs ea{x, R, D, ld, diag_other, nsplit, packed, seg_pad, out_pitch > 0 ? out_pitch : seg_pad * (nsplit == 3 ? 2 : 1),
                   packed8, seg8, n2_out, pd_out, unscale_out};
  float4 e[EV];
  embed_row_load<EV>(ea, row, lane, e);
  embed_row_finish<EV>(ea, row, lane, e);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_BLOCK)
k_ggn_row_prep(const float* __restrict__ x, int64_t R, int64_t D, int64_t ld, float opscale, int nsplit, int side,
               __half* __restrict__ xhat, int64_t d_pad, float* __restrict__ inv_norm, float* __restrict__ w_raw,
               float* __restrict__ w_sum) {
  __shared__ float s_w[WARPS_PER_BLOCK];
  const int warp = threadIdx.x >> 5;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + warp;
  const int lane = threadIdx.x & 31;
  float wr = 0.f;
  if (row < R) {
    const float* xr = x + row * ld;
    float n2 = 0.f;
    for (int64_t j = lane; j < D; j += 32) n2 = fmaf(xr[j], xr[j], n2);
    n2 = warp_sum(n2);
    const float inv = 1.0f / sqrtf(n2);
    const float mul = inv * opscale;
    __half* o = xhat + row * d_pad * (nsplit == 3 ? 2 : 1);  // split operands are stored once as [hi | lo]
    for (int64_t j = 2 * lane; j < d_pad; j += 64) {
      const float v0 = j < D ? xr[j] * mul : 0.f;
      const float v1 = j + 1 < D ? xr[j + 1] * mul : 0.f;
      const __half2 hi = __floats2half2_rn(v0, v1);
      *reinterpret_cast<__half2*>(o + j) = hi;
      if (nsplit == 3) {  // hi/lo split; the GEMM plan forms hi.hi + lo.hi + hi.lo from the two stored segments
        const float2 hf = __half22float2(hi);
        *reinterpret_cast<__half2*>(o + d_pad + j) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      }
    }
    wr = inv * inv;
    if (lane == 0) {
      inv_norm[row] = inv;
      if (w_raw != nullptr) w_raw[row] = wr;
    }
  }
  if (w_sum != nullptr) {
    if (lane == 0) s_w[warp] = wr;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < WARPS_PER_BLOCK; ++i) t += s_w[i];
      atomicAdd(w_sum, t);
    }
  }
}

__global__ void k_normalize_weights(const float* __restrict__ w_raw, const float* __restrict__ w_sum, int64_t R,
                                    float* __restrict__ w) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < R) w[i] = w_raw[i] * (static_cast<float>(R) / *w_sum);
}

// ------------------------------------------------------------------------------------------------
// Per-source finalisation of the pivot-centred GGN (see kfac.cu). One warp per source row; writes the fp16 ROW-MAJOR
// stacked operands of pass 4 directly (they are MN-major operands of the tensor-core kernel: no transposes).
__device__ __forceinline__ void store_row_f16(__half* dst, int64_t j, int64_t D, float v0, float v1) {
  // j even; pads beyond D with zeros
  *reinterpret_cast<__half2*>(dst + j) = __floats2half2_rn(j < D ? v0 : 0.f, j + 1 < D ? v1 : 0.f);
}

__global__ void __launch_bounds__(ROW_BLOCK)
k_ggn_row_finalize(const float* __restrict__ x, int64_t B, int64_t D, int64_t Dp, int64_t ldx,
                   const float* __restrict__ inv_norm, const float* __restrict__ w, const float* __restrict__ y,
                   int64_t ldy, const float* __restrict__ inv_norm_y, const int* __restrict__ pivot,
                   const float* __restrict__ rest, const float* __restrict__ inv_gamma, const float* __restrict__ Nraw,
                   const float* __restrict__ Rraw, int64_t ldm, float unscale_n, float unscale_r, int siglip, float g,
                   __half* __restrict__ LA, __half* __restrict__ RA, __half* __restrict__ LB, __half* __restrict__ RB,
                   int64_t ldl, int64_t ldr) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float inv = inv_norm[row];
  const float* xr = x + row * ldx;
  const float* rr = Rraw + row * ldm;
  __half* lb = LB + row * ldl;
  __half* rb = RB + row * ldr;
  if (siglip) {
    // u = r (cosine-weighted), a = u.xh ;  L_B = -2 kappa xh,  R_B = kappa (u - a/2 xh),  kappa = sqrt(w / gamma)
    const float kappa = g * sqrtf(fmaxf(w[row] * (*inv_gamma), 0.f));
    float a = 0.f;
    for (int64_t j = lane; j < D; j += 32) a = fmaf(rr[j] * unscale_r, xr[j] * inv, a);
    a = warp_sum(a);
    for (int64_t j = 2 * lane; j < Dp; j += 64) {
      float lv[2], rv[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t jj = j + t;
        const float xh = jj < D ? xr[jj] * inv : 0.f;
        const float u = jj < D ? rr[jj] * unscale_r : 0.f;
        lv[t] = -2.f * kappa * xh;
        rv[t] = kappa * fmaf(-0.5f * a, xh, u);
      }
      store_row_f16(lb, j, D, lv[0], lv[1]);
      store_row_f16(rb, j, D, rv[0], rv[1]);
    }
    return;
  }
  // conditional (rho-free) quantities: nbar = E[yh | c != pivot], ebar = nbar - g, rbar = E[d yh | c != pivot] ...
  const int pv = pivot[row];
  const float* gr = y + static_cast<int64_t>(pv) * ldy;
  const float ginv = inv_norm_y[pv];
  const float rs = rest[row];
  const float rho = rs / (1.f + rs);         // 1 - softmax(pivot), no cancellation
  const float sq = sqrtf(1.f / (1.f + rs));  // sqrt(softmax(pivot))
  const float c1 = 1.f / (1.f + sq);         // (1 - sqrt p*) / rho
  const float c2 = 1.f + sq;
  const float kappa = g * sqrtf(fmaxf(w[row] * rho * (*inv_gamma), 0.f));
  const float* nr = Nraw + row * ldm;
  __half* la = LA + row * ldl;
  __half* ra = RA + row * ldr;
  float tau = 0.f;  // ebar . xh
  for (int64_t j = lane; j < D; j += 32) {
    const float e = fmaf(nr[j], unscale_n, -gr[j] * ginv);
    tau = fmaf(e, xr[j] * inv, tau);
  }
  tau = warp_sum(tau);
  float a = 0.f;  // ubar . xh,  ubar = rbar - tau (g + rho ebar)
  for (int64_t j = lane; j < D; j += 32) {
    const float gg = gr[j] * ginv;
    const float e = fmaf(nr[j], unscale_n, -gg);
    const float u = fmaf(-tau, fmaf(rho, e, gg), rr[j] * unscale_r);
    a = fmaf(u, xr[j] * inv, a);
  }
  a = warp_sum(a);
  for (int64_t j = 2 * lane; j < Dp; j += 64) {
    float v[4][2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int64_t jj = j + t;
      const bool ok = jj < D;
      const float gg = ok ? gr[jj] * ginv : 0.f;
      const float xh = ok ? xr[jj] * inv : 0.f;
      const float e = ok ? fmaf(nr[jj], unscale_n, -gg) : 0.f;
      const float u = ok ? fmaf(-tau, fmaf(rho, e, gg), rr[jj] * unscale_r) : 0.f;
      v[0][t] = -kappa * fmaf(c1, gg, e);              // L_A
      v[1][t] = kappa * fmaf(rho, e, c2 * gg);         // R_A
      v[2][t] = -2.f * kappa * xh;                     // L_B
      v[3][t] = kappa * fmaf(-0.5f * a, xh, u);        // R_B
    }
    store_row_f16(la, j, D, v[0][0], v[0][1]);
    store_row_f16(ra, j, D, v[1][0], v[1][1]);
    store_row_f16(lb, j, D, v[2][0], v[2][1]);
    store_row_f16(rb, j, D, v[3][0], v[3][1]);
  }
}

// 16-byte vectorised variant of k_ggn_row_finalize (D % 4 == 0, aligned rows): same math, a quarter of the load
// instructions; the three sweeps over the row hit L1 after the first.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4h(__half* dst, float a, float b, float c, float d) {
  const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0);
  pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(dst) = pk;
}
__global__ void __launch_bounds__(ROW_BLOCK)
k_ggn_row_finalize_vec(const float* __restrict__ x, int64_t B, int64_t D, int64_t Dp, int64_t ldx,
                       const float* __restrict__ inv_norm, const float* __restrict__ w, const float* __restrict__ y,
                       int64_t ldy, const float* __restrict__ inv_norm_y, const int* __restrict__ pivot,
                       const float* __restrict__ rest, const float* __restrict__ inv_gamma, const float* __restrict__ Nraw,
                       const float* __restrict__ Rraw, int64_t ldm, float unscale_n, float unscale_r, int siglip, float g,
                       __half* __restrict__ LA, __half* __restrict__ RA, __half* __restrict__ LB, __half* __restrict__ RB,
                       int64_t ldl, int64_t ldr) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float inv = inv_norm[row];
  const float* xr = x + row * ldx;
  const float* rr = Rraw + row * ldm;
  __half* lb = LB + row * ldl;
  __half* rb = RB + row * ldr;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (siglip) {
    const float kappa = g * sqrtf(fmaxf(w[row] * (*inv_gamma), 0.f));
    float a = 0.f;
    for (int64_t j = 4 * lane; j < D; j += 128) {
      const float4 r4 = ld4(rr + j), x4 = ld4(xr + j);
      a = fmaf(r4.x, x4.x, fmaf(r4.y, x4.y, fmaf(r4.z, x4.z, fmaf(r4.w, x4.w, a))));
    }
    a = warp_sum(a) * unscale_r * inv;
    for (int64_t j = 4 * lane; j < Dp; j += 128) {
      const bool ok = j < D;
      const float4 r4 = ok ? ld4(rr + j) : z4, x4 = ok ? ld4(xr + j) : z4;
      const float k2 = -2.f * kappa * inv, ha = -0.5f * a * inv;
      st4h(lb + j, k2 * x4.x, k2 * x4.y, k2 * x4.z, k2 * x4.w);
      st4h(rb + j, kappa * fmaf(ha, x4.x, r4.x * unscale_r), kappa * fmaf(ha, x4.y, r4.y * unscale_r),
           kappa * fmaf(ha, x4.z, r4.z * unscale_r), kappa * fmaf(ha, x4.w, r4.w * unscale_r));
    }
    return;
  }
  const int pv = pivot[row];
  const float* gr = y + static_cast<int64_t>(pv) * ldy;
  const float ginv = inv_norm_y[pv];
  const float rs = rest[row];
  const float rho = rs / (1.f + rs);
  const float sq = sqrtf(1.f / (1.f + rs));
  const float c1 = 1.f / (1.f + sq), c2 = 1.f + sq;
  const float kappa = g * sqrtf(fmaxf(w[row] * rho * (*inv_gamma), 0.f));
  const float* nr = Nraw + row * ldm;
  __half* la = LA + row * ldl;
  __half* ra = RA + row * ldr;
  float tau = 0.f;  // ebar . xh
  for (int64_t j = 4 * lane; j < D; j += 128) {
    const float4 n4 = ld4(nr + j), g4 = ld4(gr + j), x4 = ld4(xr + j);
    tau = fmaf(fmaf(n4.x, unscale_n, -g4.x * ginv), x4.x, tau);
    tau = fmaf(fmaf(n4.y, unscale_n, -g4.y * ginv), x4.y, tau);
    tau = fmaf(fmaf(n4.z, unscale_n, -g4.z * ginv), x4.z, tau);
    tau = fmaf(fmaf(n4.w, unscale_n, -g4.w * ginv), x4.w, tau);
  }
  tau = warp_sum(tau) * inv;
  float a = 0.f;  // ubar . xh
#define BVLM_U(nn, gg, rrv) fmaf(-tau, fmaf(rho, fmaf(nn, unscale_n, -(gg) * ginv), (gg) * ginv), (rrv) * unscale_r)
  for (int64_t j = 4 * lane; j < D; j += 128) {
    const float4 n4 = ld4(nr + j), g4 = ld4(gr + j), x4 = ld4(xr + j), r4 = ld4(rr + j);
    a = fmaf(BVLM_U(n4.x, g4.x, r4.x), x4.x, a);
    a = fmaf(BVLM_U(n4.y, g4.y, r4.y), x4.y, a);
    a = fmaf(BVLM_U(n4.z, g4.z, r4.z), x4.z, a);
    a = fmaf(BVLM_U(n4.w, g4.w, r4.w), x4.w, a);
  }
  a = warp_sum(a) * inv;
  for (int64_t j = 4 * lane; j < Dp; j += 128) {
    const bool ok = j < D;
    const float4 n4 = ok ? ld4(nr + j) : z4, g4 = ok ? ld4(gr + j) : z4, x4 = ok ? ld4(xr + j) : z4, r4 = ok ? ld4(rr + j) : z4;
    float vla[4], vra[4], vlb[4], vrb[4];
    const float nn[4] = {n4.x, n4.y, n4.z, n4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w},
                rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float gv = gg[t] * ginv, xh = xx[t] * inv;
      const float e = ok ? fmaf(nn[t], unscale_n, -gv) : 0.f;
      const float u = fmaf(-tau, fmaf(rho, e, gv), rv[t] * unscale_r);
      vla[t] = -kappa * fmaf(c1, gv, e);
      vra[t] = kappa * fmaf(rho, e, c2 * gv);
      vlb[t] = -2.f * kappa * xh;
      vrb[t] = kappa * fmaf(-0.5f * a, xh, u);
    }
    st4h(la + j, vla[0], vla[1], vla[2], vla[3]);
    st4h(ra + j, vra[0], vra[1], vra[2], vra[3]);
    st4h(lb + j, vlb[0], vlb[1], vlb[2], vlb[3]);
    st4h(rb + j, vrb[0], vrb[1], vrb[2], vrb[3]);
  }
#undef BVLM_U
}

// out[c, :] = fp16( y_c / |y_c| * q_c * inv_gamma * mult ), zero padded to Dp columns: the scaled side of Yh^T diag(q) Yh
__global__ void __launch_bounds__(ROW_BLOCK)
k_ggn_scale_targets(const float* __restrict__ y, int64_t C, int64_t D, int64_t Dp, int64_t ldy,
                    const float* __restrict__ inv_norm_y, const float* __restrict__ q, const float* __restrict__ inv_gamma,
                    float mult, __half* __restrict__ out, int64_t ldo) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= C) return;
  const float m = inv_norm_y[row] * fmaxf(q[row] * (*inv_gamma), 0.f) * mult;
  const float* yr = y + row * ldy;
  __half* o = out + row * ldo;
  for (int64_t j = 2 * lane; j < Dp; j += 64)
    store_row_f16(o, j, D, j < D ? yr[j] * m : 0.f, j + 1 < D ? yr[j + 1] * m : 0.f);
}

// combine the per-column-range online-softmax partials of pass 1: [B, S] -> [B] (written to the first B entries)
__global__ void k_merge_rowstats(float* __restrict__ m, float* __restrict__ rest, int* __restrict__ piv, int64_t B, int S) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float mm = m[b * S], rr = rest[b * S];
  int pp = piv[b * S];
  for (int s = 1; s < S; ++s) {
    const float m2 = m[b * S + s], r2 = rest[b * S + s];
    const int p2 = piv[b * S + s];
    if (m2 > mm) {
      rr = r2 + (rr + 1.f) * exp2f(mm - m2);
      mm = m2;
      pp = p2;
    } else if (m2 > -INFINITY) {
      rr = rr + (r2 + 1.f) * exp2f(m2 - mm);
    }
  }
  m[B * S + b] = mm;  // merged stats live behind the partials (no aliasing with rows other blocks still read)
  rest[B * S + b] = rr;
  piv[B * S + b] = pp;
}

// per-source constants of pass 2, packed so that the epilogue needs ONE 16-byte load per row and tile:
//   InfoNCE: {m2, lgw = log2(rest) - 12 (+inf if rest == 0), wq = w * rest/(1+rest) / 4096, pivot}   SigLIP: {0, 0, w/4096, 0}
__global__ void k_ggn_rowinfo(const float* __restrict__ m2, const float* __restrict__ rest, const int* __restrict__ piv,
                              const float* __restrict__ w, int64_t B, int siglip, float4* __restrict__ out) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float wscale_inv = 1.0f / 4096.f;
  if (siglip) {
    out[b] = make_float4(0.f, 0.f, w[b] * wscale_inv, 0.f);
  } else {
    const float rs = rest[b];
    out[b] = make_float4(m2[b], rs > 0.f ? log2f(rs) - 12.f : INFINITY, w[b] * (rs / (1.f + rs)) * wscale_inv,
                         __int_as_float(piv[b]));
  }
}

// gamma = max_c q_c (q >= 0): atomicMax on the float bit pattern
__global__ void k_max_nonneg(const float* __restrict__ q, int64_t n, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = q[i];
    if (v > m && isfinite(v)) m = v;
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

// scalars: [0] sum_b 1/|x_b|^2, [1] wbar, [2] gamma, [3] wbar * gamma, [4] 1/gamma (0 if gamma == 0)
__global__ void k_ggn_scalars(float* __restrict__ sc, float inv_count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float wbar = sc[0] * inv_count;
    const float gamma = sc[2];
    sc[1] = wbar;
    sc[3] = wbar * gamma;
    sc[4] = gamma > 0.f ? 1.0f / gamma : 0.f;
  }
}

// out (+)= alpha * alpha_dev * (S + S^T) / 2
__global__ void k_sym_add(const float* __restrict__ S, int64_t d, int64_t lds, float* __restrict__ out, int64_t ldo,
                          float alpha, const float* __restrict__ alpha_dev, int accumulate) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= d || i >= d) return;
  const float a = alpha_dev != nullptr ? alpha * (*alpha_dev) : alpha;
  const float v = 0.5f * a * (S[i * lds + j] + S[j * lds + i]);
  out[i * ldo + j] = accumulate ? out[i * ldo + j] + v : v;
}

// per-feature |max| over the rows (atomicMax on the bit pattern of non-negative floats)
__global__ void __launch_bounds__(256) k_col_absmax(const float* __restrict__ x, int64_t n, int64_t d, int64_t ld,
                                                    int rows_per_block, unsigned int* __restrict__ amax_bits) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  if (j >= d) return;
  float m = 0.f;
  const int64_t r1 = r0 + rows_per_block < n ? r0 + rows_per_block : n;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) m = fmaxf(m, fabsf(x[r * ld + j]));
  if (m > 0.f && isfinite(m)) atomicMax(amax_bits + j, __float_as_uint(m));
}

// out[r, j] = fp16(x[r, j] * scale[j]) row-major [n, ldo] (a ones column at j == d when append_one, zeros up to ldo):
// the MN-major SYRK operand -- no transpose. One warp per row, 16-byte loads when the rows are aligned.
__global__ void __launch_bounds__(ROW_BLOCK)
k_scale_cols_f16(const float* __restrict__ x, int64_t n, int64_t d, int64_t ld, const float* __restrict__ scale,
                 int append_one, __half* __restrict__ out, int64_t ldo, int vec_ok) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const float* xr = x + r * ld;
  __half* o = out + r * ldo;
  for (int64_t j = 4 * lane; j < ldo; j += 128) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec_ok && j + 3 < d) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + j));
      const float4 sv = __ldg(reinterpret_cast<const float4*>(scale + j));
      v[0] = xv.x * sv.x, v[1] = xv.y * sv.y, v[2] = xv.z * sv.z, v[3] = xv.w * sv.w;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t jj = j + t;
        if (jj < d) v[t] = xr[jj] * scale[jj];
        else if (jj == d && append_one) v[t] = scale[jj];
      }
    }
    const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    if (vec_ok && j + 3 < ldo) {
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(o + j) = pk;
    } else {  // ldo is even: at most one trailing pair
      *reinterpret_cast<__half2*>(o + j) = h0;
      if (j + 2 < ldo) *reinterpret_cast<__half2*>(o + j + 2) = h1;
    }
  }
}

// scale[j] = 2^e with absmax * 2^e in [512, 1024); unscale[j] = 2^-e. Feature d (the ones column) gets absmax 1.
__global__ void k_col_pow2_scale(const unsigned int* __restrict__ amax_bits, int64_t d, int append_one,
                                 float* __restrict__ scale, float* __restrict__ unscale) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= d + (append_one ? 1 : 0)) return;
  const float amax = j < d ? __uint_as_float(amax_bits[j]) : 1.f;
  int e = 0;
  if (amax > 0.f) {
    int ex;
    frexpf(amax, &ex);
    e = 10 - ex;
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
  }
  scale[j] = ldexpf(1.f, e);
  unscale[j] = ldexpf(1.f, -e);
}

// ------------------------------------------------------------------------------------------------
// P3: probs = softmax_j( mean_j / sqrt(1 + pi/8 var_j) )  (scripts/zeroshot.py:119-120).  One warp per row; rows of up to 1024
// classes make ONE pass over HBM: every load of the row (32 + 32 per lane, 128-bit when the row pitch allows) is issued
// before the first dependent instruction, so a warp keeps 8 KB in flight; z is evaluated with MUFU.RSQ / MUFU.EX2
// (the precise division + sqrt of the first version serialised loads and arithmetic: 2.9 TB/s, bench r2 `with_probs`).
template <bool VEC4>
__global__ void __launch_bounds__(ROW_BLOCK)
k_probit_softmax(const float* __restrict__ mean, const float* __restrict__ var, int64_t N, int64_t C, int64_t ld,
                 float* __restrict__ probs) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* m = mean + row * ld;
  const float* v = var + row * ld;
  float* p = probs + row * ld;
  constexpr float kPi8 = 0.39269908169872414f;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int CACHE = 32;  // rows up to 1024 classes stay in registers
  if (C <= CACHE * 32) {
    float z[CACHE], w[CACHE];
    if constexpr (VEC4) {  // lane owns columns 4 * (lane + 32 i) .. + 3
#pragma unroll
      for (int i = 0; i < CACHE / 4; ++i) {
        const int64_t j = 4 * (lane + 32 * i);
        const int64_t jc = j < C ? j : 0;  // C is a multiple of 4 here; clamped, not predicated: the loads hoist freely
        const float4 a = __ldcs(reinterpret_cast<const float4*>(m + jc));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(v + jc));
        z[4 * i] = a.x, z[4 * i + 1] = a.y, z[4 * i + 2] = a.z, z[4 * i + 3] = a.w;
        w[4 * i] = b.x, w[4 * i + 1] = b.y, w[4 * i + 2] = b.z, w[4 * i + 3] = b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CACHE; ++i) {
        const int64_t j = lane + 32 * i;
        const int64_t jc = j < C ? j : 0;
        z[i] = __ldcs(m + jc);
        w[i] = __ldcs(v + jc);
      }
    }
    asm volatile("" ::: "memory");  // every load of the row is issued before the first use
    float zmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < CACHE; ++i) {
      const int64_t j = VEC4 ? 4 * (lane + 32 * (i / 4)) + (i & 3) : lane + 32 * i;
      z[i] = j < C ? z[i] * (rsqrtf(fmaf(kPi8, w[i], 1.0f)) * kLog2e) : -INFINITY;
      zmax = fmaxf(zmax, z[i]);
    }
    zmax = warp_max(zmax);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CACHE; ++i) {
      z[i] = fast_exp2(z[i] - zmax);
      s += z[i];
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    if constexpr (VEC4) {
#pragma unroll
      for (int i = 0; i < CACHE / 4; ++i) {
        const int64_t j = 4 * (lane + 32 * i);
        if (j < C)
          __stcs(reinterpret_cast<float4*>(p + j),
                 make_float4(z[4 * i] * inv, z[4 * i + 1] * inv, z[4 * i + 2] * inv, z[4 * i + 3] * inv));
      }
    } else {
#pragma unroll
      for (int i = 0; i < CACHE; ++i) {
        const int64_t j = lane + 32 * i;
        if (j < C) __stcs(p + j, z[i] * inv);
      }
    }
  } else {
    float zmax = -INFINITY;
    for (int64_t j = lane; j < C; j += 32) zmax = fmaxf(zmax, m[j] * (rsqrtf(fmaf(kPi8, v[j], 1.0f)) * kLog2e));
    zmax = warp_max(zmax);
    float s = 0.f;
    for (int64_t j = lane; j < C; j += 32) s += fast_exp2(m[j] * (rsqrtf(fmaf(kPi8, v[j], 1.0f)) * kLog2e) - zmax);
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int64_t j = lane; j < C; j += 32) p[j] = fast_exp2(m[j] * (rsqrtf(fmaf(kPi8, v[j], 1.0f)) * kLog2e) - zmax) * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// Monte-Carlo methods of ProbabilisticLogits (vlm.py:68-103 softmax, :142-159 expected_aleatoric_entropy) for a diagonal
// logit covariance: for G noise draws eps_g [N, C] (one torch.randn call each: the reference's RNG contract)
//     p_g = softmax(mean + eps_g * std)        acc_p += p_g        acc_h += -sum_j p_gj log p_gj
// One warp per row, the row's mean / std / running sums in registers (C <= 32 * NI), one pass over the noise: the
// reference's per-draw sequence (mul, add, softmax, +=: ~13 tensor passes per draw) becomes ONE read of eps_g.  The sums
// start from the values already in acc_p / acc_h and grow draw by draw in the reference's order (same association).
// ------------------------------------------------------------------------------------------------
template <int NI>
__global__ void __launch_bounds__(ROW_BLOCK)
k_mc_softmax(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ eps, int64_t N, int64_t C,
             int G, float* __restrict__ acc_p, float* __restrict__ acc_h) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* m = mean + row * C;
  const float* v = var + row * C;
  float mu[NI], sd[NI], ap[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int64_t j = lane + 32 * i;
    const bool ok = j < C;
    mu[i] = ok ? m[j] : -INFINITY;  // padded classes: exp(-inf) = 0, never stored
    sd[i] = ok ? sqrtf(v[j]) : 0.f;
    ap[i] = (ok && acc_p != nullptr) ? acc_p[row * C + j] : 0.f;
  }
  float ah = acc_h != nullptr ? acc_h[row] : 0.f;
  for (int g = 0; g < G; ++g) {
    const float* e = eps + (static_cast<int64_t>(g) * N + row) * C;
    float z[NI];
    float zmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int64_t j = lane + 32 * i;
      const float ev = j < C ? __ldcs(e + j) : 0.f;
      z[i] = __fadd_rn(mu[i], __fmul_rn(ev, sd[i]));  // torch: mean + (randn * std), two roundings
      zmax = fmaxf(zmax, z[i]);
    }
    zmax = warp_max(zmax);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      z[i] = expf(z[i] - zmax);
      s += z[i];
    }
    s = warp_sum(s);
    float h = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const float p = z[i] / s;
      ap[i] += p;
      if (acc_h != nullptr && lane + 32 * i < C) h += p * logf(p);  // (0 * log 0 = NaN, as in the reference)
    }
    if (acc_h != nullptr) ah -= warp_sum(h);
  }
  if (acc_p != nullptr) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int64_t j = lane + 32 * i;
      if (j < C) acc_p[row * C + j] = ap[i];
    }
  }
  if (acc_h != nullptr && lane == 0) acc_h[row] = ah;
}

__global__ void k_symmetrize_scale(float* __restrict__ A, int64_t d, int64_t ld, float scale) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j > i || i >= d) return;
  const float v = A[i * ld + j] * scale;
  A[i * ld + j] = v;
  if (i != j) A[j * ld + i] = v;
}

}  // namespace

// ================================================================================================
int launch_rows_to_16(const float* in, int64_t R, int64_t d, int64_t ld, int append_one, int fmt, int row_pow2_scale,
                      float gmult, void* out, int64_t k_pad, float* row_unscale, cudaStream_t st) {
  if (R <= 0) return BVLM_OK;
  if (k_pad < d + (append_one ? 1 : 0) || (k_pad & 1)) return BVLM_EINVAL;
  k_rows_to_16<<<row_grid(R), ROW_BLOCK, 0, st>>>(in, R, d, ld, append_one, fmt, row_pow2_scale, gmult,
                                                  static_cast<uint16_t*>(out), k_pad, row_unscale);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_predictive_row_prep(const float* x, int64_t R, int64_t D, int64_t ld, const float* quad,
                               const float* diag_other, float sum_diag_self, float kappa, float s2, int side,
                               int nsplit, float opscale, __half* packed, int64_t seg_pad, int64_t out_pitch,
                               uint8_t* packed8, int64_t seg8, float* out0, float* out1, cudaStream_t st) {
  if (R <= 0) return BVLM_OK;
  if ((nsplit != 1 && nsplit != 2 && nsplit != 3) || seg_pad < D || (seg_pad & 1)) return BVLM_EINVAL;
  if (nsplit == 2 && (packed8 == nullptr || seg8 < D || (seg8 & 1))) return BVLM_EINVAL;
  k_predictive_row_prep<<<row_grid(R), ROW_BLOCK, 0, st>>>(x, R, D, ld, quad, diag_other, sum_diag_self, kappa, s2, side,
                                                           nsplit, opscale, packed, seg_pad, out_pitch, packed8, seg8, out0,
                                                           out1);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_predictive_embed_prep(const float* x, int64_t R, int64_t D, int64_t ld, const float* diag_other, int nsplit,
                                 __half* packed, int64_t seg_pad, int64_t out_pitch, uint8_t* packed8, int64_t seg8, float* n2,
                                 float* pd, float* unscale, const float* act, int64_t d_act, int64_t ld_act, int append_one,
                                 __half* act16, int64_t act_kpad, float* act_unscale, cudaStream_t st) {
  if (R <= 0) return BVLM_OK;
  if (act != nullptr && (act16 == nullptr || act_unscale == nullptr || act_kpad < d_act + (append_one ? 1 : 0) || (act_kpad & 1)))
    return BVLM_EINVAL;
  if ((nsplit != 1 && nsplit != 2 && nsplit != 3) || seg_pad < D || (seg_pad & 1)) return BVLM_EINVAL;
  if (nsplit == 2 && (packed8 == nullptr || seg8 < D || (seg8 & 1))) return BVLM_EINVAL;
  // fast path: 16-byte aligned rows that fit the register-resident kernel (covers every model of the reference)
  const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int64_t ecols = seg_pad > seg8 && nsplit == 2 ? seg_pad : (nsplit == 2 ? seg8 : seg_pad);
  const bool vec_ok = act != nullptr && (x == nullptr || al16(x)) && al16(act) && al16(diag_other) && al16(packed) && al16(act16) &&
                      (ld % 4) == 0 && (ld_act % 4) == 0 && (seg_pad % 4) == 0 && (act_kpad % 4) == 0 &&
                      ((out_pitch > 0 ? out_pitch : seg_pad) % 4) == 0 && (nsplit != 2 || ((seg8 % 4) == 0 && al16(packed8))) &&
                      ecols <= 1024 && act_kpad <= 3200;
  if (vec_ok) {
#define BVLM_PREP_VEC(EVV, AVV)                                                                                            \
    k_predictive_prep_vec<EVV, AVV><<<row_grid(R), ROW_BLOCK, 0, st>>>(x, R, D, ld, diag_other, nsplit, packed, seg_pad,      \
                                                                       out_pitch, packed8, seg8, n2, pd, unscale, act, d_act, \
                                                                       ld_act, append_one, act16, act_kpad, act_unscale)
    if (act_kpad <= 1024) BVLM_PREP_VEC(8, 8);
    else if (act_kpad <= 1536) BVLM_PREP_VEC(8, 12);
    else BVLM_PREP_VEC(8, 25);
#undef BVLM_PREP_VEC
  } else {
    k_predictive_embed_prep<<<row_grid(R), ROW_BLOCK, 0, st>>>(x, R, D, ld, diag_other, nsplit, packed, seg_pad, out_pitch,
                                                               packed8, seg8, n2, pd, unscale, act, d_act, ld_act, append_one,
                                                               act16, act_kpad, act_unscale);
  }
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_ggn_row_prep(const float* x, int64_t R, int64_t D, int64_t ld, float opscale, int nsplit, int side, __half* xhat,
                        int64_t d_pad, float* inv_norm, float* w_raw, float* w_sum, cudaStream_t st) {
  if (R <= 0) return BVLM_OK;
  if (nsplit != 1 && nsplit != 3) return BVLM_EINVAL;
  k_ggn_row_prep<<<row_grid(R), ROW_BLOCK, 0, st>>>(x, R, D, ld, opscale, nsplit, side, xhat, d_pad, inv_norm, w_raw, w_sum);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_normalize_weights(const float* w_raw, const float* w_sum, int64_t R, float* w, cudaStream_t st) {
  if (R <= 0) return BVLM_OK;
  k_normalize_weights<<<static_cast<unsigned>((R + 255) / 256), 256, 0, st>>>(w_raw, w_sum, R, w);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_ggn_row_finalize(const float* x, int64_t B, int64_t D, int64_t Dp, int64_t ldx, const float* inv_norm, const float* w,
                            const float* y, int64_t ldy, const float* inv_norm_y, const int* pivot, const float* rest,
                            const float* inv_gamma, const float* Nraw, const float* Rraw, int64_t ldm, float unscale_n,
                            float unscale_r, int siglip, float g, __half* LA, __half* RA, __half* LB, __half* RB, int64_t ldl,
                            int64_t ldr, cudaStream_t st) {
  if (B <= 0) return BVLM_OK;
  if ((Dp & 1) || (ldl & 1) || (ldr & 1)) return BVLM_EINVAL;
  const auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const bool vec = (D % 4) == 0 && (Dp % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0 && (ldm % 4) == 0 && (ldl % 4) == 0 &&
                   (ldr % 4) == 0 && al(x, 16) && al(y, 16) && al(Nraw, 16) && al(Rraw, 16) && al(LB, 8) && al(RB, 8) &&
                   (siglip || (al(LA, 8) && al(RA, 8)));
  if (vec) {
    k_ggn_row_finalize_vec<<<row_grid(B), ROW_BLOCK, 0, st>>>(x, B, D, Dp, ldx, inv_norm, w, y, ldy, inv_norm_y, pivot, rest,
                                                              inv_gamma, Nraw, Rraw, ldm, unscale_n, unscale_r, siglip, g, LA,
                                                              RA, LB, RB, ldl, ldr);
  } else {
    k_ggn_row_finalize<<<row_grid(B), ROW_BLOCK, 0, st>>>(x, B, D, Dp, ldx, inv_norm, w, y, ldy, inv_norm_y, pivot, rest,
                                                          inv_gamma, Nraw, Rraw, ldm, unscale_n, unscale_r, siglip, g, LA, RA,
                                                          LB, RB, ldl, ldr);
  }
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_ggn_scale_targets(const float* y, int64_t C, int64_t D, int64_t Dp, int64_t ldy, const float* inv_norm_y,
                             const float* q, const float* inv_gamma, float mult, __half* out, int64_t ldo, cudaStream_t st) {
  if (C <= 0) return BVLM_OK;
  k_ggn_scale_targets<<<row_grid(C), ROW_BLOCK, 0, st>>>(y, C, D, Dp, ldy, inv_norm_y, q, inv_gamma, mult, out, ldo);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_merge_rowstats(float* m, float* rest, int* piv, int64_t B, int S, cudaStream_t st) {
  if (B <= 0) return BVLM_OK;
  k_merge_rowstats<<<static_cast<unsigned>((B + 255) / 256), 256, 0, st>>>(m, rest, piv, B, S);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_ggn_rowinfo(const float* m2, const float* rest, const int* piv, const float* w, int64_t B, int siglip, float4* out,
                       cudaStream_t st) {
  if (B <= 0) return BVLM_OK;
  k_ggn_rowinfo<<<static_cast<unsigned>((B + 255) / 256), 256, 0, st>>>(m2, rest, piv, w, B, siglip, out);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_ggn_scalars(const float* q, int64_t C, float* scalars, float inv_count, cudaStream_t st) {
  unsigned grid = static_cast<unsigned>((C + 255) / 256);
  if (grid > 1024) grid = 1024;
  k_max_nonneg<<<grid, 256, 0, st>>>(q, C, reinterpret_cast<unsigned int*>(scalars + 2));
  count_launch();
  k_ggn_scalars<<<1, 32, 0, st>>>(scalars, inv_count);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_sym_add(const float* S, int64_t d, int64_t lds, float* out, int64_t ldo, float alpha, const float* alpha_dev,
                   int accumulate, cudaStream_t st) {
  if (d <= 0) return BVLM_OK;
  dim3 grid(static_cast<unsigned>((d + 255) / 256), static_cast<unsigned>(d));
  k_sym_add<<<grid, 256, 0, st>>>(S, d, lds, out, ldo, alpha, alpha_dev, accumulate);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_scale_cols_f16(const float* x, int64_t n, int64_t d, int64_t ld, const float* scale, int append_one, __half* out,
                          int64_t ldo, cudaStream_t st) {
  if (n <= 0) return BVLM_OK;
  if (ldo < d + (append_one ? 1 : 0) || (ldo & 1)) return BVLM_EINVAL;
  const int vec_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(scale) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 7) == 0 && (ld % 4) == 0 && (ldo % 4) == 0;
  k_scale_cols_f16<<<row_grid(n), ROW_BLOCK, 0, st>>>(x, n, d, ld, scale, append_one, out, ldo, vec_ok);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_col_pow2_scale(const float* x, int64_t n, int64_t d, int64_t ld, int append_one, unsigned int* amax_bits,
                          float* scale, float* unscale, cudaStream_t st) {
  if (n <= 0 || d <= 0) return BVLM_OK;
  BVLM_CUDA_TRY(cudaMemsetAsync(amax_bits, 0, static_cast<size_t>(d) * sizeof(unsigned int), st));
  const int rows_per_block = 256;
  dim3 grid(static_cast<unsigned>((d + 31) / 32), static_cast<unsigned>((n + rows_per_block - 1) / rows_per_block));
  k_col_absmax<<<grid, dim3(32, 8), 0, st>>>(x, n, d, ld, rows_per_block, amax_bits);
  count_launch();
  const int64_t dA = d + (append_one ? 1 : 0);
  k_col_pow2_scale<<<static_cast<unsigned>((dA + 255) / 256), 256, 0, st>>>(amax_bits, d, append_one, scale, unscale);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_mc_softmax(const float* mean, const float* var, const float* eps, int64_t N, int64_t C, int G, float* acc_p,
                      float* acc_h, cudaStream_t st) {
  if (N <= 0 || C <= 0 || G <= 0) return BVLM_OK;
  if (C > 1024) return BVLM_ENOTSUP;  // the row must fit the warp's registers
  const unsigned grid = row_grid(N);
  timing_begin(TAG_PROBIT, st);
  if (C <= 128) k_mc_softmax<4><<<grid, ROW_BLOCK, 0, st>>>(mean, var, eps, N, C, G, acc_p, acc_h);
  else if (C <= 256) k_mc_softmax<8><<<grid, ROW_BLOCK, 0, st>>>(mean, var, eps, N, C, G, acc_p, acc_h);
  else if (C <= 512) k_mc_softmax<16><<<grid, ROW_BLOCK, 0, st>>>(mean, var, eps, N, C, G, acc_p, acc_h);
  else k_mc_softmax<32><<<grid, ROW_BLOCK, 0, st>>>(mean, var, eps, N, C, G, acc_p, acc_h);
  timing_end(TAG_PROBIT, st);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_probit_softmax(const float* mean, const float* var, int64_t N, int64_t C, int64_t ld, float* probs,
                          cudaStream_t st) {
  if (N <= 0 || C <= 0) return BVLM_OK;
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec4 = C % 4 == 0 && ld % 4 == 0 && al16(mean) && al16(var) && al16(probs);
  timing_begin(TAG_PROBIT, st);
  if (vec4) k_probit_softmax<true><<<row_grid(N), ROW_BLOCK, 0, st>>>(mean, var, N, C, ld, probs);
  else k_probit_softmax<false><<<row_grid(N), ROW_BLOCK, 0, st>>>(mean, var, N, C, ld, probs);
  timing_end(TAG_PROBIT, st);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

int launch_symmetrize_scale(float* A, int64_t d, int64_t ld, float scale, cudaStream_t st) {
  if (d <= 0) return BVLM_OK;
  dim3 grid(static_cast<unsigned>((d + 255) / 256), static_cast<unsigned>(d));
  k_symmetrize_scale<<<grid, 256, 0, st>>>(A, d, ld, scale);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

}  // namespace bvlm
