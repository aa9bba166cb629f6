// Device-side row conversion of the predictive's source embeddings, shared by the stand-alone prep kernel and by the
// quadratic-form GEMM, whose (otherwise idle) epilogue warps convert the embedding rows of their row panel while the
// tensor cores work on the activations: HBM-bound and tensor-bound work overlap inside one kernel.
//
// Per row r (16-byte aligned, up to 128*EV floats), with an exact power-of-two row scale 2^k (2^5 more in the fp16+fp8 mode):
//   packed  [r, :]  fp16( x 2^k )                    (+ fp16 lo at column seg_pad when nsplit == 3)
//   packed8 [r, :]  E4M3 [ 32 (v - fp16 v) | v/32 ]  (nsplit == 2)
//   n2 = |x|^2, pd = sum_d x_d^2 diag_other_d, unscale = 2^-k
#pragma once
#include <cuda_fp8.h>

#include "common.cuh"

namespace bvlm {

struct EmbedPrepArgs {
  const float* x;           // [R, ld] fp32; nullptr = nothing to do
  int64_t R, D, ld;
  const float* diag_other;  // [D]
  int nsplit;               // 1 | 2 (fp16 + fp8) | 3 (fp16 hi/lo)
  __half* packed;
  int64_t seg_pad, pitch;
  uint8_t* packed8;
  int64_t seg8;
  float* n2;
  float* pd;
  float* unscale;
};

template <int EV>
__device__ __forceinline__ void embed_row_load(const EmbedPrepArgs& a, int64_t row, int lane, float4 (&e)[EV]) {
  const float* xr = a.x + row * a.ld;
#pragma unroll
  for (int i = 0; i < EV; ++i) {
    const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * 4;
    e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < a.R) {
      if (c + 3 < a.D) e[i] = __ldg(reinterpret_cast<const float4*>(xr + c));
      else if (c < a.D) {
        e[i].x = xr[c];
        if (c + 1 < a.D) e[i].y = xr[c + 1];
        if (c + 2 < a.D) e[i].z = xr[c + 2];
      }
    }
  }
}

// EXACT: D == seg_pad (== seg8 in the fp16 + fp8 mode) == 128 * EV -- no bounds checks, 32-bit offsets (the epilogue-warp
// caller has only two warps per scheduler to hide latency with, so every instruction counts there).
struct EmbedRowStats {
  float n2, pd;
  int ee;  // row scale exponent
};

// pass 1: |x|^2, sum_d x_d^2 diag_other_d, row scale (ends in warp reductions: every lane's e[] has been consumed on return)
// `dl_cached` (optional): the lane's float4 slices of diag_other kept in registers by the caller -- in the GEMM epilogue a
// per-row reload misses L1 (bulk copies and stores stream through it) and stalls the FMAs on the long scoreboard.
template <int EV, bool EXACT = false>
__device__ __forceinline__ EmbedRowStats embed_row_stats(const EmbedPrepArgs& a, int lane, const float4 (&e)[EV],
                                                         const float4* dl_cached = nullptr) {
  // packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2): half the issue slots of the scalar form
  float2 n2v = make_float2(0.f, 0.f), pdv = make_float2(0.f, 0.f);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < EV; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dl_cached != nullptr) dl = dl_cached[i];
    else if (EXACT || c + 3 < a.D) dl = __ldg(reinterpret_cast<const float4*>(a.diag_other + c));
    else if (c < a.D) {
      dl.x = a.diag_other[c];
      if (c + 1 < a.D) dl.y = a.diag_other[c + 1];
      if (c + 2 < a.D) dl.z = a.diag_other[c + 2];
    }
    const float2 e0 = make_float2(e[i].x, e[i].y), e1 = make_float2(e[i].z, e[i].w);
    const float2 q0 = __fmul2_rn(e0, e0), q1 = __fmul2_rn(e1, e1);
    n2v = __fadd2_rn(n2v, __fadd2_rn(q0, q1));
    pdv = __ffma2_rn(q0, make_float2(dl.x, dl.y), __ffma2_rn(q1, make_float2(dl.z, dl.w), pdv));
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(e[i].x), fabsf(e[i].y)), fmaxf(fabsf(e[i].z), fabsf(e[i].w))));
  }
  const float n2 = warp_sum(n2v.x + n2v.y);
  const float pd = warp_sum(pdv.x + pdv.y);
  amax = warp_max(amax);
  int ee = 0;
  if (amax > 0.f && isfinite(amax)) {
    ee = (a.nsplit == 2 ? 8 : 9) - (static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 126);  // frexp exponent
    if ((__float_as_uint(amax) >> 23) == 0u) ee = 60;                                                // subnormal
    ee = ee < -60 ? -60 : (ee > 60 ? 60 : ee);
  }
  return EmbedRowStats{n2, pd, ee};
}

// pass 2: scaled fp16 (+ fp16 lo | + fp8 compensation terms) row and its statistics
template <int EV, bool EXACT = false>
__device__ __forceinline__ void embed_row_store(const EmbedPrepArgs& a, int64_t row, int lane, const float4 (&e)[EV],
                                                const EmbedRowStats& rs) {
  const int ee = rs.ee;
  // fp16 + fp8 mode: the fp16 operand carries an extra 2^5 so that hi.hi, lo8.t8 and e8.tlo8 share the scale 2^10
  const float sc = __uint_as_float(static_cast<uint32_t>(ee + (a.nsplit == 2 ? 5 : 0) + 127) << 23);  // exact power of two
  const float2 sc2 = make_float2(sc, sc), m1 = make_float2(-1.f, -1.f), k32 = make_float2(32.f, 32.f),
               r32 = make_float2(0.03125f, 0.03125f);
  __half* o = a.packed + row * a.pitch;
  uint8_t* o8 = a.nsplit == 2 ? a.packed8 + row * 2 * a.seg8 : nullptr;
  const int seg_pad = static_cast<int>(a.seg_pad), seg8 = static_cast<int>(a.seg8);
#pragma unroll
  for (int i = 0; i < EV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float2 v0 = __fmul2_rn(make_float2(e[i].x, e[i].y), sc2), v1 = __fmul2_rn(make_float2(e[i].z, e[i].w), sc2);
    const __half2 h0 = __float22half2_rn(v0), h1 = __float22half2_rn(v1);
    const float2 d0 = __ffma2_rn(__half22float2(h0), m1, v0), d1 = __ffma2_rn(__half22float2(h1), m1, v1);  // exact residuals
    if (EXACT || c < seg_pad) {
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(o + c) = pk;
      if (a.nsplit == 3) {
        const __half2 l0 = __float22half2_rn(d0), l1 = __float22half2_rn(d1);
        pk.x = *reinterpret_cast<const uint32_t*>(&l0);
        pk.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(o + seg_pad + c) = pk;
      }
    }
    if (a.nsplit == 2 && (EXACT || c < seg8)) {
      const uint32_t lo8 =
          static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(__fmul2_rn(d0, k32), __NV_SATFINITE, __NV_E4M3)) |
          (static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(__fmul2_rn(d1, k32), __NV_SATFINITE, __NV_E4M3)) << 16);
      const uint32_t x8 =
          static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(__fmul2_rn(v0, r32), __NV_SATFINITE, __NV_E4M3)) |
          (static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(__fmul2_rn(v1, r32), __NV_SATFINITE, __NV_E4M3)) << 16);
      *reinterpret_cast<uint32_t*>(o8 + c) = lo8;
      *reinterpret_cast<uint32_t*>(o8 + seg8 + c) = x8;
    }
  }
  if (lane == 0) {
    a.n2[row] = rs.n2;
    a.pd[row] = rs.pd;
    a.unscale[row] = __uint_as_float(static_cast<uint32_t>(127 - ee) << 23);
  }
}

template <int EV, bool EXACT = false>
__device__ __forceinline__ void embed_row_finish(const EmbedPrepArgs& a, int64_t row, int lane, const float4 (&e)[EV]) {
  if (row >= a.R) return;  // warp-uniform
  const EmbedRowStats rs = embed_row_stats<EV, EXACT>(a, lane, e);
  embed_row_store<EV, EXACT>(a, row, lane, e, rs);
}

}  // namespace bvlm
