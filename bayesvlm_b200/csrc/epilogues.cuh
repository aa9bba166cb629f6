// Epilogue functors for gemm_engine.cuh. Each thread of the four epilogue warps owns ONE accumulator row of the
// 128-row tile (row = 128*m + 32*ew + lane) and receives it 32 columns at a time in registers.
#pragma once
#include "gemm_engine.cuh"
#include "rowprep_device.cuh"

namespace bvlm {

__device__ __forceinline__ int epi_row(const EpiCtx& ctx, const TileCoord& tc) {
  return tc.row0 + ctx.ew * 32 + ctx.lane;
}

// Store 32 consecutive fp32 of one row; vectorised when the slice is complete and 16-byte aligned.
__device__ __forceinline__ void store_row32_f32(float* dst, const float (&v)[32], int n_valid, bool aligned16) {
  if (n_valid >= 32 && aligned16) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) __stcs(d4 + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_valid) __stcs(dst + j, v[j]);
  }
}
__device__ __forceinline__ void store_row32_f16(__half* dst, const float (&v)[32], int n_valid, bool aligned16) {
  if (n_valid >= 32 && aligned16) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 h0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]);
      __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
      __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
      __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2);
      pk.w = *reinterpret_cast<uint32_t*>(&h3);
      d4[j] = pk;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < n_valid) dst[j] = __float2half_rn(v[j]);
  }
}

// After the call, v[0] of lane j holds sum over the 32 lanes of their v[j] (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---------------------------------------------------------------------------------------------
// Plain fp32 output:  C = alpha * acc   (store)   or   C += alpha * acc   (red.global.add; split-K, SYRK)
// mirror=1 additionally accumulates the transposed element for off-diagonal tiles (lower-triangle SYRK).
// ---------------------------------------------------------------------------------------------
template <int BN>
struct EpiStoreF32 {
  // two 4 KB staging slabs per epilogue warp for the bulk-reduction path (use_tma)
  static constexpr size_t scratch_bytes(int warps) { return static_cast<size_t>(warps) * 2 * SLAB_BYTES; }
  struct Params {
    float* C;
    int64_t ldc;
    float alpha;
    int atomic;              // 1: accumulate into C (split-K / accumulate), 0: plain store
    int lower_only;          // 1: only elements with row >= col are written (SYRK; mirrored by a later pass)
    const float* alpha_dev;  // optional device scalar multiplied into alpha
    const float* unscale;    // optional per-index multiplier u: C[i,j] (+)= alpha * u[i] * u[j] * acc (SYRK operand scaling)
    // atomic accumulation as TMA bulk reductions (cp.reduce.async.bulk.tensor ... .add): 32 x 32 fp32 boxes staged in shared
    // memory, the adds done by the L2 on whole 128-byte lines.  The per-element red.global.add of the first version made a
    // warp instruction touch 32 rows (32 separate L2 atomics): the split-K SYRK of 65536 rows spent 2/3 of its time there.
    int use_tma;             // 1: tm_c is valid (C 16-byte aligned, ldc * 4 a multiple of 16) and atomic == 1
    CUtensorMap tm_c;        // [rows, cols] fp32, box {32 cols, 32 rows}, SWIZZLE_128B
  };
  struct State {
    float alpha;
    int sidx;
  };
  static constexpr bool ALL_CHUNKS = false;
  static constexpr bool UNROLL_CHUNKS = true;
  static constexpr bool DRAIN_FIRST = false;
  __device__ static void kernel_begin(State& st, const Params&, const EpiCtx&) { st.sidx = 0; }
  __device__ static void kernel_end(State&, const Params& p, const EpiCtx& ctx) {
    if (p.use_tma && ctx.lane == 0) tma_store_wait_all<0>();
  }
  __device__ static void item_begin(State& st, const Params& p, const EpiCtx&, const TileCoord&) {
    st.alpha = p.alpha_dev != nullptr ? p.alpha * (*p.alpha_dev) : p.alpha;
  }
  __device__ static void tile_begin(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void chunk(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc, float (&v)[32], int c) {
    const int row = epi_row(ctx, tc);
    const int col0 = tc.n * BN + c * 32;
    int n_valid = ctx.N - col0;
    if (p.use_tma) {
      // every lane takes part (warp-collective staging); rows / columns beyond the matrix are clipped by the tensor map,
      // masked elements (upper triangle of a diagonal SYRK tile) add an exact zero
      if (p.lower_only && col0 > tc.row0 + ctx.ew * 32 + 31) return;  // the whole 32 x 32 block lies above the diagonal (warp-uniform)
      int lim = 32;
      if (p.lower_only) {
        lim = row - col0 + 1;
        lim = lim < 0 ? 0 : lim;
      }
      float a = st.alpha;
      if (p.unscale != nullptr) {
        a *= row < ctx.M ? p.unscale[row] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= (j < lim && j < n_valid) ? a * p.unscale[col0 + j] : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= j < lim ? a : 0.f;
      }
      const uint32_t slab = ctx.scratch_u32 + static_cast<uint32_t>(ctx.wid * 2 + st.sidx) * SLAB_BYTES;
      slab_wait_free<1>(ctx.lane);  // the reduction issued two chunks ago has finished reading its slab
      slab_write_f32(slab, ctx.lane, v);
      slab_issue_add(&p.tm_c, slab, ctx.lane, col0, tc.row0 + ctx.ew * 32);
      slab_commit(ctx.lane);
      st.sidx ^= 1;
      return;
    }
    if (row >= ctx.M) return;
    if (p.lower_only) {
      const int lim = row - col0 + 1;  // columns col0 .. row
      n_valid = n_valid < lim ? n_valid : lim;
      if (n_valid <= 0) return;
    }
    float a = st.alpha;
    if (p.unscale != nullptr) {
      a *= p.unscale[row];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= a * ((j < n_valid) ? p.unscale[col0 + j] : 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= a;
    }
    float* dst = p.C + static_cast<int64_t>(row) * p.ldc + col0;
    if (!p.atomic) {
      store_row32_f32(dst, v, n_valid, (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_valid) red_add_f32(dst + j, v[j]);
    }
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
};

// ---------------------------------------------------------------------------------------------
// Row sum of squares over all N tiles of a row panel:  out[row] = scale * rowscale[row] * sum_n acc[row,n]^2
// (quadratic forms a^T A^-1 a evaluated as |G a|^2 with A^-1 = G^T G;  vlm.py:662-663)
// ---------------------------------------------------------------------------------------------
template <int BN>
struct EpiRowSumSq {
  static constexpr size_t scratch_bytes(int warps) { return warps == 8 ? 128 * sizeof(float) : 16; }
  struct Params {
    float* out;
    const float* row_scale;  // optional per-row multiplier (undoes the per-row power-of-two operand scaling)
    float scale;
  };
  struct State {
    float acc;
  };
  static constexpr bool ALL_CHUNKS = false;
  static constexpr bool UNROLL_CHUNKS = true;
  static constexpr bool DRAIN_FIRST = false;
  __device__ static void kernel_begin(State&, const Params&, const EpiCtx&) {}
  __device__ static void kernel_end(State&, const Params&, const EpiCtx&) {}
  __device__ static void item_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    st.acc = 0.f;
  }
  __device__ static void tile_begin(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void chunk(State& st, const Params&, const EpiCtx&, const TileCoord&, float (&v)[32], int) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      s0 = fmaf(v[j], v[j], s0);
      s1 = fmaf(v[j + 1], v[j + 1], s1);
    }
    st.acc += s0 + s1;
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int row = epi_row(ctx, tc);
    if (ctx.n_warps == 8) {  // two warps share a row (one per column half): merge through shared memory
      const uint32_t slot = ctx.scratch_u32 + 4u * static_cast<uint32_t>(ctx.ew * 32 + ctx.lane);
      if (ctx.wid >= 4) sts_f32(slot, st.acc);
      epi_bar_sync(ctx);
      if (ctx.wid < 4) st.acc += lds_f32(slot);
      epi_bar_sync(ctx);
      if (ctx.wid >= 4) return;
    }
    if (row < ctx.M) {
      float r = st.acc * p.scale;
      if (p.row_scale != nullptr) r *= p.row_scale[row];
      p.out[row] = r;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// EpiRowSumSq + a side job for its (nearly idle) epilogue warps: convert the EMBEDDING rows of the CTA's row panel (fp32 ->
// fp16 / fp8 operands of the mean GEMM + row statistics, rowprep_device.cuh) while the tensor cores work on the activations,
// so that the HBM-bound conversion overlaps the tensor-bound quadratic forms inside one kernel.
// Each warp owns rows row0 + wid + n_warps * j and a private ring of shared-memory row slots fed by bulk copies
// (cp.async.bulk + mbarrier): PREP_RING_BYTES per warp are in flight at all times, whatever the warp itself is doing --
// ordinary loads (two rows per warp in flight) left the memory system half idle.
constexpr int PREP_RING_BYTES = 12288;
constexpr int PREP_MAX_SLOTS = 8;

// EVX > 0: the embedding rows are exactly 128 * EVX floats wide and stored unpadded (no bounds checks); 0: generic, <= 1024.
template <int BN, int EVX = 0>
struct EpiQuadformPrep {
  using Base = EpiRowSumSq<BN>;
  static constexpr size_t scratch_bytes(int warps) { return 1024 + static_cast<size_t>(warps) * PREP_RING_BYTES; }
  struct Params {
    typename Base::Params base;
    EmbedPrepArgs prep;
    int row_bytes;  // D * 4 (multiple of 16)
    int slots;      // largest power of two <= min(PREP_MAX_SLOTS, PREP_RING_BYTES / row_bytes)
    int slot_shift; // log2(slots)
  };
  struct State {
    typename Base::State base;
    uint32_t issued, consumed;
    uint32_t ring;
    uint64_t* bars;
    float4 dl[EVX > 0 ? EVX : 1];  // this lane's slices of diag_other (exact-width instantiations)
  };
  static constexpr bool ALL_CHUNKS = false;
  static constexpr bool UNROLL_CHUNKS = true;
  static constexpr bool DRAIN_FIRST = false;

  __device__ static void issue(State& st, const Params& p, const EpiCtx& ctx, int64_t row) {
    const uint32_t slot = st.issued & static_cast<uint32_t>(p.slots - 1);
    if (ctx.lane == 0) {
      mbar_arrive_expect_tx(&st.bars[slot], static_cast<uint32_t>(p.row_bytes));
      bulk_load_1d(st.ring + slot * static_cast<uint32_t>(p.row_bytes), p.prep.x + row * p.prep.ld,
                   static_cast<uint32_t>(p.row_bytes), &st.bars[slot]);
    }
    ++st.issued;
  }
  __device__ static void kernel_begin(State& st, const Params& p, const EpiCtx& ctx) {
    st.issued = st.consumed = 0;
    st.bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ctx.scratch) + 512) + ctx.wid * PREP_MAX_SLOTS;
    st.ring = ctx.scratch_u32 + 1024u + static_cast<uint32_t>(ctx.wid) * PREP_RING_BYTES;
    if constexpr (EVX > 0) {
#pragma unroll
      for (int i = 0; i < EVX; ++i) st.dl[i] = __ldg(reinterpret_cast<const float4*>(p.prep.diag_other + (i * 32 + ctx.lane) * 4));
    }
    if (ctx.lane == 0) {
      for (int i = 0; i < p.slots; ++i) mbar_init(&st.bars[i], 1);
      fence_barrier_init();
      fence_proxy_async_smem();
    }
    __syncwarp();
  }
  __device__ static void kernel_end(State&, const Params&, const EpiCtx&) {}
  __device__ static void item_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    Base::item_begin(st.base, p.base, ctx, tc);
    const int per_warp = GEMM_BM / ctx.n_warps;
    for (int j = 0; j < p.slots && j < per_warp; ++j) {
      const int64_t row = tc.row0 + ctx.wid + ctx.n_warps * j;
      if (row >= p.prep.R) break;
      issue(st, p, ctx, row);
    }
  }
  __device__ static void tile_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    // column tile n of T converts the j range proportional to its share (n + 1) / (T (T + 1) / 2) of the triangular K
    // loop, BEFORE waiting for the tile's accumulator: the conversion is spread over the main loop of the row panel
    const int T = (ctx.N + BN - 1) / BN, per_warp = GEMM_BM / ctx.n_warps;
    const int j0 = per_warp * tc.n * (tc.n + 1) / (T * (T + 1)), j1 = per_warp * (tc.n + 1) * (tc.n + 2) / (T * (T + 1));
    for (int j = j0; j < j1; ++j) {
      const int64_t row = tc.row0 + ctx.wid + ctx.n_warps * j;
      if (row >= p.prep.R) break;
      const uint32_t slot = st.consumed & static_cast<uint32_t>(p.slots - 1);
      mbar_wait(&st.bars[slot], (st.consumed >> p.slot_shift) & 1u);
      ++st.consumed;
      const uint32_t src = st.ring + slot * static_cast<uint32_t>(p.row_bytes);
      constexpr int EV = EVX > 0 ? EVX : 8;
      float4 e[EV];
#pragma unroll
      for (int i = 0; i < EV; ++i) {
        const int c = (i * 32 + ctx.lane) * 4;
        e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EVX > 0 || c < p.prep.D) e[i] = lds_v4(src + static_cast<uint32_t>(c) * 4u);
      }
      // pass 1 ends in warp reductions over values derived from every lane's e[]: all reads of the slot have completed,
      // so it can be handed back to the copy engine (write-after-read needs no proxy fence, as in any TMA pipeline)
      const EmbedRowStats rs = embed_row_stats<EV, (EVX > 0)>(p.prep, ctx.lane, e, EVX > 0 ? st.dl : nullptr);
      const int jn = j + p.slots;
      const int64_t row_n = tc.row0 + ctx.wid + ctx.n_warps * jn;
      if (jn < per_warp && row_n < p.prep.R) issue(st, p, ctx, row_n);
      embed_row_store<EV, (EVX > 0)>(p.prep, row, ctx.lane, e, rs);
    }
  }
  __device__ static void chunk(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc, float (&v)[32], int c) {
    Base::chunk(st.base, p.base, ctx, tc, v, c);
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    Base::item_end(st.base, p.base, ctx, tc);
  }
};

// ---------------------------------------------------------------------------------------------
// Kronecker-Laplace predictive (vlm.py:630-684, collapsed):
//   mean_ij = mean_scale * acc_ij
//   var_ij  = u_i * a_j + v_i * b_j
// with u_i = s^2 p_i / E_i, v_i = s^2 alpha_i / E_i, a_j = gamma_j / E_j, b_j = (gamma_j kappa + q_j) / E_j.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct EpiPredictive {
  // per warp: mean + var slabs, double-buffered with 4 epilogue warps (16 KB per warp), single with 8 (8 KB per warp)
  static constexpr size_t scratch_bytes(int) { return 16 * SLAB_BYTES; }
  struct Params {
    CUtensorMap tm_mean, tm_var;  // [N, C] fp32, box {32 cols, 32 rows}, SWIZZLE_128B (used when use_tma)
    float* mean;
    float* var;
    int64_t ld;
    // per-row inputs (vlm.py:659-668): alpha_i = a_i^T A^-1 a_i, n2_i = |e_i|^2, pd_i = sum_d e_id^2 delta_d, esc_i = 2^-k_i;
    // E_i = n2_i + alpha_i sum_beta, u_i = s2 pd_i / E_i, v_i = s2 alpha_i / E_i, mean_ij = acc_ij mean_scale esc_i / sqrt(E_i)
    const float* alpha;
    const float* n2;
    const float* pd;
    const float* esc;
    float sum_beta;
    float ls;             // log-space logit scale (host copy), used when ls_dev == nullptr
    const float* ls_dev;  // optional DEVICE scalar holding the log-space logit scale: read per tile, so a parameter updated
                          // in place (the `.data` idiom of reference epig.py:230) is always current and the host never syncs
    float mean_unscale;   // 1 / (operand scaling of the two unit-energy embeddings)
    const float* a;   // padded to a multiple of BN entries (zeros beyond C)
    const float* b;
    int use_tma;      // 0: row pitch not a multiple of 16 bytes -> direct stores
  };
  struct State {
    float u, v, rm;
  };
  static constexpr bool ALL_CHUNKS = false;
  static constexpr bool UNROLL_CHUNKS = true;
  static constexpr bool DRAIN_FIRST = true;   // 8-warp configuration: free the TMEM buffer before the stores
  __device__ static void kernel_begin(State&, const Params&, const EpiCtx&) {}
  __device__ static void kernel_end(State&, const Params& p, const EpiCtx& ctx) {
    if (p.use_tma && ctx.lane == 0) tma_store_wait_all<0>();
  }
  __device__ static void item_begin(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void tile_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int row = epi_row(ctx, tc);
    st.u = st.v = st.rm = 0.f;
    if (row < ctx.M) {
      const float al = p.alpha[row];
      const float E = fmaf(al, p.sum_beta, p.n2[row]);
      const float rE = 1.0f / E;
      const float s = expf(p.ls_dev != nullptr ? __ldg(p.ls_dev) : p.ls);
      st.u = s * s * p.pd[row] * rE;
      st.v = s * s * al * rE;
      st.rm = s * p.mean_unscale * p.esc[row] * rsqrtf(E);
    }
  }
  __device__ static void chunk(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc, float (&v)[32], int c) {
    const int row = epi_row(ctx, tc);
    const int col0 = tc.n * BN + c * 32;
    const float4* a4p = reinterpret_cast<const float4*>(p.a + col0);  // same address in every lane: broadcast loads
    const float4* b4p = reinterpret_cast<const float4*>(p.b + col0);
    float var[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a4 = __ldg(a4p + j);
      const float4 b4 = __ldg(b4p + j);
      var[4 * j + 0] = fmaf(st.u, a4.x, st.v * b4.x);
      var[4 * j + 1] = fmaf(st.u, a4.y, st.v * b4.y);
      var[4 * j + 2] = fmaf(st.u, a4.z, st.v * b4.z);
      var[4 * j + 3] = fmaf(st.u, a4.w, st.v * b4.w);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= st.rm;
#ifdef BVLM_DIAG  // diagnostic build only (BVLM_DEBUG_EPI): which part of the output path costs the main loop its rate
    if (p.use_tma == 2) return;  // 1: epilogue math only, no shared-memory staging, no stores
    if (p.use_tma == 3) {        // 2: slabs written to shared memory, never stored
      const uint32_t slab_m = ctx.scratch_u32 + static_cast<uint32_t>(ctx.wid) * (4 * SLAB_BYTES) + static_cast<uint32_t>(c & 1) * (2 * SLAB_BYTES);
      slab_write_f32(slab_m, ctx.lane, v);
      slab_write_f32(slab_m + SLAB_BYTES, ctx.lane, var);
      return;
    }
    if (p.use_tma == 6) {        // 5: as 3, but every 32 x 32 box leaves as TWO 32 x 16 boxes (tm_mean / tm_var encoded with 16 rows): is the
                                 //    output path sensitive to the NUMBER of bulk stores or to their bytes?
      const uint32_t slab_m = ctx.scratch_u32 + static_cast<uint32_t>(ctx.wid) * (4 * SLAB_BYTES) + static_cast<uint32_t>(c & 1) * (2 * SLAB_BYTES);
      slab_wait_free<1>(ctx.lane);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        slab_issue(&p.tm_mean, slab_m + h * (SLAB_BYTES / 2), ctx.lane, col0, tc.row0 + ctx.ew * 32 + 16 * h);
        slab_issue(&p.tm_var, slab_m + SLAB_BYTES + h * (SLAB_BYTES / 2), ctx.lane, col0, tc.row0 + ctx.ew * 32 + 16 * h);
      }
      slab_commit(ctx.lane);
      return;
    }
    if (p.use_tma == 4) {        // 3: bulk stores of (stale) slabs, nothing written to shared memory
      const uint32_t slab_m = ctx.scratch_u32 + static_cast<uint32_t>(ctx.wid) * (4 * SLAB_BYTES) + static_cast<uint32_t>(c & 1) * (2 * SLAB_BYTES);
      slab_wait_free<1>(ctx.lane);
      slab_issue(&p.tm_mean, slab_m, ctx.lane, col0, tc.row0 + ctx.ew * 32);
      slab_issue(&p.tm_var, slab_m + SLAB_BYTES, ctx.lane, col0, tc.row0 + ctx.ew * 32);
      slab_commit(ctx.lane);
      return;
    }
#endif
    if (p.use_tma) {
      // rows beyond N and columns beyond C are clipped by the tensor map
      const bool dbl = ctx.n_warps == 4;
      const uint32_t slab_m = ctx.scratch_u32 + static_cast<uint32_t>(ctx.wid) * ((dbl ? 4 : 2) * SLAB_BYTES) +
                              static_cast<uint32_t>(dbl ? (c & 1) : 0) * (2 * SLAB_BYTES);
      const uint32_t slab_v = slab_m + SLAB_BYTES;
      const int row0 = tc.row0 + ctx.ew * 32;
      // one bulk group (mean + var) per chunk: double-buffered -> only the previous chunk's may still be reading
      if (dbl) slab_wait_free<1>(ctx.lane);
      else slab_wait_free<0>(ctx.lane);
      slab_write_f32(slab_m, ctx.lane, v);
      slab_write_f32(slab_v, ctx.lane, var);
      slab_issue(&p.tm_mean, slab_m, ctx.lane, col0, row0);
      slab_issue(&p.tm_var, slab_v, ctx.lane, col0, row0);
      slab_commit(ctx.lane);
      return;
    }
    if (row >= ctx.M) return;
    const int n_valid = ctx.N - col0;
    const bool al = (p.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mean) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.var) & 15) == 0;
    const int64_t off = static_cast<int64_t>(row) * p.ld + col0;
    store_row32_f32(p.mean + off, v, n_valid, al);
    store_row32_f32(p.var + off, var, n_valid, al);
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
};

// ---------------------------------------------------------------------------------------------
// InfoNCE pass 1 (hessians.py:24-27): per source row, over all targets (online across the N tiles of a row panel):
//   pivot = argmax_c l_bc,  m = max_c l_bc,  rest = sum_{c != pivot} 2^(l_bc - m)      with l = s*log2(e)*<xh_b, yh_c>
// so that softmax(pivot) = 1/(1+rest) and 1 - softmax(pivot) = rest/(1+rest) are known WITHOUT cancellation.
// (Peaked softmaxes make Yh^T diag(p) Yh - (Yh^T p)(Yh^T p)^T cancel catastrophically; the GGN pipeline therefore
//  centres every row on its pivot target -- see kfac.cu.)
// ---------------------------------------------------------------------------------------------
// merge two online-softmax partials (pivot excluded from `rest`); on ties the first (a) stays the pivot
__device__ __forceinline__ void merge_rowstats(float& m, float& rest, int& piv, float m2, float rest2, int piv2) {
  if (m2 > m) {
    rest = rest2 + (rest + 1.f) * fast_exp2(m - m2);
    m = m2;
    piv = piv2;
  } else if (m2 > -INFINITY) {
    rest = rest + (rest2 + 1.f) * fast_exp2(m2 - m);
  }
}

template <int BN>
struct EpiRowLse {
  static constexpr size_t scratch_bytes(int warps) { return warps > 4 ? (warps / 4 - 1) * 3 * 128 * sizeof(float) : 16; }
  struct Params {
    float* rowmax2;  // [B, n_split] m   (log2 units)
    float* rest;     // [B, n_split]
    int* pivot;      // [B, n_split]
    float s_log2e;   // exp(logit_scale) * log2(e) / (operand scaling)
    int n_split;     // row panels are cut into n_split column ranges; k_merge_rowstats combines the partials
  };
  struct State {
    float m, rest;
    int piv;
  };
  static constexpr bool ALL_CHUNKS = false;
  static constexpr bool UNROLL_CHUNKS = false;
  static constexpr bool DRAIN_FIRST = false;
  __device__ static void kernel_begin(State&, const Params&, const EpiCtx&) {}
  __device__ static void kernel_end(State&, const Params&, const EpiCtx&) {}
  __device__ static void item_begin(State& st, const Params&, const EpiCtx&, const TileCoord&) {
    st.m = -INFINITY;
    st.rest = 0.f;
    st.piv = 0;
  }
  __device__ static void tile_begin(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void chunk(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc, float (&v)[32], int c) {
    // st.m is kept in RAW accumulator units (the scale s_log2e > 0 is folded into the exp2 argument)
    const int col0 = tc.n * BN + c * 32;
    const int n_valid = ctx.N - col0;
    if (n_valid < 32) {  // ragged last chunk only
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < n_valid ? v[j] : -INFINITY;
    }
    float cm = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) cm = fmaxf(cm, v[j]);
    if (cm > st.m) {  // a new row maximum (rare after the first chunks): the old pivot joins the rest, the new one leaves
      st.rest = (st.rest + 1.f) * fast_exp2((st.m - cm) * p.s_log2e);  // 0 on the first chunk (st.m = -inf)
      st.m = cm;
      int ci = 31;
#pragma unroll
      for (int j = 30; j >= 0; --j) ci = (v[j] == cm) ? j : ci;  // first maximum stays the pivot on ties
      st.piv = col0 + ci;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (j == ci) ? -INFINITY : v[j];
    }
    const float off = -st.m * p.s_log2e;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      s0 += fast_exp2(fmaf(v[j], p.s_log2e, off));
      s1 += fast_exp2(fmaf(v[j + 1], p.s_log2e, off));
      s2 += fast_exp2(fmaf(v[j + 2], p.s_log2e, off));
      s3 += fast_exp2(fmaf(v[j + 3], p.s_log2e, off));
    }
    st.rest += (s0 + s1) + (s2 + s3);
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int row = epi_row(ctx, tc);
    st.m *= p.s_log2e;  // raw accumulator units -> log2 units
    const int groups = ctx.n_warps / 4;  // column groups per accumulator row (one warp each)
    if (groups > 1) {  // the warps of the upper column groups hand their partials to the first one
      const int g = ctx.wid >> 2;
      const uint32_t base = ctx.scratch_u32 + 12u * static_cast<uint32_t>(ctx.ew * 32 + ctx.lane);
      if (g > 0) {
        const uint32_t slot = base + static_cast<uint32_t>(g - 1) * (12u * 128u);
        sts_f32(slot, st.m);
        sts_f32(slot + 4, st.rest);
        sts_f32(slot + 8, __int_as_float(st.piv));
      }
      epi_bar_sync(ctx);
      if (g == 0) {
        for (int k = 0; k + 1 < groups; ++k) {
          const uint32_t slot = base + static_cast<uint32_t>(k) * (12u * 128u);
          merge_rowstats(st.m, st.rest, st.piv, lds_f32(slot), lds_f32(slot + 4), __float_as_int(lds_f32(slot + 8)));
        }
      }
      epi_bar_sync(ctx);
      if (g > 0) return;
    }
    if (row < ctx.M) {
      const int64_t o = static_cast<int64_t>(row) * p.n_split + tc.split;
      p.rowmax2[o] = st.m;
      p.rest[o] = st.rest;
      p.pivot[o] = st.piv;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// GGN pass 2. For every (source b, target c) the per-pair curvature weight
//     InfoNCE : omega = softmax_c(s L_b.) / (1 - p*_b) with the pivot target of the row masked to 0   (hessians.py:27)
//     SigLIP  : omega = sigma(z)(1 - sigma(z)), z = s L + bias   (hessians.py:88-94, without the s^2 factor)
// is written as fp16 (scaled by WSCALE so that 1/C-sized probabilities stay normal numbers) together with
//     InfoNCE : omega * d,  d = l_bc - m_b  (log2-unit logit distance to the pivot, <= 0)
//     SigLIP  : omega * L   (cosine)
// and the weighted column sums q_c = sum_b w_b omega_bc are reduced in registers across the M tiles of a column
// panel and written once (no atomics).
// ---------------------------------------------------------------------------------------------
constexpr float GGN_WSCALE = 4096.f;
// InfoNCE: |d| reaches 2 s log2(e) (~290 at s = 100) while the conditional weight of the runner-up target is ~1, so
// omega * d is stored with a smaller scale than omega itself to stay inside fp16 (max 65504).
constexpr float GGN_WDSCALE = 64.f;
constexpr int GGN_W_SLABS = 2;  // output slabs per epilogue warp of the weights pass (2: 64 KB, leaves room for a 4th operand stage)

template <int BN, bool SIGLIP>
struct EpiGgnWeights {
  // per epilogue warp NSLAB output slabs (4 KB each), then 4 x BN floats of per-quadrant column sums
  static constexpr int NSLAB = GGN_W_SLABS;
  static constexpr size_t scratch_bytes(int warps) { return warps * NSLAB * SLAB_BYTES + 4 * BN * sizeof(float); }
  struct Params {
    CUtensorMap tm_w, tm_wl;  // [B, Cp] fp16, box {64 cols, 32 rows}, SWIZZLE_128B
    const float4* rowinfo;  // [B] per source: {m2, lgw, wq, pivot bits} (InfoNCE) / {0, 0, wq, 0} (SigLIP); see k_ggn_rowinfo
    float* q;               // [N], accumulated with red.global.add (zeroed by the host)
    float s_log2e;         // InfoNCE: s*log2e/opscale ; SigLIP: s/opscale
    float l_scale;         // 1/opscale: acc -> cosine
    float bias;            // SigLIP logit bias
  };
  struct State {
    float m2, lgw, w;  // lgw = log2(rest) - log2(WSCALE): omega * WSCALE = 2^(d - lgw)
    int piv;
    int sidx;          // running bulk-store index (slab = sidx % 3)
    float4 nxt;        // row info of the NEXT tile, fetched one tile ahead (hides the global-load latency)
    int nxt_row;
  };
  static constexpr bool ALL_CHUNKS = true;      // slabs are issued per pair of 32-column chunks
  static constexpr bool UNROLL_CHUNKS = false;  // the body is large: keep one copy
  static constexpr bool DRAIN_FIRST = false;
  __device__ static uint32_t qsum_addr(const EpiCtx& ctx) {
    return ctx.scratch_u32 + static_cast<uint32_t>(ctx.n_warps) * (NSLAB * SLAB_BYTES);
  }
  __device__ static void kernel_begin(State& st, const Params&, const EpiCtx&) {
    st.sidx = 0;
    st.nxt_row = -1;
  }
  __device__ static void kernel_end(State&, const Params&, const EpiCtx& ctx) {
    if (ctx.lane == 0) tma_store_wait_all<0>();
  }
  __device__ static void item_begin(State&, const Params&, const EpiCtx& ctx, const TileCoord&) {
    // zero this warp's share of the column sums: quadrant ew, the columns of the chunks this warp owns
    const int per = (BN / 32) / (ctx.n_warps / 4);
    const uint32_t qs = qsum_addr(ctx) + 4u * static_cast<uint32_t>(ctx.ew * BN + (ctx.wid / 4) * per * 32 + ctx.lane);
    for (int i = 0; i < per; ++i) sts_f32(qs + 128u * i, 0.f);
  }
  __device__ static float4 load_info(const Params& p, const EpiCtx& ctx, int row) {
    // rows beyond B: lgw = +inf (omega = 0 without per-element masking), wq = 0
    return row < ctx.M ? __ldg(p.rowinfo + row) : make_float4(0.f, INFINITY, 0.f, __int_as_float(-1));
  }
  __device__ static void tile_begin(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const int row = epi_row(ctx, tc);
    const float4 ri = (st.nxt_row == row) ? st.nxt : load_info(p, ctx, row);
    st.m2 = ri.x;
    st.lgw = ri.y;
    st.w = ri.z;
    st.piv = __float_as_int(ri.w);
    if (tc.row0_next >= 0) {
      st.nxt_row = tc.row0_next + ctx.ew * 32 + ctx.lane;
      st.nxt = load_info(p, ctx, st.nxt_row);
    } else {
      st.nxt_row = -1;
    }
  }
  __device__ static void chunk(State& st, const Params& p, const EpiCtx& ctx, const TileCoord& tc, float (&v)[32], int c) {
    static_assert(GGN_WSCALE == 4096.f, "the exponent fold assumes WSCALE = 2^12");
    const int col0 = tc.n * BN + c * 32;
    float om[32];
    // Columns in [C, Cp) receive finite garbage here; the host zeroes that K padding after the pass. Columns >= Cp and
    // rows >= B are clipped by the tensor maps; q is only written for columns < C.
    if constexpr (SIGLIP) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float z = fmaf(v[j], p.s_log2e, p.bias);
        const float t = fast_exp2(-fabsf(z) * 1.4426950408889634f);
        const float d = 1.f + t;
        om[j] = __fdividef(t * GGN_WSCALE, d * d);
        v[j] *= p.l_scale;  // cosine
      }
    } else {
      const int pj = st.piv - col0;  // pivot position inside this chunk (outside [0,32) if elsewhere)
      const float dsc = GGN_WDSCALE / GGN_WSCALE;
      const float o_off = -(st.m2 + st.lgw), s_d = p.s_log2e * dsc, d_off = -st.m2 * dsc;
      const float2 se2 = make_float2(p.s_log2e, p.s_log2e), oo2 = make_float2(o_off, o_off), sd2 = make_float2(s_d, s_d),
                   do2 = make_float2(d_off, d_off);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {  // packed fp32x2 FMAs: half the issue slots
        const float2 vv = make_float2(v[j], v[j + 1]);
        const float2 ex = __ffma2_rn(vv, se2, oo2);  // d - lgw, d = l - m <= 0 (exactly 0 at the pivot)
        const float2 dd = __ffma2_rn(vv, sd2, do2);  // d * WDSCALE / WSCALE
        om[j] = fast_exp2(ex.x);
        om[j + 1] = fast_exp2(ex.y);
        v[j] = dd.x;
        v[j + 1] = dd.y;
      }
      if (pj >= 0 && pj < 32) {  // the pivot itself carries no conditional weight
#pragma unroll
        for (int j = 0; j < 32; ++j) om[j] = (j == pj) ? 0.f : om[j];
      }
    }
    // ---- stage fp16 omega / omega*(d|L) in the warp's rotating slabs (two 32-column chunks fill one 64-column slab)
    const uint32_t wbase = ctx.scratch_u32 + static_cast<uint32_t>(ctx.wid) * (NSLAB * SLAB_BYTES);
    const int h = c & 1;
    const int b0 = st.sidx % NSLAB;
    const int b1 = SIGLIP ? b0 : (st.sidx + 1) % NSLAB;
    // NSLAB = 3: every bulk store but the most recent one has finished reading; NSLAB = 2 (InfoNCE uses both per chunk pair):
    // all of them have -- they were issued two chunks (~2 us) ago
    if (h == 0) slab_wait_free<(NSLAB >= 3 ? 1 : 0)>(ctx.lane);
    if constexpr (!SIGLIP) slab_write_f16_half(wbase + static_cast<uint32_t>(b0) * SLAB_BYTES, ctx.lane, h, om);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 pr = __fmul2_rn(make_float2(v[j], v[j + 1]), make_float2(om[j], om[j + 1]));
      v[j] = pr.x;
      v[j + 1] = pr.y;
    }
    slab_write_f16_half(wbase + static_cast<uint32_t>(b1) * SLAB_BYTES, ctx.lane, h, v);
    if (h == 1) {
      const int row0 = tc.row0 + ctx.ew * 32;
      if constexpr (!SIGLIP) {
        slab_issue(&p.tm_w, wbase + static_cast<uint32_t>(b0) * SLAB_BYTES, ctx.lane, col0 - 32, row0);
        slab_commit(ctx.lane);
      }
      slab_issue(&p.tm_wl, wbase + static_cast<uint32_t>(b1) * SLAB_BYTES, ctx.lane, col0 - 32, row0);
      slab_commit(ctx.lane);
      st.sidx += SIGLIP ? 1 : 2;
      if (st.sidx >= NSLAB) st.sidx -= NSLAB;
    }
    // ---- weighted column sums of this 32 x 32 block: butterfly over the rows, accumulated in shared memory
    const float2 w2 = make_float2(st.w, st.w);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 pr = __fmul2_rn(make_float2(om[j], om[j + 1]), w2);
      om[j] = pr.x;
      om[j + 1] = pr.y;
    }
    const float qv = warp_transpose_reduce32(om, ctx.lane);
    const uint32_t qa = qsum_addr(ctx) + 4u * static_cast<uint32_t>(ctx.ew * BN + c * 32 + ctx.lane);
    sts_f32(qa, lds_f32(qa) + qv);
  }
  __device__ static void tile_end(State&, const Params&, const EpiCtx&, const TileCoord&) {}
  __device__ static void item_end(State&, const Params& p, const EpiCtx& ctx, const TileCoord& tc) {
    const uint32_t s = qsum_addr(ctx);
    epi_bar_sync(ctx);
    for (int i = ctx.wid * 32 + ctx.lane; i < BN; i += ctx.n_warps * 32) {
      const int col = tc.n * BN + i;
      if (col < ctx.N) {
        const float qs = lds_f32(s + 4u * i) + lds_f32(s + 4u * (BN + i)) + lds_f32(s + 4u * (2 * BN + i)) +
                         lds_f32(s + 4u * (3 * BN + i));
        red_add_f32(p.q + col, qs);
      }
    }
    epi_bar_sync(ctx);  // nobody zeroes the sums of the next item before they are read
  }
};

}  // namespace bvlm
