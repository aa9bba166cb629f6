// Persistent, warp-specialised tcgen05 GEMM engine for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T
//
//   * both operands are 16-bit (fp16 or bf16), K-major, row pitch a multiple of 16 bytes;
//   * TMA (128-byte swizzle) stages 128 x 64 / BN x 64 operand tiles into a STAGES-deep smem ring;
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM accumulator;
//   * four epilogue warps drain TMEM with tcgen05.ld (32 lanes x 32 columns at a time) and hand every
//     32-column slice to a pluggable epilogue functor -- that is where the Laplace math lives
//     (rank-2 variance, row sum-of-squares, online log-sum-exp, softmax/sigmoid weights, EPIG xlogy ...).
//
// The schedule is described by a GemmPlan: plain output tiles (optionally split along K, optionally only the
// lower triangle for SYRK), row panels (one M-tile, all N-tiles: row reductions keep state in registers) or
// column panels (one N-tile, all M-tiles: column reductions keep state in registers).
#pragma once
#include <atomic>

#include "common.cuh"
#include "tmap.cuh"

namespace bvlm {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 256;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-7 epilogue
constexpr int GEMM_EPI_WARP0 = 4;
constexpr int GEMM_AUX_BYTES = 256;

enum SchedMode : int { SCHED_TILES = 0, SCHED_ROW_PANEL = 1, SCHED_COL_PANEL = 2, SCHED_TRI_TILES = 3 };

struct GemmPlan {
  int M;         // valid output rows
  int N;         // valid output columns
  int kb_total;  // number of 64-wide K blocks
  int m_tiles;
  int n_tiles;
  int mode;      // SchedMode
  int splits;    // split-K factor (SCHED_TILES / SCHED_TRI_TILES); number of M (N) ranges a column (row) panel is cut into
  int tri_k;     // 1: operand B is lower triangular in (n,k) -> K loop stops at the diagonal block
  uint32_t idesc;
  int a_is_3d;       // 1: operand A is fetched through a 3-D map at (k, 0, m * a_outer_step)
  int a_outer_step;  // rows (2-D) or outer-dimension items (3-D) consumed per M tile
  uint32_t a_tx_bytes;
  uint32_t b_tx_bytes;
  // hi/lo split operands are stored once as [hi | lo] (2 segments of seg_kb K blocks); the three products
  // hi.hi + lo.hi + hi.lo walk 3 virtual segments whose source segment is a_seg[i] / b_seg[i].  seg_kb == 0: plain.
  // Bit v of a_seg_mask / b_seg_mask is the stored segment (0 = hi, 1 = lo) virtual segment v reads.
  int seg_kb;
  uint32_t a_seg_mask;
  uint32_t b_seg_mask;
  // CTA-pair engine only: K blocks >= kb_alt come from a second pair of tensor maps holding FP8 (E4M3) operands -- 128
  // elements per 128-byte K block -- and are multiplied with tcgen05.mma.kind::f8f6f4 (idesc_alt) into the SAME fp32
  // accumulator (error-compensation terms at twice the fp16 rate).  kb_alt >= kb_total: no alternate phase.
  int kb_alt;
  uint32_t idesc_alt;
};

struct TileCoord {
  int m, n, kb0, kb1;
  int split;  // panel split index (row / column panels)
  int row0;  // first accumulator row of this CTA's 128-row slab of the tile (set by the kernel, engine specific)
  int row0_next;  // row0 of the next tile of the same item when it is known to follow (column panels), else -1
};

template <int BN>
__host__ __device__ inline int plan_num_items(const GemmPlan& p) {
  switch (p.mode) {
    case SCHED_ROW_PANEL: return p.m_tiles * p.splits;
    case SCHED_COL_PANEL: return p.n_tiles * p.splits;
    case SCHED_TRI_TILES: return (p.m_tiles * (p.m_tiles + 1) / 2) * p.splits;
    default: return p.m_tiles * p.n_tiles * p.splits;
  }
}
// (32-bit unsigned arithmetic on purpose: these run per item in EVERY thread of the role loops, and a 64-bit division is
//  ~100 instructions; split * tiles < 2^31 for every schedule the library builds)
__host__ __device__ inline int col_panel_m0(const GemmPlan& p, int split) {
  return static_cast<int>(static_cast<uint32_t>(split) * static_cast<uint32_t>(p.m_tiles) / static_cast<uint32_t>(p.splits));
}
__host__ __device__ inline int row_panel_n0(const GemmPlan& p, int split) {
  return static_cast<int>(static_cast<uint32_t>(split) * static_cast<uint32_t>(p.n_tiles) / static_cast<uint32_t>(p.splits));
}
template <int BN>
__host__ __device__ inline int plan_inner(const GemmPlan& p, int item) {
  switch (p.mode) {
    case SCHED_ROW_PANEL: {
      if (p.splits == 1) return p.n_tiles;
      const int split = static_cast<int>(static_cast<uint32_t>(item) % static_cast<uint32_t>(p.splits));
      return row_panel_n0(p, split + 1) - row_panel_n0(p, split);
    }
    case SCHED_COL_PANEL: {
      if (p.splits == 1) return p.m_tiles;
      const int split = static_cast<int>(static_cast<uint32_t>(item) / static_cast<uint32_t>(p.n_tiles));
      return col_panel_m0(p, split + 1) - col_panel_m0(p, split);
    }
    default: return 1;
  }
}
template <int BN>
__device__ __forceinline__ int plan_tri_kb_end(const GemmPlan& p, int n) {
  int kb_end = p.kb_total;
  if (p.tri_k) {
    const int lim = ((n + 1) * BN + GEMM_BK - 1) / GEMM_BK;
    kb_end = lim < kb_end ? lim : kb_end;
  }
  return kb_end;
}
// First tile (inner = 0) of an item.
template <int BN>
__device__ __forceinline__ TileCoord plan_tile(const GemmPlan& p, int item, int inner = 0) {
  TileCoord t;
  uint32_t ksplit = 0, ksplits = 1;  // split-K index / count (tile schedules only)
  const uint32_t uitem = static_cast<uint32_t>(item);
  t.split = 0;
  if (p.mode == SCHED_ROW_PANEL) {
    // consecutive items share the M tile (its A panel stays hot in L2) and walk different N ranges
    if (p.splits == 1) {
      t.m = item;
      t.n = inner;
    } else {
      t.m = static_cast<int>(uitem / static_cast<uint32_t>(p.splits));
      t.split = item - t.m * p.splits;
      t.n = row_panel_n0(p, t.split) + inner;
    }
  } else if (p.mode == SCHED_COL_PANEL) {
    if (p.splits == 1) {
      t.n = item;
      t.m = inner;
    } else {
      t.split = static_cast<int>(uitem / static_cast<uint32_t>(p.n_tiles));
      t.n = item - t.split * p.n_tiles;
      t.m = col_panel_m0(p, t.split) + inner;
    }
  } else if (p.mode == SCHED_TRI_TILES) {
    const int tri = p.m_tiles * (p.m_tiles + 1) / 2;
    int idx = item;
    if (p.splits > 1) {
      ksplit = uitem / static_cast<uint32_t>(tri);
      idx = item - static_cast<int>(ksplit) * tri;
      ksplits = static_cast<uint32_t>(p.splits);
    }
    int m = static_cast<int>((sqrtf(8.0f * static_cast<float>(idx) + 1.0f) - 1.0f) * 0.5f);
    while ((m + 1) * (m + 2) / 2 <= idx) ++m;
    while (m * (m + 1) / 2 > idx) --m;
    t.m = m;
    t.n = idx - m * (m + 1) / 2;
  } else {
    const int tiles = p.m_tiles * p.n_tiles;
    int idx = item;
    if (p.splits > 1) {
      ksplit = uitem / static_cast<uint32_t>(tiles);
      idx = item - static_cast<int>(ksplit) * tiles;
      ksplits = static_cast<uint32_t>(p.splits);
    }
    t.m = static_cast<int>(static_cast<uint32_t>(idx) / static_cast<uint32_t>(p.n_tiles));
    t.n = idx - t.m * p.n_tiles;
  }
  const int kb_end = plan_tri_kb_end<BN>(p, t.n);
  if (ksplits == 1) {
    t.kb0 = 0;
    t.kb1 = kb_end;
  } else {
    t.kb0 = static_cast<int>(ksplit * static_cast<uint32_t>(kb_end) / ksplits);
    t.kb1 = static_cast<int>((ksplit + 1) * static_cast<uint32_t>(kb_end) / ksplits);
  }
  t.row0 = t.m * GEMM_BM;
  return t;
}
// Tile `inner` of the item whose first tile is `base`: panels advance by one N (row panel) or M (column panel) tile, which
// needs no division at all -- the role loops call this once per tile.
template <int BN>
__device__ __forceinline__ TileCoord plan_tile_at(const GemmPlan& p, const TileCoord& base, int inner) {
  TileCoord t = base;
  if (p.mode == SCHED_ROW_PANEL) {
    t.n = base.n + inner;
    t.kb1 = plan_tri_kb_end<BN>(p, t.n);  // (panels are never split along K: kb0 = 0)
  } else if (p.mode == SCHED_COL_PANEL) {
    t.m = base.m + inner;
    t.row0 = t.m * GEMM_BM;
  }
  return t;
}

// Number of ranges (1 .. max_splits) to cut every panel into so that `panels * splits` items fill whole rounds of `pairs`
// persistent CTA pairs: the smallest count whose round efficiency is within 2 % of the best (each extra range reloads the
// panel's stationary operand and flushes its partial results once more).  Host side.
inline int balanced_panel_splits(int panels, int inner_tiles, int pairs, int max_splits) {
  if (panels <= 0 || pairs <= 0) return 1;
  if (max_splits > inner_tiles) max_splits = inner_tiles;
  if (max_splits < 1) max_splits = 1;
  auto eff = [&](int sp) {
    const long long items = static_cast<long long>(panels) * sp;
    const long long rounds = (items + pairs - 1) / pairs;
    return static_cast<double>(items) / static_cast<double>(rounds * pairs);
  };
  double best = 0.0;
  for (int sp = 1; sp <= max_splits; ++sp) best = eff(sp) > best ? eff(sp) : best;
  for (int sp = 1; sp <= max_splits; ++sp)
    if (eff(sp) >= best - 0.02) return sp;
  return 1;
}

// Per-thread view the epilogue functors get.
struct EpiCtx {
  int ew;          // TMEM lane quadrant 0..3 of this warp (rows 32*ew .. 32*ew+31 of the CTA's slab)
  int wid;         // epilogue warp index 0..n_warps-1 (wid / 4 selects the column half when there are 8 warps)
  int n_warps;     // 4 or 8 epilogue warps
  int lane;        // lane in warp
  int M, N;        // valid extents
  float* scratch;        // Epi::SCRATCH_BYTES of shared memory (1024-byte aligned), shared by the four epilogue warps
  uint32_t scratch_u32;  // the same block as a shared-window address (for st.shared / ld.shared / TMA stores)
};

// ---------------------------------------------------------------------------------------------
// Output staging for TMA stores: every epilogue warp owns 32 accumulator rows; a "slab" is 32 rows x 128 bytes of
// shared memory in the 128-byte-swizzle pattern (16-byte chunk k of row r lives at chunk k ^ (r & 7)), which is what
// a SWIZZLE_128B tensor map with a {128 bytes, 32 rows} box reads.  Threads write their own row with conflict-free
// st.shared.v4 (each quarter-warp hits eight different 16-byte bank groups); one lane then issues the bulk store, so
// global memory sees full 128-byte lines instead of 32 scattered 16-byte pieces per instruction.
// ---------------------------------------------------------------------------------------------
constexpr int SLAB_BYTES = 32 * 128;

__device__ __forceinline__ void slab_write_f32(uint32_t slab, int lane, const float (&v)[32]) {
  const uint32_t row = slab + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    sts_v4(row + (static_cast<uint32_t>(k ^ (lane & 7)) << 4), v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
// 32 fp16 columns (64 bytes) into half `h` (0/1) of the thread's 128-byte slab row
__device__ __forceinline__ void slab_write_f16_half(uint32_t slab, int lane, int h, const float (&v)[32]) {
  const uint32_t row = slab + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h2 = __floats2half2_rn(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    sts_v4_u32(row + (static_cast<uint32_t>((4 * h + k) ^ (lane & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
}
// make the warp's st.shared visible to the async proxy, then one lane issues the bulk tensor store
__device__ __forceinline__ void slab_issue(const CUtensorMap* tm, uint32_t slab, int lane, int c0, int c1) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(tm)), "r"(slab), "r"(c0), "r"(c1)
                 : "memory");
  }
}
// same, as a bulk REDUCTION: global[box] += slab (fp32 add performed by the L2 on whole lines)
__device__ __forceinline__ void slab_issue_add(const CUtensorMap* tm, uint32_t slab, int lane, int c0, int c1) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(tm)), "r"(slab), "r"(c0), "r"(c1)
                 : "memory");
  }
}
__device__ __forceinline__ void slab_commit(int lane) {
  if (lane == 0) tma_store_commit();
}
// wait until at most N of this lane's committed bulk groups still read shared memory, then release the warp
template <int N>
__device__ __forceinline__ void slab_wait_free(int lane) {
  if (lane == 0) tma_store_wait_read<N>();
  __syncwarp();
}

__device__ __forceinline__ void epi_bar_sync(const EpiCtx& ctx) {
  asm volatile("bar.sync 1, %0;" ::"r"(ctx.n_warps * 32) : "memory");
}

template <int BN, int STAGES, class Epi>
constexpr size_t gemm_smem_bytes() {
  return 1024 + static_cast<size_t>(STAGES) * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + GEMM_AUX_BYTES +
         ((Epi::scratch_bytes(4) + 1023) / 1024) * 1024;
}

template <int BN, int STAGES, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmPlan plan,
               const __grid_constant__ typename Epi::Params ep) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32,256]");
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* scratch = reinterpret_cast<float*>(sB + STAGES * B_BYTES);  // 1024-byte aligned
  uint8_t* aux = sB + STAGES * B_BYTES + ((Epi::scratch_bytes(4) + 1023) / 1024) * 1024;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  static_assert((2 * STAGES + 4) * 8 + 4 <= GEMM_AUX_BYTES, "aux region too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = plan_num_items<BN>(plan);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n_inner = plan_inner<BN>(plan, item);
        const TileCoord tbase = plan_tile<BN>(plan, item);
        for (int inner = 0; inner < n_inner; ++inner) {
          const TileCoord tc = plan_tile_at<BN>(plan, tbase, inner);
          int seg_v = 0, seg_off = tc.kb0;  // virtual segment / K block inside it (split plans always start at kb0 = 0)
          for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
            int col_a = kb * GEMM_BK, col_b = col_a;
            if (plan.seg_kb != 0) {
              col_a = (static_cast<int>((plan.a_seg_mask >> seg_v) & 1u) * plan.seg_kb + seg_off) * GEMM_BK;
              col_b = (static_cast<int>((plan.b_seg_mask >> seg_v) & 1u) * plan.seg_kb + seg_off) * GEMM_BK;
              if (++seg_off == plan.seg_kb) {
                seg_off = 0;
                ++seg_v;
              }
            }
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], plan.a_tx_bytes + plan.b_tx_bytes);
            if (plan.a_is_3d) {
              tma_load_3d(sA + stage * A_BYTES, &tmA, &full_bar[stage], kb * GEMM_BK, 0, tc.m * plan.a_outer_step);
            } else {
              tma_load_2d(sA + stage * A_BYTES, &tmA, &full_bar[stage], col_a, tc.m * plan.a_outer_step);
            }
            tma_load_2d(sB + stage * B_BYTES, &tmB, &full_bar[stage], col_b, tc.n * BN);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n_inner = plan_inner<BN>(plan, item);
        const TileCoord tbase = plan_tile<BN>(plan, item);
        for (int inner = 0; inner < n_inner; ++inner) {
          const TileCoord tc = plan_tile_at<BN>(plan, tbase, inner);
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
          for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = make_smem_desc_kmajor_sw128(smem_u32(sA + stage * A_BYTES));
            const uint64_t db = make_smem_desc_kmajor_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
            for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
              // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in 16-byte units
              umma_f16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), plan.idesc,
                          (kb > tc.kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&tfull_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ------------------------------------------------ epilogue warps
    EpiCtx ctx;
    ctx.ew = warp - GEMM_EPI_WARP0;
    ctx.wid = ctx.ew;
    ctx.n_warps = 4;
    ctx.lane = lane;
    ctx.M = plan.M;
    ctx.N = plan.N;
    ctx.scratch = scratch;
    ctx.scratch_u32 = smem_u32(scratch);
    typename Epi::State st;
    Epi::kernel_begin(st, ep, ctx);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int n_inner = plan_inner<BN>(plan, item);
      Epi::item_begin(st, ep, ctx, plan_tile<BN>(plan, item, 0));
      for (int inner = 0; inner < n_inner; ++inner) {
        TileCoord tc = plan_tile<BN>(plan, item, inner);
        tc.row0_next = (plan.mode == SCHED_COL_PANEL && inner + 1 < n_inner) ? tc.row0 + GEMM_BM : -1;
        Epi::tile_begin(st, ep, ctx, tc);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ctx.ew * 32) << 16) + static_cast<uint32_t>(acc * BN);
        const int n_valid = plan.N - tc.n * BN;
        if constexpr (Epi::UNROLL_CHUNKS) {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            if (!Epi::ALL_CHUNKS && c * 32 >= n_valid) break;
            float v[32];
            tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            Epi::chunk(st, ep, ctx, tc, v, c);
          }
        } else {  // one copy of the epilogue body: large functors otherwise thrash the instruction cache
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            if (!Epi::ALL_CHUNKS && c * 32 >= n_valid) break;
            float v[32];
            tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            Epi::chunk(st, ep, ctx, tc, v, c);
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        Epi::tile_end(st, ep, ctx, tc);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      Epi::item_end(st, ep, ctx, plan_tile<BN>(plan, item, 0));
    }
    Epi::kernel_end(st, ep, ctx);  // e.g. drain outstanding bulk stores before shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host-side launch helper
// ---------------------------------------------------------------------------------------------
template <int BN>
inline GemmPlan make_plan(int M, int N, int K_padded, int mode, int splits, int a_fmt, int b_fmt) {
  GemmPlan p{};
  p.M = M;
  p.N = N;
  p.kb_total = K_padded / GEMM_BK;
  p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.mode = mode;
  p.splits = splits < 1 ? 1 : splits;
  if (p.splits > p.kb_total) p.splits = p.kb_total;
  p.tri_k = 0;
  p.idesc = make_idesc_f16(GEMM_BM, BN, a_fmt, b_fmt);
  p.a_is_3d = 0;
  p.a_outer_step = GEMM_BM;
  p.a_tx_bytes = GEMM_BM * GEMM_BK * 2;
  p.b_tx_bytes = BN * GEMM_BK * 2;
  p.seg_kb = 0;
  p.kb_alt = 0x7fffffff;
  p.idesc_alt = 0;
  return p;
}

// Three-product plan over operands stored as [hi | lo] with `seg` (multiple of 64) elements per segment.
template <int BN>
inline GemmPlan make_split_plan(int M, int N, int seg, int mode, int fmt) {
  GemmPlan p = make_plan<BN>(M, N, 3 * seg, mode, 1, fmt, fmt);
  p.seg_kb = seg / GEMM_BK;
  p.a_seg_mask = 0b010u;  // hi lo hi
  p.b_seg_mask = 0b100u;  // hi hi lo
  return p;
}

template <int BN, int STAGES, class Epi>
inline int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmPlan& plan,
                       const typename Epi::Params& ep, cudaStream_t stream, int tag) {
  if (plan.kb_total <= 0 || plan.M <= 0 || plan.N <= 0) return BVLM_EINVAL;
  constexpr size_t smem = gemm_smem_bytes<BN, STAGES, Epi>();
  static_assert(smem <= 232448, "shared memory budget exceeded");
  auto kfn = gemm_tn_kernel<BN, STAGES, Epi>;
  static std::atomic<int> configured_dev[BVLM_MAX_DEVICES];  // per instantiation AND per device
  std::atomic<int>& configured = configured_dev[current_device_slot()];
  if (!configured.load(std::memory_order_acquire)) {
    BVLM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured.store(1, std::memory_order_release);
  }
  const int items = plan_num_items<BN>(plan);
  int grid = device_sm_count();
  if (items < grid) grid = items;
  timing_begin(tag, stream);
  kfn<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, plan, ep);
  timing_end(tag, stream);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

// 16-bit K-major operand descriptor handed around on the host.
struct Operand16 {
  const void* ptr;
  int64_t rows;    // valid rows
  int64_t k_pad;   // padded K (multiple of 64)
  int fmt;         // OperandFormat
  int64_t pitch = 0;  // row pitch in elements (0: k_pad)
};

// Row pitch for library-owned K-major operands: pitches that are a multiple of 1 KiB alias in the L2 slice hash when a
// tile's rows are fetched 128 bytes at a time, so such pitches get one extra 128-byte line (BVLM_PITCH_PAD overrides).
int64_t operand_pitch(int64_t k_elems);

template <int BOX_ROWS>
inline int operand_tmap(CUtensorMap* out, const Operand16& op) {
  const int64_t pitch = op.pitch > 0 ? op.pitch : op.k_pad;
  return make_tmap_2d(out, op.ptr, op.fmt == FMT_BF16 ? TM_BF16 : TM_F16, static_cast<uint64_t>(op.k_pad),
                      static_cast<uint64_t>(op.rows), static_cast<uint64_t>(pitch) * 2, GEMM_BK, BOX_ROWS, 1);
}

}  // namespace bvlm
