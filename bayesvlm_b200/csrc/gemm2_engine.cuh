// Two-CTA (cta_group::2) variant of the tcgen05 GEMM engine:  D[M,N] = A[M,K] * B[N,K]^T on CTA PAIRS.
//
// A cluster of two CTAs (two SMs of one TPC) computes a 256 x BN output tile with tcgen05.mma.cta_group::2 (M = 256):
//   * CTA r of the pair owns accumulator rows [256 m + 128 r, +128) in ITS tensor memory and loads ITS 128 x 64 slice of
//     A plus HALF of the B tile (BN/2 x 64) per K block -- 32 KB per stage instead of 48 KB, so six stages fit where the
//     one-CTA engine has four (the one-CTA main loop was TMA-latency bound: ncu, profiles/r1a_*), and every B byte is
//     fetched from L2 once per pair instead of once per CTA;
//   * both producers signal ONE full barrier in the leader CTA (cp.async.bulk.tensor ... .cta_group::2 with the peer bit
//     of the barrier address cleared); the leader's elected thread issues the MMAs, and tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to the barriers at the same offset in both CTAs;
//   * 4 or 8 epilogue warps per CTA drain the CTA's own 128 TMEM lanes (8 warps: two per lane quadrant, each taking half
//     of the BN columns) and run the same pluggable epilogue functors as the one-CTA engine.
// Schedules: SCHED_TILES (optionally split-K), SCHED_ROW_PANEL, SCHED_COL_PANEL, with M tiles of 256 rows.
#pragma once
#include <cstdlib>

#include "gemm_engine.cuh"

namespace bvlm {

constexpr int GEMM2_BM = 256;         // rows per cluster tile
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (+ expect_tx) on the barrier at the same offset in the leader CTA, from either CTA of the pair
__device__ __forceinline__ void mbar_arrive_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(bar) & PEER_BIT_MASK), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
// TMA load into THIS CTA's shared memory, transaction bytes credited to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// same, multicast to every CTA of the cluster whose bit is set in cta_mask (same shared-memory offset in each; the
// transaction bytes are credited to the leader CTA of each destination's pair)
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0,
                                                    int32_t c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1),
        "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0,
                                                 int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `cta_mask` once all previously issued MMAs of the pair are done
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

template <int BN, int STAGES, int EPI_WARPS, class Epi>
constexpr size_t gemm2_smem_bytes() {
  return 1024 + static_cast<size_t>(STAGES) * (GEMM_BM * GEMM_BK * 2 + (BN / 2) * GEMM_BK * 2) + GEMM_AUX_BYTES +
         ((Epi::scratch_bytes(EPI_WARPS) + 1023) / 1024) * 1024;
}

// A_MN / B_MN: the operand is MN-major (its M / N index is contiguous in global memory, i.e. the tensor is stored
// [K rows, M or N columns] row-major -- a plain row-major activation matrix is the "transposed" operand of X^T X without
// any transposing pass). Such a tile is fetched as 64-column boxes of [64 K rows x 128 bytes].
// PAIRS = 2: clusters of FOUR CTAs = two MMA pairs working on two consecutive 256-row M tiles of the SAME N tile; the B
// tile is fetched ONCE per cluster and multicast into both pairs (pair 0's producers issue it), which cuts the operand
// traffic over the L2 fabric by a quarter -- the bound of the FP8-assisted predictive GEMM.
template <int BN, int STAGES, int EPI_WARPS, class Epi, bool A_MN = false, bool B_MN = false, int PAIRS = 1>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8, const GemmPlan plan,
                const __grid_constant__ typename Epi::Params ep) {
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 64 in [64,256]");
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8 || EPI_WARPS == 16, "4, 8 or 16 epilogue warps");
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;    // this CTA's 128 rows of A
  constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;   // this CTA's half of the B tile
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr int CHUNKS = BN / 32;
  constexpr int CH_PER_WARP = CHUNKS / (EPI_WARPS / 4);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* scratch = reinterpret_cast<float*>(sB + STAGES * B_BYTES);
  uint8_t* aux = sB + STAGES * B_BYTES + ((Epi::scratch_bytes(EPI_WARPS) + 1023) / 1024) * 1024;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);   // used in the leader CTA: 2 producer arrivals + tx bytes
  uint64_t* empty_bar = full_bar + STAGES;                  // per CTA: 1 arrival (multicast commit)
  uint64_t* tfull_bar = empty_bar + STAGES;                 // per CTA: 1 arrival (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;                     // used in the leader CTA: 2 * EPI_WARPS arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  static_assert((2 * STAGES + 4) * 8 + 4 <= GEMM_AUX_BYTES, "aux region too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();  // 0 .. 2 * PAIRS - 1
  const uint32_t rank = crank & 1u;          // rank inside the MMA pair
  const uint32_t pair = crank >> 1;
  const bool leader = rank == 0;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair));
  const uint16_t all_mask = static_cast<uint16_t>((1u << (2 * PAIRS)) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], PAIRS);  // every pair's MMAs must be done with a stage before anybody refills it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // (programmatic dependent launch: everything above overlapped the tail of the preceding kernel; no-ops for a normal launch)
  grid_launch_dependents();
  grid_dependency_wait();

  const int n_items = plan_num_items<BN>(plan);
  const int cluster_id = blockIdx.x / (2 * PAIRS);
  const int n_clusters = gridDim.x / (2 * PAIRS);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (one per CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int n_inner = plan_inner<BN>(plan, item);
        const TileCoord tbase = plan_tile<BN>(plan, item);
        for (int inner = 0; inner < n_inner; ++inner) {
          const TileCoord tc = plan_tile_at<BN>(plan, tbase, inner);
          const int row_a = (tc.m * PAIRS + static_cast<int>(pair)) * GEMM2_BM + static_cast<int>(rank) * GEMM_BM;
          const int row_b = tc.n * BN + static_cast<int>(rank) * (BN / 2);
          const bool load_b = PAIRS == 1 || pair == 0;  // pair 0 multicasts the shared B tile to the other pair
          const uint16_t b_mask = static_cast<uint16_t>(0x5u << rank);  // CTAs with the same in-pair rank
          int seg_v = 0, seg_off = tc.kb0;
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
            mbar_wait<true>(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx_leader(&full_bar[stage], plan.a_tx_bytes + B_BYTES);
            if (kb >= plan.kb_alt) {  // FP8 phase: same 128-byte K blocks, 128 elements each, second pair of tensor maps
              const int col8 = (kb - plan.kb_alt) * 128;
              tma_load_2d_pair(sA + stage * A_BYTES, &tmA8, &full_bar[stage], col8, row_a);
              if (PAIRS == 1) tma_load_2d_pair(sB + stage * B_BYTES, &tmB8, &full_bar[stage], col8, row_b);
              else if (load_b) tma_load_2d_pair_mc(sB + stage * B_BYTES, &tmB8, &full_bar[stage], col8, row_b, b_mask);
            } else {
            if constexpr (A_MN) {  // [K, M] storage: two 64-column boxes of 64 K rows
#pragma unroll
              for (int blk = 0; blk < GEMM_BM / 64; ++blk)
                tma_load_2d_pair(sA + stage * A_BYTES + blk * (GEMM_BK * 128), &tmA, &full_bar[stage], row_a + blk * 64, col_a);
            } else if (plan.a_is_3d) {  // a_outer_step outer items per CTA slab (EPIG: pool rows x classes x K)
              tma_load_3d_pair(sA + stage * A_BYTES, &tmA, &full_bar[stage], col_a, 0,
                               ((tc.m * PAIRS + static_cast<int>(pair)) * 2 + static_cast<int>(rank)) * plan.a_outer_step);
            } else {
              tma_load_2d_pair(sA + stage * A_BYTES, &tmA, &full_bar[stage], col_a, row_a);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int blk = 0; blk < (BN / 2) / 64; ++blk)
                tma_load_2d_pair(sB + stage * B_BYTES + blk * (GEMM_BK * 128), &tmB, &full_bar[stage], row_b + blk * 64, col_b);
            } else if (PAIRS == 1) {
              tma_load_2d_pair(sB + stage * B_BYTES, &tmB, &full_bar[stage], col_b, row_b);
            } else if (load_b) {
              tma_load_2d_pair_mc(sB + stage * B_BYTES, &tmB, &full_bar[stage], col_b, row_b, b_mask);
            }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int n_inner = plan_inner<BN>(plan, item);
        const TileCoord tbase = plan_tile<BN>(plan, item);
        for (int inner = 0; inner < n_inner; ++inner) {
          const TileCoord tc = plan_tile_at<BN>(plan, tbase, inner);
          mbar_wait<true>(&tempty_bar[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
          for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = A_MN ? make_smem_desc_mnmajor_sw128(smem_u32(sA + stage * A_BYTES), GEMM_BK * 128)
                                     : make_smem_desc_kmajor_sw128(smem_u32(sA + stage * A_BYTES));
            const uint64_t db = B_MN ? make_smem_desc_mnmajor_sw128(smem_u32(sB + stage * B_BYTES), GEMM_BK * 128)
                                     : make_smem_desc_kmajor_sw128(smem_u32(sB + stage * B_BYTES));
            // one UMMA consumes 16 K elements: 32 bytes along a K-major row, 16 rows (2048 bytes) of an MN-major block
            constexpr uint64_t STEP_A = A_MN ? (16 * 128) >> 4 : 2, STEP_B = B_MN ? (16 * 128) >> 4 : 2;
            if (kb >= plan.kb_alt) {  // FP8 phase: K = 32 per instruction = 32 bytes, same descriptor step as fp16
              const uint64_t da8 = make_smem_desc_kmajor_sw128(smem_u32(sA + stage * A_BYTES));
              const uint64_t db8 = make_smem_desc_kmajor_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f8_ss_pair(tmem_d, da8 + static_cast<uint64_t>(2 * k), db8 + static_cast<uint64_t>(2 * k), plan.idesc_alt,
                                (kb > tc.kb0 || k > 0) ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
                umma_f16_ss_pair(tmem_d, da + STEP_A * static_cast<uint64_t>(k), db + STEP_B * static_cast<uint64_t>(k),
                                 plan.idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit_pair(&empty_bar[stage], all_mask);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_pair(&tfull_bar[acc], pair_mask);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ------------------------------------------------ epilogue warps (both CTAs, own TMEM lanes)
    EpiCtx ctx;
    ctx.wid = warp - GEMM_EPI_WARP0;
    ctx.ew = ctx.wid & 3;
    ctx.n_warps = EPI_WARPS;
    ctx.lane = lane;
    ctx.M = plan.M;
    ctx.N = plan.N;
    ctx.scratch = scratch;
    ctx.scratch_u32 = smem_u32(scratch);
    const int chalf = ctx.wid >> 2;
    typename Epi::State st;
    Epi::kernel_begin(st, ep, ctx);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int n_inner = plan_inner<BN>(plan, item);
      TileCoord t0 = plan_tile<BN>(plan, item, 0);
      t0.row0 = (t0.m * PAIRS + static_cast<int>(pair)) * GEMM2_BM + static_cast<int>(rank) * GEMM_BM;
      t0.row0_next = -1;
      Epi::item_begin(st, ep, ctx, t0);
      for (int inner = 0; inner < n_inner; ++inner) {
        TileCoord tc = plan_tile_at<BN>(plan, t0, inner);
        tc.row0 = (tc.m * PAIRS + static_cast<int>(pair)) * GEMM2_BM + static_cast<int>(rank) * GEMM_BM;
        tc.row0_next = (plan.mode == SCHED_COL_PANEL && inner + 1 < n_inner) ? tc.row0 + PAIRS * GEMM2_BM : -1;
        Epi::tile_begin(st, ep, ctx, tc);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ctx.ew * 32) << 16) + static_cast<uint32_t>(acc * BN);
        const int n_valid = plan.N - tc.n * BN;
        if constexpr (Epi::DRAIN_FIRST && EPI_WARPS == 8) {
          // Drain this warp's whole share of the accumulator into registers and hand the TMEM buffer back to the MMA
          // issuer BEFORE the (store-latency bound) epilogue work: tensor memory is then occupied for ~1 us per tile,
          // whatever the output path takes.
          float vv[CH_PER_WARP][32];
#pragma unroll
          for (int i = 0; i < CH_PER_WARP; ++i)
            tmem_ld_32x32(taddr + static_cast<uint32_t>((chalf * CH_PER_WARP + i) * 32), vv[i]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
#pragma unroll
          for (int i = 0; i < CH_PER_WARP; ++i) {
            const int c = chalf * CH_PER_WARP + i;
            if (!Epi::ALL_CHUNKS && c * 32 >= n_valid) continue;
            Epi::chunk(st, ep, ctx, tc, vv[i], c);
          }
          Epi::tile_end(st, ep, ctx, tc);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
          continue;
        }
        if constexpr (Epi::UNROLL_CHUNKS) {
#pragma unroll
          for (int c = 0; c < CHUNKS; ++c) {
            if (c / CH_PER_WARP != chalf) continue;
            if (!Epi::ALL_CHUNKS && c * 32 >= n_valid) continue;
            float v[32];
            tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            Epi::chunk(st, ep, ctx, tc, v, c);
          }
        } else {  // one copy of the epilogue body: large functors otherwise thrash the instruction cache
#pragma unroll 1
          for (int c = chalf * CH_PER_WARP; c < (chalf + 1) * CH_PER_WARP; ++c) {
            if (!Epi::ALL_CHUNKS && c * 32 >= n_valid) break;
            float v[32];
            tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            Epi::chunk(st, ep, ctx, tc, v, c);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
        Epi::tile_end(st, ep, ctx, tc);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      Epi::item_end(st, ep, ctx, t0);
    }
    Epi::kernel_end(st, ep, ctx);
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the peer may still read its shared memory or signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN>
inline GemmPlan make_plan2(int M, int N, int K_padded, int mode, int splits, int fmt) {
  GemmPlan p = make_plan<BN>(M, N, K_padded, mode, splits, fmt, fmt);
  p.m_tiles = (M + GEMM2_BM - 1) / GEMM2_BM;
  p.idesc = make_idesc_f16(GEMM2_BM, BN, fmt, fmt);
  return p;
}
// tensor map of an MN-major operand stored [k_rows, mn_cols] row-major (pitch in elements), 16-bit
inline int operand_tmap_mn(CUtensorMap* out, const void* ptr, int64_t k_rows, int64_t mn_cols, int64_t pitch, int fmt) {
  return make_tmap_2d(out, ptr, fmt == FMT_BF16 ? TM_BF16 : TM_F16, static_cast<uint64_t>(mn_cols),
                      static_cast<uint64_t>(k_rows), static_cast<uint64_t>(pitch) * 2, 64, GEMM_BK, 1);
}

template <int BN>
inline GemmPlan make_split_plan2(int M, int N, int seg, int mode, int fmt) {
  GemmPlan p = make_split_plan<BN>(M, N, seg, mode, fmt);
  p.m_tiles = (M + GEMM2_BM - 1) / GEMM2_BM;
  p.idesc = make_idesc_f16(GEMM2_BM, BN, fmt, fmt);
  return p;
}

// PAIRS = 2 clusters cover 512 rows per item
inline void plan_use_pairs(GemmPlan& p, int pairs) { p.m_tiles = (p.M + pairs * GEMM2_BM - 1) / (pairs * GEMM2_BM); }

// number of CTA pairs that can be co-resident for this instantiation (queried once)
template <int BN, int STAGES, int EPI_WARPS, class Epi, bool A_MN = false, bool B_MN = false, int PAIRS = 1>
inline int gemm2_max_clusters(size_t smem) {
  static std::atomic<int> cached_dev[BVLM_MAX_DEVICES];  // per device (zero = not queried yet)
  std::atomic<int>& cached = cached_dev[current_device_slot()];
  if (const int c = cached.load(std::memory_order_relaxed); c > 0) return c;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(device_sm_count() / (2 * PAIRS) * (2 * PAIRS)), 1, 1);
  cfg.blockDim = dim3(128 + 32 * EPI_WARPS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * PAIRS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm2_tn_kernel<BN, STAGES, EPI_WARPS, Epi, A_MN, B_MN, PAIRS>, &cfg) != cudaSuccess ||
      n <= 0) {
    (void)cudaGetLastError();
    n = device_sm_count() / (2 * PAIRS);
  }
  cached.store(n, std::memory_order_relaxed);
  return n;
}

// K-major B operand: tensor map box rows = BN / 2 (each CTA of the pair loads half of the B tile).
// MN-major operands: tensor map over the [K, M|N] storage with box {64 columns, 64 K rows} (operand_tmap_mn).
template <int BN, int STAGES, int EPI_WARPS, class Epi, bool A_MN = false, bool B_MN = false, int PAIRS = 1>
inline int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmPlan& plan,
                        const typename Epi::Params& ep, cudaStream_t stream, int tag, const CUtensorMap* tmA8 = nullptr,
                        const CUtensorMap* tmB8 = nullptr, bool pdl = false) {
  if (plan.kb_total <= 0 || plan.M <= 0 || plan.N <= 0) return BVLM_EINVAL;
  constexpr size_t smem = gemm2_smem_bytes<BN, STAGES, EPI_WARPS, Epi>();
  static_assert(smem <= 232448, "shared memory budget exceeded");
  auto kfn = gemm2_tn_kernel<BN, STAGES, EPI_WARPS, Epi, A_MN, B_MN, PAIRS>;
  static std::atomic<int> configured_dev[BVLM_MAX_DEVICES];  // per instantiation AND per device (the attribute is per device)
  std::atomic<int>& configured = configured_dev[current_device_slot()];
  if (!configured.load(std::memory_order_acquire)) {
    BVLM_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured.store(1, std::memory_order_release);
  }
  const int items = plan_num_items<BN>(plan);
  int clusters = gemm2_max_clusters<BN, STAGES, EPI_WARPS, Epi, A_MN, B_MN, PAIRS>(smem);
  if (items < clusters) clusters = items;
#ifdef BVLM_DIAG  // diagnostic build only: run on fewer CTA pairs (is a limit per SM or chip wide?)
  if (const char* e = getenv("BVLM_DEBUG_CLUSTERS"); e != nullptr && atoi(e) > 0 && atoi(e) < clusters) clusters = atoi(e);
#endif
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * PAIRS * clusters), 1, 1);
  cfg.blockDim = dim3(128 + 32 * EPI_WARPS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * PAIRS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  timing_begin(tag, stream);
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kfn, tmA, tmB, tmA8 != nullptr ? *tmA8 : tmA, tmB8 != nullptr ? *tmB8 : tmB,
                                            plan, ep);
  timing_end(tag, stream);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  BVLM_CUDA_TRY(cudaGetLastError());
  return BVLM_OK;
}

}  // namespace bvlm
