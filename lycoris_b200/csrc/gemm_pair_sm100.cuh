// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM: two CTAs on one TPC cooperate on
// a 256 x BLOCK_N output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile, so the
// L2 -> SM feed per CTA drops from (128 + BLOCK_N) to (128 + BLOCK_N/2) rows per k-block — the 1-CTA
// kernel is feed-bound at 128x256 (~96 B/clk/SM), the pair is balanced against the tensor pipe.
//
//   leader CTA (cluster rank 0): one thread issues tcgen05.mma.cta_group::2 (M = 256) for both
//   both CTAs: TMA producer (signals the LEADER's full barrier), TMEM alloc (cta_group::2), epilogue of
//              their own 128 accumulator rows
// Barriers (same smem offsets in both CTAs):
//   full[s]   leader only, count 2 (one arrive.expect_tx per CTA producer) + TMA bytes of both CTAs
//   empty[s]  per CTA, count 1, multicast tcgen05.commit from the leader's MMA thread
//   tfull[a]  per CTA, count 1, multicast tcgen05.commit
//   tempty[a] leader only, count 8 (4 epilogue warps x 2 CTAs; the peer arrives remotely)
#pragma once
#include "gemm_sm100.cuh"

namespace lyco {

constexpr int PAIR_BLOCK_M = 256;

template <int BLOCK_N>
struct PairCfg {
  static constexpr int B_BYTES = (BLOCK_N / 2) * GEMM_BLOCK_K * 2;  // this CTA's half of B
  static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
  static constexpr int STAGES = (GEMM_RING_BYTES / STAGE_BYTES) > 8 ? 8 : (GEMM_RING_BYTES / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = PairCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int HALF_N = BLOCK_N / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 2);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pairs = gridDim.x >> 1;
  const int pair_id = blockIdx.x >> 1;
  const int total = p.m_tiles * p.n_tiles * p.splits;  // m_tiles counts 256-row pair tiles

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------- TMA producer (every CTA)
    int stage = 0;
    uint32_t phase = 0;
    for (int w = pair_id; w < total; w += pairs) {
      const int split = w % p.splits;
      const int tile = w / p.splits;
      const int n_idx = tile % p.n_tiles;
      const int m_idx = tile / p.n_tiles;
      const int kb0 = static_cast<int>(static_cast<int64_t>(split) * p.k_blocks / p.splits);
      const int kb1 = static_cast<int>(static_cast<int64_t>(split + 1) * p.k_blocks / p.splits);
      const int m0 = m_idx * PAIR_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M;
      const int n0 = n_idx * BLOCK_N + static_cast<int>(rank) * HALF_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
        ptx::mbar_expect_tx_cluster(full_leader, STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + GEMM_A_BYTES;
        if (!A_MN) {
          ptx::tma_load_2d_pair(sa, &tmap_a, full_leader, kb * GEMM_BLOCK_K, m0);
        } else {
#pragma unroll
          for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
            ptx::tma_load_2d_pair(sa + j * GEMM_ATOM_BYTES, &tmap_a, full_leader, m0 + j * 64, kb * GEMM_BLOCK_K);
        }
        if (!B_MN) {
          ptx::tma_load_2d_pair(sb, &tmap_b, full_leader, kb * GEMM_BLOCK_K, n0);
        } else {
#pragma unroll
          for (int j = 0; j < HALF_N / 64; ++j)
            ptx::tma_load_2d_pair(sb + j * GEMM_ATOM_BYTES, &tmap_b, full_leader, n0 + j * 64, kb * GEMM_BLOCK_K);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------------------------- MMA issuer (leader only)
    const uint32_t idesc = ptx::make_idesc_f16(p.fmt, PAIR_BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair_id; w < total; w += pairs) {
      const int split = w % p.splits;
      const int kb0 = static_cast<int>(static_cast<int64_t>(split) * p.k_blocks / p.splits);
      const int kb1 = static_cast<int>(static_cast<int64_t>(split + 1) * p.k_blocks / p.splits);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + GEMM_A_BYTES;
#pragma unroll
        for (int kk = 0; kk < GEMM_BLOCK_K / GEMM_UMMA_K; ++kk) {
          const uint64_t da = A_MN ? ptx::make_smem_desc(sa + kk * 2048, GEMM_ATOM_BYTES, 1024)
                                   : ptx::make_smem_desc(sa + kk * 32, 16, 1024);
          const uint64_t db = B_MN ? ptx::make_smem_desc(sb + kk * 2048, GEMM_ATOM_BYTES, 1024)
                                   : ptx::make_smem_desc(sb + kk * 32, 16, 1024);
          ptx::umma_f16_pair(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit_pair(&empty_bar[stage], 0b11);  // frees this slot in BOTH CTAs
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit_pair(&tfull_bar[acc], 0b11);  // accumulator halves complete in both CTAs
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (every CTA)
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair_id; w < total; w += pairs) {
      const int tile = w / p.splits;
      const int n_idx = tile % p.n_tiles;
      const int m_idx = tile / p.n_tiles;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m_idx * PAIR_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M + ew * 32 + lane;
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_row + c * 32, r);
        ptx::tmem_ld_wait();
        if (c == BLOCK_N / 32 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty_bar[acc]), 0));
        }
        store_chunk<EPI>(r, row, n_idx * BLOCK_N + c * 32, p);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();  // neither CTA may exit (or free TMEM) while its partner still uses it
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace lyco
