// Implicit-GEMM 2-D convolution on the tcgen05 GEMM pipeline: the activation operand is gathered by
// TMA in im2col mode (cp.async.bulk.tensor.4d ... .im2col) straight from the NHWC tensor — no unfold
// buffer, no cuDNN.  Same warp roles / mbarrier pipelines / TMEM double buffering as gemm_sm100.cuh.
//
//   MODE_FPROP  Y[(n,p,q), o]  = sum_{r,s,c} X[n, p*st-pad+r, q*st-pad+s, c] * Wk[o, (r,s,c)]   (+ bias)
//               A = im2col(X): 128 output pixels x 64 channels per (r, s, channel-block) step, K-major
//               B = Wk [O, R*S*C] K-major, plain tiled TMA.   dgrad runs the same kernel on dY with the
//               flipped / transposed filter [C, (r,s,o)] and pad' = R-1-pad (stride 1).
//   MODE_WGRAD  dW[o, (r,s,c)] = sum_{(n,p,q)} dY[(n,p,q), o] * X[n, p*st-pad+r, q*st-pad+s, c]
//               A = dY [M, O] MN-major (64-pixel x 64-channel boxes), B = im2col(X) 64 pixels x 64 channels:
//               an im2col box IS an MN-major operand atom (row = reduction index, 128 B = 64 channels).
//               Output columns are tiled per filter tap; reduction over pixels is split across CTAs.
#pragma once
#include "gemm_sm100.cuh"

namespace lyco {

enum { MODE_FPROP = 1, MODE_WGRAD = 2 };

struct ConvParams {
  GemmParams g;       // g.M/N/K: GEMM view (fprop: M = pixels, N = O, K = R*S*C; wgrad: M = O, N = R*S*C, K = pixels)
  int PQ, Q;          // output pixels per image / per row
  int C, CB;          // input channels of the gathered tensor, channel blocks of 64
  int S;              // filter width (taps per filter row)
  int stride;         // traversal stride
  int low_w, low_h;   // pixel-box lower corner (= -pad)
  int tiles_per_tap;  // wgrad: N-tiles per filter tap
};

__device__ __forceinline__ void tma_im2col_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c,
                                              int32_t w, int32_t h, int32_t n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}

// Same gather issued by a member of a CTA pair: completion is signalled on the LEADER's mbarrier.
__device__ __forceinline__ void tma_im2col_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                   int32_t c, int32_t w, int32_t h, int32_t n, uint16_t off_w,
                                                   uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}

// PAIR: a CTA pair (cta_group::2) computes a 256-row tile; each member gathers its own 128 rows of A and half of the
// B columns, the leader issues the MMAs (same arrangement as gemm_sm100_kernel<true, ...>).
template <bool PAIR, int MODE, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const ConvParams cp) {
  constexpr bool A_MN = (MODE == MODE_WGRAD);
  constexpr bool B_MN = (MODE == MODE_WGRAD);
  const GemmParams& p = cp.g;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* epi_smem = smem + GEMM_RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + GEMM_EPI_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* tlast_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tlast_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0;  // 0 = leader
  const int block_n = p.block_n;
  const int b_cols = PAIR ? block_n / 2 : block_n;  // B columns (N) this CTA stages
  const int stage_bytes = GEMM_A_BYTES + b_cols * GEMM_BLOCK_K * 2;
  const int stages = p.stages;
  const int rows_per_tile = PAIR ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (epi_uses_tma(EPI)) ptx::prefetch_tmap(&tmap_c);
  }
  gemm_setup<PAIR>(full_bar, empty_bar, tfull_bar, tempty_bar, tlast_bar, tmem_slot, stages, warp, lane);
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // everything above overlapped the previous kernel; its results are visible from here (pdl.cuh)
  pdl_trigger();

  const int workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (WorkIter it(p, worker, workers); it.next();) {
      const int n_idx = it.tile % p.n_tiles;
      const int m_idx = it.tile / p.n_tiles;
      const int kb0 = it.kb0, kb1 = it.kb1;
      if (MODE == MODE_FPROP) {
        // base output pixel of this CTA's 128 rows -> (image, row, col) -> input-space pixel coordinate
        const int m0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M;
        const int n0 = n_idx * block_n + static_cast<int>(rank) * b_cols;
        const int img = m0 / cp.PQ, rem = m0 - img * cp.PQ;
        const int py = rem / cp.Q, px = rem - py * cp.Q;
        const int w0 = px * cp.stride + cp.low_w, h0 = py * cp.stride + cp.low_h;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tap = kb / cp.CB, cb = kb - tap * cp.CB;
          const int fr = tap / cp.S, fs = tap - fr * cp.S;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + GEMM_A_BYTES;
          if (PAIR) {
            const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
            ptx::mbar_expect_tx_cluster(full_leader, stage_bytes);
            tma_im2col_4d_pair(sa, &tmap_a, full_leader, cb * 64, w0, h0, img, static_cast<uint16_t>(fs),
                               static_cast<uint16_t>(fr));
            ptx::tma_load_2d_pair(sb, &tmap_b, full_leader, kb * GEMM_BLOCK_K, n0);
          } else {
            ptx::mbar_expect_tx(&full_bar[stage], stage_bytes);
            tma_im2col_4d(sa, &tmap_a, &full_bar[stage], cb * 64, w0, h0, img, static_cast<uint16_t>(fs),
                          static_cast<uint16_t>(fr));
            ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      } else {
        const int tap = n_idx / cp.tiles_per_tap;
        const int c0 = (n_idx - tap * cp.tiles_per_tap) * block_n + static_cast<int>(rank) * b_cols;
        const int fr = tap / cp.S, fs = tap - fr * cp.S;
        const int o0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int m0 = kb * GEMM_BLOCK_K;  // first pixel of this reduction block
          const int img = m0 / cp.PQ, rem = m0 - img * cp.PQ;
          const int py = rem / cp.Q, px = rem - py * cp.Q;
          const int w0 = px * cp.stride + cp.low_w, h0 = py * cp.stride + cp.low_h;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + GEMM_A_BYTES;
          if (PAIR) {
            const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
            ptx::mbar_expect_tx_cluster(full_leader, stage_bytes);
#pragma unroll
            for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
              ptx::tma_load_2d_pair(sa + j * GEMM_ATOM_BYTES, &tmap_a, full_leader, o0 + j * 64, kb * GEMM_BLOCK_K);
            for (int j = 0; j < b_cols / 64; ++j)
              tma_im2col_4d_pair(sb + j * GEMM_ATOM_BYTES, &tmap_b, full_leader, c0 + j * 64, w0, h0, img,
                                 static_cast<uint16_t>(fs), static_cast<uint16_t>(fr));
          } else {
            ptx::mbar_expect_tx(&full_bar[stage], stage_bytes);
#pragma unroll
            for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
              ptx::tma_load_2d(sa + j * GEMM_ATOM_BYTES, &tmap_a, &full_bar[stage], o0 + j * 64, kb * GEMM_BLOCK_K);
            for (int j = 0; j < b_cols / 64; ++j)
              tma_im2col_4d(sb + j * GEMM_ATOM_BYTES, &tmap_b, &full_bar[stage], c0 + j * 64, w0, h0, img,
                            static_cast<uint16_t>(fs), static_cast<uint16_t>(fr));
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ------------------------------------------------- MMA issuer (leader CTA only for a pair)
    const uint32_t idesc = ptx::make_idesc_f16(p.fmt, rows_per_tile, block_n, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (WorkIter it(p, worker, workers); it.next();) {
      const int kb0 = it.kb0, kb1 = it.kb1;
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
        gemm_issue_kblock<PAIR, A_MN, B_MN>(sa, sa + GEMM_A_BYTES, d_tmem, idesc, kb == kb0, &empty_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (PAIR) ptx::umma_commit_pair(&tfull_bar[acc], 0b11);
      else ptx::umma_commit(&tfull_bar[acc]);
      if (it.last()) {  // wakes warps 0..3 for their share of the last drain (gemm_last_unit_helper)
        if (PAIR) ptx::umma_commit_pair(tlast_bar, 0b11);
        else ptx::umma_commit(tlast_bar);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;
    uint8_t* stage_buf = epi_smem + ew * 4096;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (WorkIter it(p, worker, workers); it.next();) {
      const int n_idx = it.tile % p.n_tiles;
      const int m_idx = it.tile / p.n_tiles;
      int col_base, col_limit;
      if (MODE == MODE_FPROP) {
        col_base = n_idx * block_n;
        col_limit = p.N;
      } else {
        const int tap = n_idx / cp.tiles_per_tap;
        col_base = tap * cp.C + (n_idx - tap * cp.tiles_per_tap) * block_n;
        col_limit = (tap + 1) * cp.C;  // columns past this tap's channels are padding of the tile
      }
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M + ew * 32;
      const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(ew * 32) << 16);
      const int chunks = it.last() ? main_chunks_of_last_unit(block_n) : (block_n >> 5);
      gemm_epilogue_tile<PAIR, EPI>(t_row, row0, col_base, col_limit, 0, chunks, p, &tmap_c, stage_buf, buf,
                                    &tempty_bar[acc], lane);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (epi_uses_tma(EPI) && lane == 0) ptx::tma_store_wait_read<0>();
  }

  if (warp < 4) {
    __syncwarp();
    const int total = p.m_tiles * p.n_tiles * p.splits;
    if (worker < total) {
      const int u = (total - worker + workers - 1) / workers - 1;  // this worker's last unit
      const int tile = (worker + u * workers) / p.splits;
      const int n_idx = tile % p.n_tiles, m_idx = tile / p.n_tiles;
      int col_base, col_limit;
      if (MODE == MODE_FPROP) {
        col_base = n_idx * block_n;
        col_limit = p.N;
      } else {
        const int tap = n_idx / cp.tiles_per_tap;
        col_base = tap * cp.C + (n_idx - tap * cp.tiles_per_tap) * block_n;
        col_limit = (tap + 1) * cp.C;
      }
      const int row0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M + warp * 32;
      gemm_last_unit_helper<PAIR, EPI>(tmem_base, u & 1, row0, col_base, col_limit, block_n, p, &tmap_c, smem, tlast_bar,
                                       warp, lane);
    }
  }

  gemm_teardown<PAIR>(tmem_base, warp);
}

}  // namespace lyco
