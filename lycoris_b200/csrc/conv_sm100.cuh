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

template <int BLOCK_N, int MODE, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const ConvParams cp) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr bool A_MN = (MODE == MODE_WGRAD);
  constexpr bool B_MN = (MODE == MODE_WGRAD);
  const GemmParams& p = cp.g;

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

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int split = w % p.splits;
      const int tile = w / p.splits;
      const int n_idx = tile % p.n_tiles;
      const int m_idx = tile / p.n_tiles;
      const int kb0 = static_cast<int>(static_cast<int64_t>(split) * p.k_blocks / p.splits);
      const int kb1 = static_cast<int>(static_cast<int64_t>(split + 1) * p.k_blocks / p.splits);
      if (MODE == MODE_FPROP) {
        // base output pixel of this M tile -> (image, row, col) -> input-space pixel coordinate
        const int m0 = m_idx * GEMM_BLOCK_M;
        const int img = m0 / cp.PQ, rem = m0 - img * cp.PQ;
        const int py = rem / cp.Q, px = rem - py * cp.Q;
        const int w0 = px * cp.stride + cp.low_w, h0 = py * cp.stride + cp.low_h;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tap = kb / cp.CB, cb = kb - tap * cp.CB;
          const int fr = tap / cp.S, fs = tap - fr * cp.S;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
          tma_im2col_4d(sa, &tmap_a, &full_bar[stage], cb * 64, w0, h0, img, static_cast<uint16_t>(fs),
                        static_cast<uint16_t>(fr));
          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n_idx * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      } else {
        const int tap = n_idx / cp.tiles_per_tap;
        const int c0 = (n_idx - tap * cp.tiles_per_tap) * BLOCK_N;
        const int fr = tap / cp.S, fs = tap - fr * cp.S;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int m0 = kb * GEMM_BLOCK_K;  // first pixel of this reduction block
          const int img = m0 / cp.PQ, rem = m0 - img * cp.PQ;
          const int py = rem / cp.Q, px = rem - py * cp.Q;
          const int w0 = px * cp.stride + cp.low_w, h0 = py * cp.stride + cp.low_h;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
#pragma unroll
          for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
            ptx::tma_load_2d(sa + j * GEMM_ATOM_BYTES, &tmap_a, &full_bar[stage], m_idx * GEMM_BLOCK_M + j * 64,
                             kb * GEMM_BLOCK_K);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_im2col_4d(sb + j * GEMM_ATOM_BYTES, &tmap_b, &full_bar[stage], c0 + j * 64, w0, h0, img,
                          static_cast<uint16_t>(fs), static_cast<uint16_t>(fr));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // -------------------------------------------------------------- MMA issuer
    const uint32_t idesc = ptx::make_idesc_f16(p.fmt, GEMM_BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
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
          ptx::umma_f16(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(&tfull_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int tile = w / p.splits;
      const int n_idx = tile % p.n_tiles;
      const int m_idx = tile / p.n_tiles;
      int col_base, col_limit;
      if (MODE == MODE_FPROP) {
        col_base = n_idx * BLOCK_N;
        col_limit = p.N;
      } else {
        const int tap = n_idx / cp.tiles_per_tap;
        col_base = tap * cp.C + (n_idx - tap * cp.tiles_per_tap) * BLOCK_N;
        col_limit = (tap + 1) * cp.C;  // columns past this tap's channels are padding of the tile
      }
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m_idx * GEMM_BLOCK_M + ew * 32 + lane;
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_row + c * 32, r);
        ptx::tmem_ld_wait();
        if (c == BLOCK_N / 32 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        }
        store_chunk<EPI>(r, row, col_base + c * 32, p, col_limit);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace lyco
