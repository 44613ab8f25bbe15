// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias)      16-bit operands, fp32 accumulation in TMEM
//
// One CTA per SM, 256 threads:
//   warp 0      TMA producer  (one elected lane)      global -> smem ring (SWIZZLE_128B boxes)
//   warp 1      MMA issuer    (one elected lane)      tcgen05.mma 128 x BLOCK_N x 16, D in TMEM
//   warp 2      TMEM allocator / deallocator
//   warps 4..7  epilogue: tcgen05.ld -> registers -> (bias, cast) -> global
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two
// accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1), and the
// static persistent tile schedule.
//
// Either operand may be K-major (reduction index contiguous in memory) or MN-major (output index
// contiguous): the three contractions of a wrapped layer are then all served without transposes
//   forward  Y  = X  * W'^T        A = X  [M,K]  K-major    B = W' [N,K]  K-major
//   dgrad    dX = dY * W'          A = dY [M,N]  K-major    B = W' [N,K]  MN-major (as [K_out, N_red])
//   wgrad    dW'= dY^T * X         A = dY [M,N]  MN-major   B = X  [M,K]  MN-major, split over M
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx_sm100.cuh"

namespace lyco {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KiB
constexpr int GEMM_ATOM_BYTES = 64 * GEMM_BLOCK_K * 2;         // one MN-major 64x64 box, 8 KiB
constexpr int GEMM_RING_BYTES = 192 * 1024;

enum { EPI_STORE16 = 0, EPI_STORE_F32 = 1, EPI_ATOMIC_F32 = 2 };

struct GemmParams {
  void* C;
  const void* bias;
  int64_t ldc;
  int M, N, K;
  int m_tiles, n_tiles, splits, k_blocks;
  int fmt;         // operand / 16-bit output format: 0 = f16, 1 = bf16
  int bias_dtype;  // LYCO_BF16 / LYCO_F16 / LYCO_F32
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
  static constexpr int STAGES = (GEMM_RING_BYTES / STAGE_BYTES) > 8 ? 8 : (GEMM_RING_BYTES / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float load_scalar(const void* p, int dtype, int64_t i) {
  if (dtype == 2) return reinterpret_cast<const float*>(p)[i];
  if (dtype == 0) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Epilogue for one 32-column chunk of one accumulator row held in registers (fp32 bit patterns).
template <int EPI>
__device__ __forceinline__ void store_chunk(const uint32_t (&r)[32], int row, int col0, const GemmParams& p,
                                            int col_limit = -1) {
  const int N = col_limit < 0 ? p.N : col_limit;  // first column this tile must not write
  if (row >= p.M || col0 >= N) return;
  const bool full = (col0 + 32 <= N);
  if (EPI == EPI_STORE16) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias != nullptr) {
      if (full && p.bias_dtype != 2 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
        // 32 x 16-bit bias values = four 16-byte loads (every lane reads the same addresses: L1 broadcast)
        const uint4* bp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.bias) + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 bv = __ldg(bp + q);
          const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float lo, hi;
            if (p.bias_dtype == 0) {
              lo = __uint_as_float(w[t] << 16);
              hi = __uint_as_float(w[t] & 0xFFFF0000u);
            } else {
              const __half2 h = *reinterpret_cast<const __half2*>(&w[t]);
              lo = __low2float(h);
              hi = __high2float(h);
            }
            v[8 * q + 2 * t] += lo;
            v[8 * q + 2 * t + 1] += hi;
          }
        }
      } else if (full) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += load_scalar(p.bias, p.bias_dtype, col0 + j);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) v[j] += load_scalar(p.bias, p.bias_dtype, col0 + j);
      }
    }
    uint16_t* crow = reinterpret_cast<uint16_t*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack16(v[8 * q + 0], v[8 * q + 1], p.fmt);
        o.y = pack16(v[8 * q + 2], v[8 * q + 3], p.fmt);
        o.z = pack16(v[8 * q + 4], v[8 * q + 5], p.fmt);
        o.w = pack16(v[8 * q + 6], v[8 * q + 7], p.fmt);
        reinterpret_cast<uint4*>(crow)[q] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) crow[j] = static_cast<uint16_t>(pack16(v[j], 0.f, p.fmt) & 0xFFFF);
    }
  } else {
    float* crow = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
    if (full) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (EPI == EPI_STORE_F32) {
          reinterpret_cast<float4*>(crow)[q] =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                          __uint_as_float(r[4 * q + 3]));
        } else {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + 4 * q),
                       "f"(__uint_as_float(r[4 * q])), "f"(__uint_as_float(r[4 * q + 1])),
                       "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3]))
                       : "memory");
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) {
          if (EPI == EPI_STORE_F32) crow[j] = __uint_as_float(r[j]);
          else atomicAdd(crow + j, __uint_as_float(r[j]));
        }
    }
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;

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
      ptx::mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
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
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + GEMM_A_BYTES;
        if (!A_MN) {
          ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_idx * GEMM_BLOCK_M);
        } else {
#pragma unroll
          for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
            ptx::tma_load_2d(sa + j * GEMM_ATOM_BYTES, &tmap_a, &full_bar[stage],
                             m_idx * GEMM_BLOCK_M + j * 64, kb * GEMM_BLOCK_K);
        }
        if (!B_MN) {
          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n_idx * BLOCK_N);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            ptx::tma_load_2d(sb + j * GEMM_ATOM_BYTES, &tmap_b, &full_bar[stage],
                             n_idx * BLOCK_N + j * 64, kb * GEMM_BLOCK_K);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
          // K-major : advance 16 elements (32 B) inside the 128 B swizzle row
          // MN-major: advance 16 k-rows of 128 B
          const uint64_t da = A_MN ? ptx::make_smem_desc(sa + kk * 2048, GEMM_ATOM_BYTES, 1024)
                                   : ptx::make_smem_desc(sa + kk * 32, 16, 1024);
          const uint64_t db = B_MN ? ptx::make_smem_desc(sb + kk * 2048, GEMM_ATOM_BYTES, 1024)
                                   : ptx::make_smem_desc(sb + kk * 32, 16, 1024);
          ptx::umma_f16(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(&tfull_bar[acc]);  // accumulator complete
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew + 32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int tile = w / p.splits;
      const int n_idx = tile % p.n_tiles;
      const int m_idx = tile / p.n_tiles;
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
          // all of this warp's TMEM reads are done: hand the accumulator back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        }
        store_chunk<EPI>(r, row, n_idx * BLOCK_N + c * 32, p);
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
