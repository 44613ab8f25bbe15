// LoHa: both rank-r factor products of a weight tile on the tensor cores, Hadamard product + rounding chain + `W +` in
// the epilogue — the tile of dW = (w1a·w1b) ⊙ (w2a·w2b) is assembled on the fly and never written to HBM.
//
//   MODE 0 (merge, forward):     W'[n,k] = rnd(W[n,k] + chain(rnd(P1[n,k]) * rnd(P2[n,k])))          read W, write W'
//   MODE 1 (gradient operands):  G1 = rnd(g * dW'[n,k] * rnd(P2)),  G2 = rnd(g * dW'[n,k] * rnd(P1))  read dW' (fp32),
//                                                                                                    write G1, G2
// with P1 = w1a [N,r] · w1b [r,K'], P2 = w2a · w2b (reference lycoris/functional/loha.py:10-30: HadaWeight forward, and
// the re-products of its backward).  Round 1 formed P1 and P2 with two lyco_gemm launches each writing an [N, K'] 16-bit
// array, then merge_raw / grad_prep read them back: 12 (forward) + 20 (backward) bytes of HBM traffic per weight
// element against 4 + 8 here.
//
// One CTA per 128 x 128 tile, 128 threads: thread 0 stages the four factor tiles with TMA (the rank r <= 64 is ONE
// 64-wide k-block, zero-filled past r by the tensor maps) and issues the two tcgen05.mma chains into two 128-column
// TMEM accumulators; the four warps then drain their 32 TMEM lanes (= weight rows) chunk by chunk.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_sm100.cuh"
#include "weight_kernels.cuh"

namespace lyco {

constexpr int HADA_BM = 128;
constexpr int HADA_BN = 128;
constexpr int HADA_THREADS = 128;
constexpr int HADA_A_BYTES = HADA_BM * 64 * 2;               // [128 rows][64 k] 16-bit, 128-byte swizzled rows
constexpr int HADA_B_BYTES = (HADA_BN / 64) * GEMM_ATOM_BYTES;  // two MN-major 64 x 64 boxes
constexpr int HADA_SMEM_BYTES = 2 * (HADA_A_BYTES + HADA_B_BYTES) + 1024 /*align*/ + 64 /*barriers*/;

struct HadaParams {
  const void* W;     // MODE 0: base weight (16-bit)      MODE 1: dW' (fp32)
  void* out0;        // MODE 0: W'                         MODE 1: G1
  void* out1;        // MODE 1: G2
  int N, K;          // weight is [N, K] row-major
  int rank;          // r <= 64
  int fmt;           // 1 = bf16, 0 = f16 (factors, products, 16-bit outputs)
  int w_dtype;       // MODE 0: dtype of W / W'
  float m_pre, m_post1, m_post2;  // MODE 0: rounding-chain multipliers (product domain = the 16-bit factor dtype)
  float gscale;      // MODE 1
};

template <int MODE>
__global__ void __launch_bounds__(HADA_THREADS)
hada_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_b1,
                  const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b2,
                  const HadaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sa1 = smem;
  uint8_t* sb1 = sa1 + HADA_A_BYTES;
  uint8_t* sa2 = sb1 + HADA_B_BYTES;
  uint8_t* sb2 = sa2 + HADA_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sb2 + HADA_B_BYTES);
  uint64_t* mma_bar = full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.K + HADA_BN - 1) / HADA_BN;
  const int m0 = (blockIdx.x / n_tiles) * HADA_BM;
  const int n0 = (blockIdx.x % n_tiles) * HADA_BN;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_a1);
    ptx::prefetch_tmap(&tmap_b1);
    ptx::prefetch_tmap(&tmap_a2);
    ptx::prefetch_tmap(&tmap_b2);
    ptx::mbar_init(full_bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 2 * HADA_BN);  // two accumulators of 128 columns
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(full_bar, 2 * (HADA_A_BYTES + HADA_B_BYTES));
    ptx::tma_load_2d(sa1, &tmap_a1, full_bar, 0, m0);
    ptx::tma_load_2d(sa2, &tmap_a2, full_bar, 0, m0);
#pragma unroll
    for (int j = 0; j < HADA_BN / 64; ++j) {
      ptx::tma_load_2d(sb1 + j * GEMM_ATOM_BYTES, &tmap_b1, full_bar, n0 + j * 64, 0);
      ptx::tma_load_2d(sb2 + j * GEMM_ATOM_BYTES, &tmap_b2, full_bar, n0 + j * 64, 0);
    }
    ptx::mbar_wait(full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_f16(p.fmt, HADA_BM, HADA_BN, 0, 1);
    const int steps = (p.rank + GEMM_UMMA_K - 1) / GEMM_UMMA_K;  // k-steps that hold real factor columns
    for (int pair = 0; pair < 2; ++pair) {
      const uint32_t a = ptx::smem_u32(pair ? sa2 : sa1), b = ptx::smem_u32(pair ? sb2 : sb1);
      const uint32_t d = tmem_base + pair * HADA_BN;
      for (int kk = 0; kk < steps; ++kk) {
        const uint64_t da = ptx::make_smem_desc(a + kk * 32, 16, 1024);                  // K-major A
        const uint64_t db = ptx::make_smem_desc(b + kk * 2048, GEMM_ATOM_BYTES, 1024);   // MN-major B
        ptx::umma_f16(d, da, db, idesc, kk > 0 ? 1u : 0u);
      }
    }
    ptx::umma_commit(mma_bar);
  }
  __syncwarp();
  ptx::mbar_wait(mma_bar, 0);
  ptx::tc_fence_after();

  // ------------------------------------------------------------------ epilogue: this thread owns weight row `row`
  const int row = m0 + warp * 32 + lane;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const int pd = p.fmt == 1 ? LYCO_BF16 : LYCO_F16;  // product dtype
  const Chain ch{1, pd, p.w_dtype, p.m_pre, p.m_post1, p.m_post2};
#pragma unroll 1
  for (int c = 0; c < HADA_BN / 32; ++c) {
    uint32_t r1[32], r2[32];
    ptx::tmem_ld_32x32(t_row + c * 32, r1);
    ptx::tmem_ld_32x32(t_row + HADA_BN + c * 32, r2);
    const int col0 = n0 + c * 32;
    if (row >= p.N || col0 >= p.K) continue;
    const int64_t off = static_cast<int64_t>(row) * p.K + col0;
    const int ncols = min(32, p.K - col0);
    if (MODE == 0) {
      const uint16_t* w = reinterpret_cast<const uint16_t*>(p.W) + off;
      uint16_t* o = reinterpret_cast<uint16_t*>(p.out0) + off;
      if (ncols == 32 && (p.K & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w) + q);
          const uint16_t* wh = reinterpret_cast<const uint16_t*>(&wv);
          uint4 ov;
          uint16_t* oh = reinterpret_cast<uint16_t*>(&ov);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float raw = rnd(__uint_as_float(r1[8 * q + j]), pd) * rnd(__uint_as_float(r2[8 * q + j]), pd);
            oh[j] = to16(merged(cvt16(wh[j], p.w_dtype), apply_chain(raw, ch), p.w_dtype), p.w_dtype);
          }
          reinterpret_cast<uint4*>(o)[q] = ov;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {  // fully unrolled with a predicate: r1 / r2 stay in registers
          if (j < ncols) {
            const float raw = rnd(__uint_as_float(r1[j]), pd) * rnd(__uint_as_float(r2[j]), pd);
            o[j] = to16(merged(cvt16(w[j], p.w_dtype), apply_chain(raw, ch), p.w_dtype), p.w_dtype);
          }
        }
      }
    } else {
      const float* g = reinterpret_cast<const float*>(p.W) + off;
      uint16_t* o1 = reinterpret_cast<uint16_t*>(p.out0) + off;
      uint16_t* o2 = reinterpret_cast<uint16_t*>(p.out1) + off;
      if (ncols == 32 && (p.K & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + 2 * q);
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(g) + 2 * q + 1);
          const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          uint4 v1, v2;
          uint16_t* h1 = reinterpret_cast<uint16_t*>(&v1);
          uint16_t* h2 = reinterpret_cast<uint16_t*>(&v2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float s = gv[j] * p.gscale;
            h1[j] = to16(s * rnd(__uint_as_float(r2[8 * q + j]), pd), pd);
            h2[j] = to16(s * rnd(__uint_as_float(r1[8 * q + j]), pd), pd);
          }
          reinterpret_cast<uint4*>(o1)[q] = v1;
          reinterpret_cast<uint4*>(o2)[q] = v2;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < ncols) {
            const float s = g[j] * p.gscale;
            o1[j] = to16(s * rnd(__uint_as_float(r2[j]), pd), pd);
            o2[j] = to16(s * rnd(__uint_as_float(r1[j]), pd), pd);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * HADA_BN);
  }
}

}  // namespace lyco
