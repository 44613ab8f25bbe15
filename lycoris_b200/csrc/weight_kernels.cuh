// Weight-side (HBM-bound) kernels: merged-weight assembly W' = W + dW(factors) and the
// factor gradients from dW' = dY^T X.  One pass over the [N, K'] weight each; factors are tiny
// and live in shared memory / L1.  Element (n, k) of a convolution weight is addressed with
// k = c*kh*kw + i*kw + j, exactly how the reference flattens it.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/lyco_b200.h"
#include "pdl.cuh"

namespace lyco {

// ------------------------------------------------------------------ helpers --
__device__ __forceinline__ float rnd(float x, int dtype) {
  if (dtype == LYCO_BF16) return __bfloat162float(__float2bfloat16_rn(x));
  if (dtype == LYCO_F16) return __half2float(__float2half_rn(x));
  return x;
}
__device__ __forceinline__ float ld_f(const void* p, int dtype, int64_t i) {
  if (dtype == LYCO_F32) return __ldg(reinterpret_cast<const float*>(p) + i);
  if (dtype == LYCO_BF16)
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(p) + i));
  return __half2float(__ldg(reinterpret_cast<const __half*>(p) + i));
}
__device__ __forceinline__ float cvt16(uint16_t h, int dtype) {
  if (dtype == LYCO_BF16) return __uint_as_float(static_cast<uint32_t>(h) << 16);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
__device__ __forceinline__ uint16_t to16(float x, int dtype) {
  if (dtype == LYCO_BF16) {
    __nv_bfloat16 v = __float2bfloat16_rn(x);
    return *reinterpret_cast<uint16_t*>(&v);
  }
  __half v = __float2half_rn(x);
  return *reinterpret_cast<uint16_t*>(&v);
}

struct Chain {
  int pre_round, pre_dtype, w_dtype;
  float m_pre, m_post1, m_post2;
};
// raw factor product -> delta on the weight's grid (the reference's rounding points)
__device__ __forceinline__ float apply_chain(float raw, const Chain& c) {
  const int pd = c.pre_round ? c.pre_dtype : LYCO_F32;
  float d = rnd(raw, pd);
  d = rnd(d * c.m_pre, pd);
  d = rnd(d, c.w_dtype);
  d = rnd(d * c.m_post1, c.w_dtype);
  d = rnd(d * c.m_post2, c.w_dtype);
  return d;
}
__device__ __forceinline__ float merged(float w, float delta, int w_dtype) {
  return rnd(w + delta, w_dtype);
}

// Tile geometry shared by the low-rank (LoCon / DyLoRA / LoHa) kernels:
//   256 threads, tile = 32 rows x 256 columns; warp w owns rows 4w..4w+3, lane l owns the
//   8 consecutive columns 8l..8l+7 (one 16-byte vector of 16-bit weights).
constexpr int LR_ROWS = 32;
constexpr int LR_COLS = 256;
constexpr int LR_RC = 16;  // rank chunk held in shared memory

// stage [rows x rc] of `a` ([N, r] row-major) and [rc x cols] of `b` ([r, K'] row-major)
__device__ __forceinline__ void lr_stage_factors(const void* a, const void* b, int f_dtype, int N,
                                                 int K, int r, int n0, int k0, int r0, int rc,
                                                 float (*sa)[LR_RC], float (*sb)[LR_COLS],
                                                 int round_dtype, float b_mul) {
  for (int i = threadIdx.x; i < LR_ROWS * LR_RC; i += blockDim.x) {
    const int row = i / LR_RC, q = i % LR_RC;
    float v = 0.f;
    if (n0 + row < N && q < rc) v = rnd(ld_f(a, f_dtype, static_cast<int64_t>(n0 + row) * r + r0 + q), round_dtype);
    sa[row][q] = v;
  }
  for (int i = threadIdx.x; i < LR_RC * LR_COLS; i += blockDim.x) {
    const int q = i / LR_COLS, col = i % LR_COLS;
    float v = 0.f;
    if (q < rc && k0 + col < K) {
      v = ld_f(b, f_dtype, static_cast<int64_t>(r0 + q) * K + k0 + col);
      if (b_mul != 1.f) v = rnd(v * b_mul, f_dtype);  // dylora: down * (alpha/(b+1) * mult) in param dtype
      v = rnd(v, round_dtype);
    }
    sb[q][col] = v;
  }
}

// acc[4][8] += sa[rows][rc] * sb[rc][cols] for this thread's 4x8 patch
__device__ __forceinline__ void lr_accumulate(float (&acc)[4][8], const float (*sa)[LR_RC],
                                              const float (*sb)[LR_COLS], int rc, int warp,
                                              int lane) {
  for (int q = 0; q < rc; ++q) {
    const float4 b0 = *reinterpret_cast<const float4*>(&sb[q][8 * lane]);
    const float4 b1 = *reinterpret_cast<const float4*>(&sb[q][8 * lane + 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = sa[4 * warp + i][q];
      acc[i][0] = fmaf(a, b0.x, acc[i][0]);
      acc[i][1] = fmaf(a, b0.y, acc[i][1]);
      acc[i][2] = fmaf(a, b0.z, acc[i][2]);
      acc[i][3] = fmaf(a, b0.w, acc[i][3]);
      acc[i][4] = fmaf(a, b1.x, acc[i][4]);
      acc[i][5] = fmaf(a, b1.y, acc[i][5]);
      acc[i][6] = fmaf(a, b1.z, acc[i][6]);
      acc[i][7] = fmaf(a, b1.w, acc[i][7]);
    }
  }
}

// full product P[4][8] = a[n,:] . b[:,k] over all rank chunks
__device__ __forceinline__ void lr_product(float (&acc)[4][8], const void* a, const void* b,
                                           int f_dtype, int N, int K, int r, int n0, int k0,
                                           float (*sa)[LR_RC], float (*sb)[LR_COLS],
                                           int round_dtype, float b_mul) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int r0 = 0; r0 < r; r0 += LR_RC) {
    const int rc = min(LR_RC, r - r0);
    __syncthreads();
    lr_stage_factors(a, b, f_dtype, N, K, r, n0, k0, r0, rc, sa, sb, round_dtype, b_mul);
    __syncthreads();
    lr_accumulate(acc, sa, sb, rc, warp, lane);
  }
}

// ------------------------------------------------------------- merge kernels --
// LoCon / DyLoRA (LOHA = false) and LoHa (LOHA = true)
template <bool LOHA>
__global__ void __launch_bounds__(256) merge_lowrank_kernel(lyco_delta_desc_t d, const uint16_t* __restrict__ W,
                                                            uint16_t* __restrict__ Wout) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  __shared__ __align__(16) float sa[LR_ROWS][LR_RC];
  __shared__ __align__(16) float sb[LR_RC][LR_COLS];
  const int N = d.out_dim, K = d.in_dim, r = d.rank;
  const int k_tiles = (K + LR_COLS - 1) / LR_COLS;
  const int n0 = (blockIdx.x / k_tiles) * LR_ROWS;
  const int k0 = (blockIdx.x % k_tiles) * LR_COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Chain ch{d.pre_round, d.pre_dtype, d.w_dtype, d.m_pre, d.m_post1, d.m_post2};
  const int fround = d.pre_round ? d.pre_dtype : LYCO_F32;  // operands as the 16-bit matmul sees them

  float p1[4][8];
  lr_product(p1, d.f0, d.f1, d.f_dtype, N, K, r, n0, k0, sa, sb, fround, d.m_in);
  if (LOHA) {
    float p2[4][8];
    lr_product(p2, d.f2, d.f3, d.f_dtype, N, K, r, n0, k0, sa, sb, fround, 1.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) p1[i][j] = rnd(p1[i][j], fround) * rnd(p2[i][j], fround);
  }
  const int kc = k0 + 8 * lane;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * warp + i;
    if (n >= N || kc >= K) continue;
    const int64_t off = static_cast<int64_t>(n) * K + kc;
    if (kc + 8 <= K && (K % 8) == 0) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(W + off));
      const uint16_t* wh = reinterpret_cast<const uint16_t*>(&wv);
      uint4 ov;
      uint16_t* oh = reinterpret_cast<uint16_t*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        oh[j] = to16(merged(cvt16(wh[j], d.w_dtype), apply_chain(p1[i][j], ch), d.w_dtype), d.w_dtype);
      *reinterpret_cast<uint4*>(Wout + off) = ov;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (kc + j < K)
          Wout[off + j] = to16(merged(cvt16(W[off + j], d.w_dtype), apply_chain(p1[i][j], ch), d.w_dtype), d.w_dtype);
    }
  }
}

// Compile-time dtype specialisation for the one-pass merge kernels.  FD / WD = factor / weight dtype, PD = the
// product domain: PD_RUNTIME -> everything read from the descriptor (generic instantiation), PD_NONE -> no
// pre-rounding (fp32 product), otherwise pre_round with that dtype.  The kernel source is the SAME for every
// instantiation — the local constants below replace the descriptor fields, the compiler folds the dtype switches of
// rnd / ld_f / cvt16 / to16 — so a specialised kernel computes bit for bit what the generic one does.
constexpr int DT_RUNTIME = -1;
constexpr int PD_RUNTIME = -2;
constexpr int PD_NONE = -1;

// LoKr: dW[pu*vp+pv, u*vq+v] = w1[pu,u] * w2[pv,v]   (functional/lokr.py:11-20 via torch.kron)
template <int VEC, int FD = DT_RUNTIME, int WD = DT_RUNTIME, int PD = PD_RUNTIME>
__global__ void __launch_bounds__(256) merge_lokr_kernel(lyco_delta_desc_t d, const uint16_t* __restrict__ W,
                                                         uint16_t* __restrict__ Wout) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  const int f_dtype = FD == DT_RUNTIME ? d.f_dtype : FD;
  const int w_dtype = WD == DT_RUNTIME ? d.w_dtype : WD;
  const int pre_round = PD == PD_RUNTIME ? d.pre_round : (PD == PD_NONE ? 0 : 1);
  const int pre_dtype = PD == PD_RUNTIME ? d.pre_dtype : (PD == PD_NONE ? LYCO_F32 : PD);
  const int K = d.in_dim;
  const int64_t total = static_cast<int64_t>(d.out_dim) * K / VEC;
  const Chain ch{pre_round, pre_dtype, w_dtype, d.m_pre, d.m_post1, d.m_post2};
  const int fround = pre_round ? pre_dtype : LYCO_F32;
  const int kv = K / VEC;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / kv);
    const int k = static_cast<int>(idx % kv) * VEC;
    const int pu = n / d.vp, pv = n % d.vp;
    const int u = k / d.vq, v = k % d.vq;  // VEC divides vq -> the vector stays inside one w1 block
    const float a = rnd(ld_f(d.f0, f_dtype, pu * d.uq + u), fround);
    const int64_t off = static_cast<int64_t>(n) * K + k;
    if (VEC == 8) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(W + off));
      const uint16_t* wh = reinterpret_cast<const uint16_t*>(&wv);
      uint4 ov;
      uint16_t* oh = reinterpret_cast<uint16_t*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b = rnd(ld_f(d.f1, f_dtype, static_cast<int64_t>(pv) * d.vq + v + j), fround);
        oh[j] = to16(merged(cvt16(wh[j], w_dtype), apply_chain(a * b, ch), w_dtype), w_dtype);
      }
      *reinterpret_cast<uint4*>(Wout + off) = ov;
    } else {
      const float b = rnd(ld_f(d.f1, f_dtype, static_cast<int64_t>(pv) * d.vq + v), fround);
      Wout[off] = to16(merged(cvt16(W[off], w_dtype), apply_chain(a * b, ch), w_dtype), w_dtype);
    }
  }
}

// Row-blocked LoKr merge (round 2): one warp walks one weight row at a time, lanes stride over its 16-byte vectors.
// The row decomposition (pu, pv) — two integer divisions in the element-indexed kernel above — is done once per row
// per warp, the 8 values of w2 a vector needs arrive as two 16-byte loads (fp32 factors) or one (16-bit factors)
// instead of eight scalar loads, the w1 row sits in registers / L1, and multipliers equal to 1 skip their rounding
// step (bit-identical: rnd(x * 1) == x once x is on the grid).  Needs vq % 8 == 0 and 16-byte aligned arrays.
template <int FD, int WD, int PD>
__global__ void __launch_bounds__(256) merge_lokr_rows_kernel(lyco_delta_desc_t d, const uint16_t* __restrict__ W,
                                                              uint16_t* __restrict__ Wout) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  constexpr int f_dtype = FD, w_dtype = WD;
  constexpr int pre_round = PD == PD_NONE ? 0 : 1;
  constexpr int pre_dtype = PD == PD_NONE ? LYCO_F32 : PD;
  constexpr int fround = pre_round ? pre_dtype : LYCO_F32;
  const int K = d.in_dim, N = d.out_dim, vq = d.vq, vp = d.vp, uq = d.uq;
  const int kv = K >> 3;                     // 16-byte vectors per row
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const bool unit_pre = d.m_pre == 1.f, unit_p1 = d.m_post1 == 1.f, unit_p2 = d.m_post2 == 1.f;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    const int pu = n / vp, pv = n - pu * vp;
    const uint4* wrow = reinterpret_cast<const uint4*>(W + static_cast<int64_t>(n) * K);
    uint4* orow = reinterpret_cast<uint4*>(Wout + static_cast<int64_t>(n) * K);
    // two vectors per lane per trip: both weight loads are issued before either is consumed
    for (int i0 = lane; i0 < kv; i0 += 64) {
      const int i1 = i0 + 32;
      const bool has1 = i1 < kv;
      const uint4 wv0 = __ldg(wrow + i0);
      const uint4 wv1 = has1 ? __ldg(wrow + i1) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t == 1 && !has1) break;
        const int i = t ? i1 : i0;
        const uint4 wv = t ? wv1 : wv0;
        const int k = i << 3;
        const int u = k / vq, v = k - u * vq;
        const float a = rnd(ld_f(d.f0, f_dtype, pu * uq + u), fround);
        float b[8];
        const int64_t boff = static_cast<int64_t>(pv) * vq + v;
        if (f_dtype == LYCO_F32) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.f1) + boff));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.f1) + boff) + 1);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        } else {
          const uint4 bv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(d.f1) + boff));
          const uint16_t* bh = reinterpret_cast<const uint16_t*>(&bv);
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = cvt16(bh[j], f_dtype);
        }
        const uint16_t* wh = reinterpret_cast<const uint16_t*>(&wv);
        uint4 ov;
        uint16_t* oh = reinterpret_cast<uint16_t*>(&ov);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // the rounding chain of apply_chain(), with the identity steps folded away
          float dv = rnd(a * rnd(b[j], fround), fround);
          if (!unit_pre) dv = rnd(dv * d.m_pre, fround);
          dv = rnd(dv, w_dtype);
          if (!unit_p1) dv = rnd(dv * d.m_post1, w_dtype);
          if (!unit_p2) dv = rnd(dv * d.m_post2, w_dtype);
          oh[j] = to16(cvt16(wh[j], w_dtype) + dv, w_dtype);
        }
        orow[i] = ov;
      }
    }
  }
}

// RAW: the rank-r products were formed on the tensor cores (lyco_gemm with K = r) and arrive as 16-bit
// [N, K'] arrays: W' = rnd(W + chain(raw1 [* raw2])).  Used for LoCon / DyLoRA (one product) and LoHa (two).
// D16 != DT_RUNTIME: products, product domain and weights all in that 16-bit dtype (the autocast / all-bf16 regimes).
template <int D16 = DT_RUNTIME>
__global__ void __launch_bounds__(256) merge_raw_kernel(lyco_delta_desc_t d, const uint16_t* __restrict__ W,
                                                        uint16_t* __restrict__ Wout) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  const int w_dtype = D16 == DT_RUNTIME ? d.w_dtype : D16;
  const int pre_round = D16 == DT_RUNTIME ? d.pre_round : 1;
  const int pre_dtype = D16 == DT_RUNTIME ? d.pre_dtype : D16;
  const int pd = D16 == DT_RUNTIME ? d.f_dtype : D16;  // dtype of the raw products
  const int64_t total8 = static_cast<int64_t>(d.out_dim) * d.in_dim / 8;
  const Chain ch{pre_round, pre_dtype, w_dtype, d.m_pre, d.m_post1, d.m_post2};
  const uint4* r1 = reinterpret_cast<const uint4*>(d.f0);
  const uint4* r2 = reinterpret_cast<const uint4*>(d.f1);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(W) + i);
    const uint4 av = __ldg(r1 + i);
    uint4 bv = make_uint4(0, 0, 0, 0);
    if (r2) bv = __ldg(r2 + i);
    const uint16_t* wh = reinterpret_cast<const uint16_t*>(&wv);
    const uint16_t* ah = reinterpret_cast<const uint16_t*>(&av);
    const uint16_t* bh = reinterpret_cast<const uint16_t*>(&bv);
    uint4 ov;
    uint16_t* oh = reinterpret_cast<uint16_t*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float raw = cvt16(ah[j], pd);
      if (r2) raw *= cvt16(bh[j], pd);
      oh[j] = to16(merged(cvt16(wh[j], w_dtype), apply_chain(raw, ch), w_dtype), w_dtype);
    }
    reinterpret_cast<uint4*>(Wout)[i] = ov;
  }
}

// G = rnd16(gscale * dW [* P]) — operand of the skinny tensor-core gradient contractions
__global__ void __launch_bounds__(256) grad_prep_kernel(const float* __restrict__ dW, const uint16_t* __restrict__ P,
                                                        uint16_t* __restrict__ G, int64_t n8, float gscale, int dtype) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(dW) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(dW) + 2 * i + 1);
    float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 pv = make_uint4(0, 0, 0, 0);
    if (P) pv = __ldg(reinterpret_cast<const uint4*>(P) + i);
    const uint16_t* ph = reinterpret_cast<const uint16_t*>(&pv);
    uint4 ov;
    uint16_t* oh = reinterpret_cast<uint16_t*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = g[j] * gscale;
      if (P) v *= cvt16(ph[j], dtype);
      oh[j] = to16(v, dtype);
    }
    reinterpret_cast<uint4*>(G)[i] = ov;
  }
}

// (IA)^3: W' = W * (1 + w*mult) on output rows or input channels (ia3.py:91-102)
__global__ void __launch_bounds__(256) merge_ia3_kernel(lyco_delta_desc_t d, const uint16_t* __restrict__ W,
                                                        uint16_t* __restrict__ Wout) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  const int K = d.in_dim;
  const int64_t total = static_cast<int64_t>(d.out_dim) * K;
  const int cd = (d.f_dtype == LYCO_F32) ? LYCO_F32 : d.f_dtype;  // dtype the scale is computed in
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / K);
    const int k = static_cast<int>(idx % K);
    const int ch = d.on_input ? (k / d.ia3_group) : n;
    float s = rnd(ld_f(d.f0, d.f_dtype, ch) * d.m_post2, cd);
    s = rnd(s + 1.f, cd);
    const float w = cvt16(W[idx], d.w_dtype);
    // product is formed in the promoted dtype (fp32 when the scale is fp32), then cast to W's dtype
    Wout[idx] = to16(rnd(w * s, cd == LYCO_F32 ? LYCO_F32 : d.w_dtype), d.w_dtype);
  }
}

// -------------------------------------------------------------- grad kernels --
// Reduce this thread's G[4][8] patch against staged factors:
//   ga[n, q] += sum_k G[n,k] * sb[q][k]     (gradient of the [N, r] factor)
//   gb[q, k] += sum_n sa[n][q] * G[n,k]     (gradient of the [r, K'] factor)
__device__ __forceinline__ void lr_grad_tile(const float (&G)[4][8], const float (*sa)[LR_RC],
                                             const float (*sb)[LR_COLS], float* red /*[8][LR_COLS]*/,
                                             int rc, int r0, int r, int N, int K, int n0, int k0,
                                             float* ga, float* gb, float gb_scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = 0; q < rc; ++q) {
    const float4 b0 = *reinterpret_cast<const float4*>(&sb[q][8 * lane]);
    const float4 b1 = *reinterpret_cast<const float4*>(&sb[q][8 * lane + 4]);
    float colsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = G[i][0] * b0.x + G[i][1] * b0.y + G[i][2] * b0.z + G[i][3] * b0.w + G[i][4] * b1.x +
                G[i][5] * b1.y + G[i][6] * b1.z + G[i][7] * b1.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int n = n0 + 4 * warp + i;
      if (lane == 0 && n < N) atomicAdd(&ga[static_cast<int64_t>(n) * r + r0 + q], s);
      const float a = sa[4 * warp + i][q];
#pragma unroll
      for (int j = 0; j < 8; ++j) colsum[j] = fmaf(a, G[i][j], colsum[j]);
    }
    // cross-warp reduction of the column sums through shared memory
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp * LR_COLS + 8 * lane + j] = colsum[j];
    __syncthreads();
    {
      const int col = threadIdx.x;  // 256 threads <-> 256 columns
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * LR_COLS + col];
      if (k0 + col < K) atomicAdd(&gb[static_cast<int64_t>(r0 + q) * K + k0 + col], s * gb_scale);
    }
  }
}

template <bool LOHA>
__global__ void __launch_bounds__(256) grad_lowrank_kernel(lyco_delta_desc_t d, const float* __restrict__ dW,
                                                           float* g0, float* g1, float* g2, float* g3) {
  __shared__ __align__(16) float sa[LR_ROWS][LR_RC];
  __shared__ __align__(16) float sb[LR_RC][LR_COLS];
  __shared__ __align__(16) float red[8 * LR_COLS];
  const int N = d.out_dim, K = d.in_dim, r = d.rank;
  const int k_tiles = (K + LR_COLS - 1) / LR_COLS;
  const int n0 = (blockIdx.x / k_tiles) * LR_ROWS;
  const int k0 = (blockIdx.x % k_tiles) * LR_COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fround = d.pre_round ? d.pre_dtype : LYCO_F32;
  const float gscale = d.m_pre * d.m_post1 * d.m_post2;

  float G[4][8];
  const int kc = k0 + 8 * lane;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * warp + i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = 0.f;
      if (n < N && kc + j < K) g = __ldg(&dW[static_cast<int64_t>(n) * K + kc + j]) * gscale;
      G[i][j] = g;
    }
  }
  if (!LOHA) {
    for (int r0 = 0; r0 < r; r0 += LR_RC) {
      const int rc = min(LR_RC, r - r0);
      __syncthreads();
      lr_stage_factors(d.f0, d.f1, d.f_dtype, N, K, r, n0, k0, r0, rc, sa, sb, fround, d.m_in);
      __syncthreads();
      lr_grad_tile(G, sa, sb, red, rc, r0, r, N, K, n0, k0, g0, g1, d.m_in);
    }
  } else {
    // dP1 = G * P2, dP2 = G * P1   (functional/loha.py:18-30, recomputed instead of cached)
    float P1[4][8], P2[4][8];
    lr_product(P1, d.f0, d.f1, d.f_dtype, N, K, r, n0, k0, sa, sb, fround, 1.f);
    lr_product(P2, d.f2, d.f3, d.f_dtype, N, K, r, n0, k0, sa, sb, fround, 1.f);
    float G1[4][8], G2[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        G1[i][j] = G[i][j] * rnd(P2[i][j], fround);
        G2[i][j] = G[i][j] * rnd(P1[i][j], fround);
      }
    for (int r0 = 0; r0 < r; r0 += LR_RC) {
      const int rc = min(LR_RC, r - r0);
      __syncthreads();
      lr_stage_factors(d.f0, d.f1, d.f_dtype, N, K, r, n0, k0, r0, rc, sa, sb, fround, 1.f);
      __syncthreads();
      lr_grad_tile(G1, sa, sb, red, rc, r0, r, N, K, n0, k0, g0, g1, 1.f);
      __syncthreads();
      lr_stage_factors(d.f2, d.f3, d.f_dtype, N, K, r, n0, k0, r0, rc, sa, sb, fround, 1.f);
      __syncthreads();
      lr_grad_tile(G2, sa, sb, red, rc, r0, r, N, K, n0, k0, g2, g3, 1.f);
    }
  }
}

// LoKr: g_w1[pu,u] = sum_{pv,v} dW[pu*vp+pv, u*vq+v] * w2[pv,v]
//       g_w2[pv,v] = sum_{pu,u} dW[...]              * w1[pu,u]
// grid = (plane tiles, up): a CTA owns VEC*256 consecutive (pv,v) plane elements of ONE w1 row pu and walks
// the uq blocks of that row (short loop, 16-byte loads); g_w1 partials are staged in shared memory
// (one global atomic per entry per CTA), g_w2 takes `up` atomics per element (g_w2 is zero-filled first).
template <int VEC>
__global__ void __launch_bounds__(256) grad_lokr_kernel(lyco_delta_desc_t d, const float* __restrict__ dW,
                                                        float* g_w1, float* g_w2) {
  extern __shared__ float s_w1[];  // [uq]
  const int K = d.in_dim;
  const int pu = blockIdx.y;
  for (int i = threadIdx.x; i < d.uq; i += blockDim.x) s_w1[i] = 0.f;
  __syncthreads();
  const int fround = d.pre_round ? d.pre_dtype : LYCO_F32;
  const float gscale = d.m_pre * d.m_post1 * d.m_post2;
  const int64_t plane = static_cast<int64_t>(d.vp) * d.vq;
  const int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
  const bool live = e0 < plane;
  const int pv = live ? static_cast<int>(e0 / d.vq) : 0;
  const int v = live ? static_cast<int>(e0 % d.vq) : 0;  // VEC divides vq: the vector stays in one row
  float b[VEC], acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    b[i] = live ? rnd(ld_f(d.f1, d.f_dtype, e0 + i), fround) : 0.f;
    acc[i] = 0.f;
  }
  const int lane = threadIdx.x & 31;
  const float* row = dW + static_cast<int64_t>(pu * d.vp + pv) * K + v;
  // Groups of four w1 columns: the four (read-only) row loads of a group are issued before the first shuffle tree, so
  // a thread keeps 64 bytes in flight instead of 16.  Per-u arithmetic and its order are unchanged.
  constexpr int UG = 4;
  for (int u0 = 0; u0 < d.uq; u0 += UG) {
    float g[UG][VEC];
#pragma unroll
    for (int j = 0; j < UG; ++j) {
      const bool on = live && (u0 + j < d.uq);
      if (VEC == 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) t = __ldg(reinterpret_cast<const float4*>(row + static_cast<int64_t>(u0 + j) * d.vq));
        g[j][0] = t.x; g[j][1 % VEC] = t.y; g[j][2 % VEC] = t.z; g[j][3 % VEC] = t.w;
      } else {
        g[j][0] = on ? __ldg(row + static_cast<int64_t>(u0 + j) * d.vq) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < UG; ++j) {
      const int u = u0 + j;
      if (u >= d.uq) break;  // warp-uniform
      const float a = rnd(ld_f(d.f0, d.f_dtype, pu * d.uq + u), fround);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float gi = g[j][i] * gscale;
        acc[i] = fmaf(a, gi, acc[i]);
        s = fmaf(gi, b[i], s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) atomicAdd(&s_w1[u], s);
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomicAdd(&g_w2[e0 + i], acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.uq; i += blockDim.x) atomicAdd(&g_w1[pu * d.uq + i], s_w1[i]);
}

// (IA)^3: g_w[c] = mult * sum dW[n,k] * W[n,k] over the row (or the input-channel columns)
__global__ void __launch_bounds__(256) grad_ia3_kernel(lyco_delta_desc_t d, const float* __restrict__ dW,
                                                       const uint16_t* __restrict__ W, float* g_w) {
  const int K = d.in_dim, N = d.out_dim;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!d.on_input) {
    // one warp per output row
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    float s = 0.f;
    for (int k = lane; k < K; k += 32)
      s = fmaf(__ldg(&dW[static_cast<int64_t>(n) * K + k]), cvt16(W[static_cast<int64_t>(n) * K + k], d.w_dtype), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) g_w[n] = s * d.m_post2;
  } else {
    // thread per column k, CTA covers a 64-row strip -> atomics across strips
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_begin = blockIdx.y * 64;
    const int n_end = min(N, n_begin + 64);
    if (k >= K) return;
    float s = 0.f;
    for (int n = n_begin; n < n_end; ++n)
      s = fmaf(__ldg(&dW[static_cast<int64_t>(n) * K + k]), cvt16(W[static_cast<int64_t>(n) * K + k], d.w_dtype), s);
    atomicAdd(&g_w[k / d.ia3_group], s * d.m_post2);
  }
}

}  // namespace lyco
