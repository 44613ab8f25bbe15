// Weight-side kernels outside the training step's inner triple: the DoRA magnitude rescale fused around the merged
// weight (reference lycoris/modules/locon.py:239-260 apply_weight_decompose and its copies in loha.py / lokr.py),
// and the standalone delta weight dW (+ its squared Frobenius norm) that merge_to / onfly_merge / apply_max_norm need
// (reference lycoris/modules/base.py:326-374, lokr.py:383-397, 442-466).  All HBM-bound, one or two passes over [N, K'].
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "weight_kernels.cuh"
#include "lokr_struct_kernels.cuh"  // unpack8 / pack8

namespace lyco {

// group of element (n, k): the output row (on_out) or the input channel k / taps (weights are [N, C*taps])
__device__ __forceinline__ int dora_group(int n, int k, int on_out, int taps) { return on_out ? n : k / taps; }

// acc[group] += sum a*b over the group's elements; A is 16-bit (the merged weight, MODE 0: b = a) or fp32 dW'' with
// B = the 16-bit merged weight (MODE 1).  CTA = 64 rows x 256 columns strip; a thread owns one column of the strip.
template <int MODE>
__global__ void __launch_bounds__(256) dora_reduce_kernel(const void* __restrict__ A, const uint16_t* __restrict__ Wm,
                                                          float* __restrict__ acc, int N, int K, int on_out, int taps,
                                                          int w_dtype) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int n0 = blockIdx.y * 64;
  const int n1 = min(N, n0 + 64);
  const int lane = threadIdx.x & 31;
  if (on_out) {
    // rows: every thread of the CTA contributes its column to 64 row sums -> warp reduce, one atomic per warp per row
    for (int n = n0; n < n1; ++n) {
      float v = 0.f;
      if (k < K) {
        const int64_t off = static_cast<int64_t>(n) * K + k;
        const float w = cvt16(Wm[off], w_dtype);
        const float a = MODE == 0 ? w : __ldg(reinterpret_cast<const float*>(A) + off);
        v = a * w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(&acc[n], v);
    }
  } else {
    if (k >= K) return;
    float s = 0.f;
    for (int n = n0; n < n1; ++n) {
      const int64_t off = static_cast<int64_t>(n) * K + k;
      const float w = cvt16(Wm[off], w_dtype);
      const float a = MODE == 0 ? w : __ldg(reinterpret_cast<const float*>(A) + off);
      s = fmaf(a, w, s);
    }
    atomicAdd(&acc[k / taps], s);
  }
}

// scale of a group from its squared norm: s = g / (sqrt(sumsq) + eps), then s <- mult * (s - 1) + 1 when mult != 1.
// `sdt` is the dtype the reference does this arithmetic in (dora_scale's: fp32 under autocast; a bf16 adapter rounds the
// norm, the sum with eps and the quotient to bf16 — a per-group relative difference of up to 2^-8 if ignored).
__device__ __forceinline__ float dora_scale_of(float sumsq, float g, float mult, float eps, int sdt) {
  const float n = rnd(sqrtf(sumsq), sdt);
  float s = rnd(g / rnd(n + eps, sdt), sdt);
  if (mult != 1.f) s = rnd(rnd(mult * rnd(s - 1.f, sdt), sdt) + 1.f, sdt);
  return s;
}

// forward: W''[n,k] = rnd16(float(Wm[n,k]) * s[group])
__global__ void __launch_bounds__(256) dora_apply_fwd_kernel(const uint16_t* __restrict__ Wm, uint16_t* __restrict__ Wout,
                                                             const float* __restrict__ sumsq, const float* __restrict__ g,
                                                             int N, int K, int on_out, int taps, float mult, float eps,
                                                             int w_dtype, int sdt) {
  const int64_t total = static_cast<int64_t>(N) * K;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / K), k = static_cast<int>(idx % K);
    const int grp = dora_group(n, k, on_out, taps);
    const float s = dora_scale_of(__ldg(sumsq + grp), __ldg(g + grp), mult, eps, sdt);
    Wout[idx] = to16(rnd(cvt16(Wm[idx], w_dtype) * s, sdt), w_dtype);
  }
}

// backward through W'' = Wm * s(Wm):  with t[grp] = sum dW''*Wm, n = sqrt(sumsq), ne = n + eps
//   d g[grp]   = t * mult / ne
//   dWm[n,k]   = s * dW''[n,k] - (mult * g * t / (ne^2 * n)) * Wm[n,k]          (in place over dW'')
__global__ void __launch_bounds__(256) dora_apply_bwd_kernel(float* __restrict__ dW, const uint16_t* __restrict__ Wm,
                                                             const float* __restrict__ sumsq, const float* __restrict__ g,
                                                             const float* __restrict__ t, float* __restrict__ dg, int N,
                                                             int K, int on_out, int taps, float mult, float eps,
                                                             int w_dtype, int groups, int sdt) {
  const int64_t total = static_cast<int64_t>(N) * K;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (dg != nullptr && tid < groups) {
    const float n = sqrtf(__ldg(sumsq + tid));
    dg[tid] = __ldg(t + tid) * mult / (n + eps);
  }
  for (int64_t idx = tid; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n_ = static_cast<int>(idx / K), k = static_cast<int>(idx % K);
    const int grp = dora_group(n_, k, on_out, taps);
    const float ss = __ldg(sumsq + grp), gg = __ldg(g + grp), tt = __ldg(t + grp);
    const float nrm = sqrtf(ss), ne = nrm + eps;
    const float s = dora_scale_of(ss, gg, mult, eps, sdt);
    const float b = nrm > 0.f ? mult * gg * tt / (ne * ne * nrm) : 0.f;
    dW[idx] = s * dW[idx] - b * cvt16(Wm[idx], w_dtype);
  }
}

// ---- vector variants (K % 8 == 0, 16-byte aligned arrays): one warp walks one weight row, lanes stride over its
// 16-byte vectors — the (n, k) decomposition costs nothing, loads and stores are 16 bytes wide, row groups need no
// atomics at all.  The scalar kernels above remain for ragged shapes.

// row groups (on_out = 1): acc[n] = sum_k a*b, written directly (no memset, no atomics)
template <int MODE>
__global__ void __launch_bounds__(256) dora_reduce_rows_vec_kernel(const float* __restrict__ A, const uint16_t* __restrict__ Wm,
                                                                   float* __restrict__ acc, int N, int K, int w_dtype,
                                                                   float scale = 1.f) {
  const int kv = K >> 3, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int fmt = w_dtype == LYCO_BF16 ? 1 : 0;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    const uint4* wrow = reinterpret_cast<const uint4*>(Wm + static_cast<int64_t>(n) * K);
    const float4* arow = MODE == 1 ? reinterpret_cast<const float4*>(A + static_cast<int64_t>(n) * K) : nullptr;
    float s = 0.f;
    for (int i = lane; i < kv; i += 32) {
      float w[8];
      unpack8(__ldg(wrow + i), w, fmt);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(w[j], w[j], s);
      } else {
        const float4 a0 = __ldg(arow + 2 * i), a1 = __ldg(arow + 2 * i + 1);
        s = fmaf(a0.x, w[0], s); s = fmaf(a0.y, w[1], s); s = fmaf(a0.z, w[2], s); s = fmaf(a0.w, w[3], s);
        s = fmaf(a1.x, w[4], s); s = fmaf(a1.y, w[5], s); s = fmaf(a1.z, w[6], s); s = fmaf(a1.w, w[7], s);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) acc[n] = s * scale;
  }
}

// column groups (on_out = 0): a thread owns one 8-column vector over a strip of 32 rows; acc must be zero-filled
template <int MODE>
__global__ void __launch_bounds__(256) dora_reduce_cols_vec_kernel(const float* __restrict__ A, const uint16_t* __restrict__ Wm,
                                                                   float* __restrict__ acc, int N, int K, int taps,
                                                                   int w_dtype, float scale = 1.f) {
  const int kv = K >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kv) return;
  const int n0 = blockIdx.y * 32, n1 = min(N, n0 + 32);
  const int fmt = w_dtype == LYCO_BF16 ? 1 : 0;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = n0; n < n1; ++n) {
    float w[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(Wm + static_cast<int64_t>(n) * K) + i), w, fmt);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = fmaf(w[j], w[j], s[j]);
    } else {
      const float4* ar = reinterpret_cast<const float4*>(A + static_cast<int64_t>(n) * K) + 2 * i;
      const float4 a0 = __ldg(ar), a1 = __ldg(ar + 1);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = fmaf(a[j], w[j], s[j]);
    }
  }
  // neighbouring columns of one input channel (taps > 1) are folded before the atomic
  int g_prev = (8 * i) / taps;
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gj = (8 * i + j) / taps;
    if (gj != g_prev) {
      atomicAdd(&acc[g_prev], run * scale);
      run = 0.f;
      g_prev = gj;
    }
    run += s[j];
  }
  atomicAdd(&acc[g_prev], run * scale);
}

// (IA)^3 merge, vectorised: W'[n, k..k+7] = rnd_w(rnd_cd(W * (1 + w[group] * mult)))  (ia3.py:91-102), warp per row
__global__ void __launch_bounds__(256) merge_ia3_vec_kernel(lyco_delta_desc_t d, const uint16_t* __restrict__ W,
                                                            uint16_t* __restrict__ Wout) {
  const int K = d.in_dim, N = d.out_dim;
  const int kv = K >> 3, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int cd = d.f_dtype;  // dtype the scale is computed in (fp32: no rounding)
  const int pd = cd == LYCO_F32 ? LYCO_F32 : d.w_dtype;
  const int fmt = d.w_dtype == LYCO_BF16 ? 1 : 0;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    const uint4* wrow = reinterpret_cast<const uint4*>(W + static_cast<int64_t>(n) * K);
    uint4* orow = reinterpret_cast<uint4*>(Wout + static_cast<int64_t>(n) * K);
    float s_row = 0.f;
    if (!d.on_input) s_row = rnd(rnd(ld_f(d.f0, d.f_dtype, n) * d.m_post2, cd) + 1.f, cd);
    for (int i = lane; i < kv; i += 32) {
      float w[8], o[8];
      unpack8(__ldg(wrow + i), w, fmt);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float sc = s_row;
        if (d.on_input) sc = rnd(rnd(ld_f(d.f0, d.f_dtype, (8 * i + j) / d.ia3_group) * d.m_post2, cd) + 1.f, cd);
        o[j] = rnd(w[j] * sc, pd);
      }
      orow[i] = pack8(o, fmt);
    }
  }
}

// forward apply, vectorised: W''[n, k..k+7] = rnd16(rnd_s(Wm * s[group]))
__global__ void __launch_bounds__(256) dora_apply_fwd_vec_kernel(const uint16_t* __restrict__ Wm, uint16_t* __restrict__ Wout,
                                                                 const float* __restrict__ sumsq, const float* __restrict__ g,
                                                                 int N, int K, int on_out, int taps, float mult, float eps,
                                                                 int w_dtype, int sdt) {
  const int kv = K >> 3, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int fmt = w_dtype == LYCO_BF16 ? 1 : 0;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    const uint4* wrow = reinterpret_cast<const uint4*>(Wm + static_cast<int64_t>(n) * K);
    uint4* orow = reinterpret_cast<uint4*>(Wout + static_cast<int64_t>(n) * K);
    const float s_row = on_out ? dora_scale_of(__ldg(sumsq + n), __ldg(g + n), mult, eps, sdt) : 0.f;
    for (int i = lane; i < kv; i += 32) {
      float w[8], o[8];
      unpack8(__ldg(wrow + i), w, fmt);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float sc = s_row;
        if (!on_out) {
          const int grp = (8 * i + j) / taps;
          sc = dora_scale_of(__ldg(sumsq + grp), __ldg(g + grp), mult, eps, sdt);
        }
        o[j] = rnd(w[j] * sc, sdt);
      }
      orow[i] = pack8(o, fmt);
    }
  }
}

// backward apply, vectorised and in place over the fp32 dW; g_scale written by the first `groups` threads
__global__ void __launch_bounds__(256) dora_apply_bwd_vec_kernel(float* __restrict__ dW, const uint16_t* __restrict__ Wm,
                                                                 const float* __restrict__ sumsq, const float* __restrict__ g,
                                                                 const float* __restrict__ t, float* __restrict__ dg, int N,
                                                                 int K, int on_out, int taps, float mult, float eps,
                                                                 int w_dtype, int groups, int sdt) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int q = tid; dg != nullptr && q < groups; q += gridDim.x * blockDim.x)
    dg[q] = __ldg(t + q) * mult / (sqrtf(__ldg(sumsq + q)) + eps);
  const int kv = K >> 3, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int fmt = w_dtype == LYCO_BF16 ? 1 : 0;
  for (int n = tid >> 5; n < N; n += warps) {
    const uint4* wrow = reinterpret_cast<const uint4*>(Wm + static_cast<int64_t>(n) * K);
    float4* drow = reinterpret_cast<float4*>(dW + static_cast<int64_t>(n) * K);
    float a_row = 0.f, b_row = 0.f;
    if (on_out) {
      const float ss = __ldg(sumsq + n), gg = __ldg(g + n), tt = __ldg(t + n);
      const float nrm = sqrtf(ss), ne = nrm + eps;
      a_row = dora_scale_of(ss, gg, mult, eps, sdt);
      b_row = nrm > 0.f ? mult * gg * tt / (ne * ne * nrm) : 0.f;
    }
    for (int i = lane; i < kv; i += 32) {
      float w[8];
      unpack8(__ldg(wrow + i), w, fmt);
      float4 d0 = drow[2 * i], d1 = drow[2 * i + 1];
      float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = a_row, b = b_row;
        if (!on_out) {
          const int grp = (8 * i + j) / taps;
          const float ss = __ldg(sumsq + grp), gg = __ldg(g + grp), tt = __ldg(t + grp);
          const float nrm = sqrtf(ss), ne = nrm + eps;
          a = dora_scale_of(ss, gg, mult, eps, sdt);
          b = nrm > 0.f ? mult * gg * tt / (ne * ne * nrm) : 0.f;
        }
        d[j] = a * d[j] - b * w[j];
      }
      drow[2 * i] = make_float4(d[0], d[1], d[2], d[3]);
      drow[2 * i + 1] = make_float4(d[4], d[5], d[6], d[7]);
    }
  }
}

// ------------------------------------------------------------------------------------------ standalone delta weight
// dW[n,k] = chain(raw(n,k)) written in out_dtype (fp32 output with an fp32 "weight" dtype reproduces the reference's
// get_diff_weight arithmetic: products and scales in the parameter dtype, no 16-bit rounding points), optional
// sum of squares into *norm_sq.  Cold path (merge / max-norm): one thread per element, factor reads from L1/L2.
__global__ void __launch_bounds__(256) delta_weight_kernel(lyco_delta_desc_t d, const void* __restrict__ W,
                                                           void* __restrict__ out, int out_dtype,
                                                           float* __restrict__ norm_sq) {
  const int K = d.in_dim;
  const int64_t total = static_cast<int64_t>(d.out_dim) * K;
  const Chain ch{d.pre_round, d.pre_dtype, d.w_dtype, d.m_pre, d.m_post1, d.m_post2};
  const int fround = d.pre_round ? d.pre_dtype : LYCO_F32;
  float sq = 0.f;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / K), k = static_cast<int>(idx % K);
    float raw = 0.f;
    if (d.algo == LYCO_ALGO_LOKR) {
      const int pu = n / d.vp, pv = n % d.vp, u = k / d.vq, v = k % d.vq;
      raw = rnd(ld_f(d.f0, d.f_dtype, pu * d.uq + u), fround) *
            rnd(ld_f(d.f1, d.f_dtype, static_cast<int64_t>(pv) * d.vq + v), fround);
    } else if (d.algo == LYCO_ALGO_IA3) {
      // dW = W * (w * mult)  (ia3.py:91-102 make_weight(diff=True)); chain is the identity for IA3
      const int c = d.on_input ? (k / d.ia3_group) : n;
      raw = ld_f(W, d.w_dtype, idx) * (ld_f(d.f0, d.f_dtype, c) * d.m_post2);
    } else {
      float p1 = 0.f, p2 = 0.f;
      for (int r = 0; r < d.rank; ++r) {
        float b = ld_f(d.f1, d.f_dtype, static_cast<int64_t>(r) * K + k);
        if (d.m_in != 1.f) b = rnd(b * d.m_in, d.f_dtype);
        p1 = fmaf(rnd(ld_f(d.f0, d.f_dtype, static_cast<int64_t>(n) * d.rank + r), fround), rnd(b, fround), p1);
        if (d.algo == LYCO_ALGO_LOHA)
          p2 = fmaf(rnd(ld_f(d.f2, d.f_dtype, static_cast<int64_t>(n) * d.rank + r), fround),
                    rnd(ld_f(d.f3, d.f_dtype, static_cast<int64_t>(r) * K + k), fround), p2);
      }
      raw = d.algo == LYCO_ALGO_LOHA ? rnd(p1, fround) * rnd(p2, fround) : p1;
    }
    const float dw = d.algo == LYCO_ALGO_IA3 ? raw : apply_chain(raw, ch);
    sq = fmaf(dw, dw, sq);
    if (out != nullptr) {
      if (out_dtype == LYCO_F32) reinterpret_cast<float*>(out)[idx] = dw;
      else reinterpret_cast<uint16_t*>(out)[idx] = to16(dw, out_dtype);
    }
  }
  if (norm_sq != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    __shared__ float part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += part[i];
      atomicAdd(norm_sq, s);
    }
  }
}

}  // namespace lyco
