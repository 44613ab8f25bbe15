// Structured LoKr factor gradients (HBM-bound helpers around two skinny tcgen05 GEMMs).
//
// dW = kron(w1 [up x uq], w2 [vp x vq]):  dW[pu*vp+pv, u*vq+v] = w1[pu,u] * w2[pv,v]   (torch.kron convention,
// reference lycoris/functional/lokr.py:11-20; the structured contraction order is the reference's own bypass path,
// lycoris/modules/lokr.py:468-538).  With dY [M, up*vp] and X [M, uq*vq] the factor gradients
//
//     g_w2[pv,v] = sum_{m,pu,u} w1[pu,u] dY[m,pu,pv] X[m,u,v]
//     g_w1[pu,u] = sum_{m,pv,v} w2[pv,v] dY[m,pu,pv] X[m,u,v]
//
// never need the dense dW' = dY^T X [N x K]:
//   mix the SMALLER activation with w1 ............... Xt[m,pu,v] = sum_u w1[pu,u] X[m,u,v]      (lokr_mix_kernel)
//   g_w2 = dY2^T Xt2, dY2 = dY as [M*up, vp] ......... ONE tcgen05 GEMM, reduction over M*up, 1/uq of the dense FLOPs
//   Q = dY2 w2  [M*up, vq] ........................... ONE tcgen05 GEMM, 1/uq of the dense FLOPs
//   g_w1[pu,u] = sum_{m,v} Q[m,pu,v] X[m,u,v] ........ lokr_w1grad_kernel (reads Q and X once)
// (or the mirror image with dY mixed instead of X when N < K).  All views are free: [M, up*vp] row-major IS
// [M*up, vp] row-major.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "weight_kernels.cuh"
#include "pdl.cuh"

namespace lyco {

constexpr int LK_MAX_G = 16;  // w1 blocks up to 16 x 16 (factor <= 16); larger w1 uses the dense path

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8], int fmt) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (fmt == 1) {
      f[2 * t] = __uint_as_float(w[t] << 16);
      f[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[t]);
      f[2 * t] = __low2float(h);
      f[2 * t + 1] = __high2float(h);
    }
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8], int fmt) {
  uint32_t w[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (fmt == 1) {
      __nv_bfloat162 v = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
      w[t] = *reinterpret_cast<uint32_t*>(&v);
    } else {
      __half2 v = __floats2half2_rn(f[2 * t], f[2 * t + 1]);
      w[t] = *reinterpret_cast<uint32_t*>(&v);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// out[m, a, c] = sum_b Wm(a, b) * in[m, b, c]      in: [M, nb, nc]   out: [M, na, nc]   16-bit, nc % 8 == 0
// Wm(a, b) = w[a * ldw + b]  (trans = 0)   or   w[b * ldw + a]  (trans = 1)
// One thread owns one (m, 8-column vector): issues the NB input loads (16 bytes each) back to back — NB * 16 bytes in
// flight per thread, which is what keeps HBM busy at ~2 resident CTAs per SM — and keeps the NA output vectors in
// registers.  2 bytes read + 2 bytes written per element against 2*na FMAs: bandwidth-bound.
template <int NA, int NB>
__global__ void __launch_bounds__(256) lokr_mix_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                       const void* __restrict__ w, int w_dtype, int ldw, int trans,
                                                       int64_t M, int na, int nb, int nc8, int fmt,
                                                       float* __restrict__ zero_buf, int64_t zero_n) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  __shared__ float sw[LK_MAX_G * LK_MAX_G];
  // optional side job: zero-fill the (small) fp32 gradient buffers the NEXT kernels reduce into with atomics — this
  // kernel precedes them on the stream, so the two memset nodes per layer-step disappear from the captured graph
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < zero_n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    zero_buf[i] = 0.f;
  for (int i = threadIdx.x; i < na * nb; i += blockDim.x) {
    const int a = i / nb, b = i % nb;
    sw[i] = ld_f(w, w_dtype, trans ? static_cast<int64_t>(b) * ldw + a : static_cast<int64_t>(a) * ldw + b);
  }
  __syncthreads();
  const int64_t total = M * nc8;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = idx / nc8;
    const int c8 = static_cast<int>(idx - m * nc8);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (m * nb) * nc8 + c8;
    float acc[NA][8];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    uint4 raw[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b)
      raw[b] = b < nb ? __ldg(src + static_cast<int64_t>(b) * nc8) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (b < nb) {
        float x[8];
        unpack8(raw[b], x, fmt);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          if (a < na) {
            const float wv = sw[a * nb + b];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a][j] = fmaf(wv, x[j], acc[a][j]);
          }
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + (m * na) * nc8 + c8;
#pragma unroll
    for (int a = 0; a < NA; ++a)
      if (a < na) dst[static_cast<int64_t>(a) * nc8] = pack8(acc[a], fmt);
  }
}

// g[a, b] = gscale * sum_{m, c} P[m, a, c] * R[m, b, c]     P: [M, na, nc]   R: [M, nb, nc]   16-bit, nc % 8 == 0
// Thread = one (m, 8-column vector) per grid-stride step; NA x NB partial sums live in registers for the whole
// kernel and are reduced once at the end (warp shuffles -> shared -> one global atomic per entry per CTA).
template <int NA, int NB>
__global__ void __launch_bounds__(256) lokr_w1grad_kernel(const uint16_t* __restrict__ P, const uint16_t* __restrict__ R,
                                                          float* __restrict__ g, int64_t M, int na, int nb, int nc8,
                                                          float gscale, int fmt) {
  float acc[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[a][b] = 0.f;
  const int64_t total = M * nc8;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = idx / nc8;
    const int c8 = static_cast<int>(idx - m * nc8);
    const uint4* pp = reinterpret_cast<const uint4*>(P) + (m * na) * nc8 + c8;
    const uint4* rp = reinterpret_cast<const uint4*>(R) + (m * nb) * nc8 + c8;
    // all NA + NB vectors of this (m, c8) are requested before any is consumed: 16 x 16 bytes in flight per thread is
    // what hides the DRAM latency at one resident CTA per SM (round-2 profile: with the P loads issued one per
    // a-iteration the kernel ran at 0.78 TB/s, latency-bound)
    uint4 rraw[NB], praw[NA];
#pragma unroll
    for (int b = 0; b < NB; ++b) rraw[b] = b < nb ? __ldg(rp + static_cast<int64_t>(b) * nc8) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int a = 0; a < NA; ++a) praw[a] = a < na ? __ldg(pp + static_cast<int64_t>(a) * nc8) : make_uint4(0, 0, 0, 0);
    float r[NB][8];
#pragma unroll
    for (int b = 0; b < NB; ++b) unpack8(rraw[b], r[b], fmt);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      float p[8];
      unpack8(praw[a], p, fmt);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float s = acc[a][b];
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(p[j], r[b][j], s);
        acc[a][b] = s;
      }
    }
  }
  // Reduction of the NA*NB per-lane partial sums.  Butterfly over the VALUE index: at each step a lane keeps one half
  // of its values and hands the other half to its partner (NA*NB - 2 shuffles in total instead of 5 per value), so
  // after five steps lane l holds the warp totals of values {2*rev(l), 2*rev(l)+1}.  Warp totals go to a per-warp row
  // of shared memory (plain stores — shared-memory float atomics are CAS loops), 64 threads add the 8 rows and issue
  // ONE global atomic per entry per CTA.
  constexpr int NV = NA * NB;
  static_assert(NV == 64 || NV == 16, "value count handled by the butterfly below");
  float v[NV];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) v[a * NB + b] = acc[a][b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;  // index of v[0] after the steps so far
#pragma unroll
  for (int off = 16, n = NV; off >= 1 && n >= 2; off >>= 1, n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    if (upper) base += half;
  }
  // NV = 64: five steps leave 2 values per lane; NV = 16: four steps (off 16..2) leave 1 value per lane PAIR member —
  // handle generically: `left` values per lane, lanes that differ only in the unused low offset bits hold duplicates
  constexpr int STEPS = (NV == 64) ? 5 : 4;
  constexpr int left = NV >> STEPS;
  if (NV == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);  // the offset the 16-value butterfly did not use
  __shared__ float red[8][NV];
  if (NV == 64 || (lane & 1) == 0) {
#pragma unroll
    for (int i = 0; i < left; ++i) red[warp][base + i] = v[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][i];
    const int a = i / NB, b = i % NB;
    if (a < na && b < nb) atomicAdd(&g[a * nb + b], s * gscale);
  }
}

}  // namespace lyco
