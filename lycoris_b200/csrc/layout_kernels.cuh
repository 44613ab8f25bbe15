// Activation layout pass for the implicit-GEMM convolutions: batched 2-D transpose with an optional
// fp32 -> 16-bit cast,  dst[b][c][r] = cast(src[b][r][c]).
//   NCHW -> NHWC : rows = C, cols = H*W   (fused with autocast's fp32 -> bf16 cast of the layer input)
//   NHWC -> NCHW : rows = H*W, cols = C
// HBM-bound: 64x64 tiles through shared memory, both the global reads (8 B / 4 B per thread along `cols`) and
// the global writes (4 B per thread along `rows`) are full 128-byte row segments.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include "pdl.cuh"

namespace lyco {

constexpr int TR_TILE = 64;

template <typename SRC>
__device__ __forceinline__ uint16_t to16(SRC v, int fmt);
template <>
__device__ __forceinline__ uint16_t to16<float>(float v, int fmt) {
  if (fmt == 1) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}
template <>
__device__ __forceinline__ uint16_t to16<uint16_t>(uint16_t v, int) { return v; }

// grid = (ceil(cols/64), ceil(rows/64), batch), block = (32, 8)
template <typename SRC>
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const SRC* __restrict__ src, uint16_t* __restrict__ dst, int rows, int cols, int fmt) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  __shared__ uint16_t tile[TR_TILE][TR_TILE + 2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c0 = blockIdx.x * TR_TILE, r0 = blockIdx.y * TR_TILE;
  const size_t plane = static_cast<size_t>(rows) * cols;
  const SRC* s = src + blockIdx.z * plane;
  uint16_t* d = dst + blockIdx.z * plane;
  const bool vec_in = (cols % 2 == 0), vec_out = (rows % 2 == 0);

  // Interior tiles (the whole 64x64 tile inside the matrix, 16-byte alignable pitches): 16-byte global accesses on
  // both sides.  Thread t loads 8 consecutive columns of row t/8 (+32), scatters them into the TRANSPOSED tile
  // (tile_t[c][r], pitch 66 halfwords), then reads 8 consecutive rows of one column as four 32-bit words
  // (bank = (c + 4*rg + q) mod 32: conflict-free) and writes them with one 16-byte store.
  if (r0 + TR_TILE <= rows && c0 + TR_TILE <= cols && cols % 8 == 0 && rows % 8 == 0) {
    uint16_t (*tile_t)[TR_TILE + 2] = tile;  // same storage, indexed [c][r]
    const int t = ty * 32 + tx;
    const int cg = t & 7, rr = t >> 3;  // 8 column groups x 32 rows per pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = rr + 32 * pass;
      const SRC* q = s + static_cast<size_t>(r0 + r) * cols + c0 + 8 * cg;
      uint16_t h[8];
      if (sizeof(SRC) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(q);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q) + 4);
        h[0] = to16<float>(a.x, fmt); h[1] = to16<float>(a.y, fmt); h[2] = to16<float>(a.z, fmt);
        h[3] = to16<float>(a.w, fmt); h[4] = to16<float>(b.x, fmt); h[5] = to16<float>(b.y, fmt);
        h[6] = to16<float>(b.z, fmt); h[7] = to16<float>(b.w, fmt);
      } else {
        *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(q);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) tile_t[8 * cg + j][r] = h[j];
    }
    __syncthreads();
    const int rg = t & 7, cc = t >> 3;  // 8 row groups x 32 columns per pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int c = cc + 32 * pass;
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&tile_t[c][8 * rg]);
      const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(d + static_cast<size_t>(c0 + c) * rows + r0 + 8 * rg) = v;
    }
    return;
  }

#pragma unroll
  for (int i = ty; i < TR_TILE; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    uint16_t a = 0, b = 0;
    if (r < rows) {
      const SRC* q = s + static_cast<size_t>(r) * cols + c;
      if (vec_in && c + 1 < cols) {
        if (sizeof(SRC) == 4) {
          const float2 v = *reinterpret_cast<const float2*>(q);
          a = to16<float>(v.x, fmt);
          b = to16<float>(v.y, fmt);
        } else {
          const uint32_t v = *reinterpret_cast<const uint32_t*>(q);
          a = static_cast<uint16_t>(v & 0xffffu);
          b = static_cast<uint16_t>(v >> 16);
        }
      } else {
        if (c < cols) a = to16<SRC>(q[0], fmt);
        if (c + 1 < cols) b = to16<SRC>(q[1], fmt);
      }
    }
    *reinterpret_cast<uint32_t*>(&tile[i][2 * tx]) = static_cast<uint32_t>(a) | (static_cast<uint32_t>(b) << 16);
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < TR_TILE; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;  // dst row = source column
    if (c >= cols) continue;
    const uint16_t a = tile[2 * tx][i], b = tile[2 * tx + 1][i];
    uint16_t* q = d + static_cast<size_t>(c) * rows + r;
    if (vec_out && r + 1 < rows) {
      *reinterpret_cast<uint32_t*>(q) = static_cast<uint32_t>(a) | (static_cast<uint32_t>(b) << 16);
    } else {
      if (r < rows) q[0] = a;
      if (r + 1 < rows) q[1] = b;
    }
  }
}

// Filter re-layouts around the implicit-GEMM convolutions (T = R*S taps, small).  All three read and write runs of
// >= 64 contiguous bytes through a shared-memory tile.
//   FILTER_FPROP : W' [O][C][T]  (PyTorch's [O,C,R,S])  -> Wk [O][T][C]        the fprop B operand
//   FILTER_DGRAD : W' [O][C][T]                          -> Wd [C][T'][O], T' = T-1-t   flipped + transposed, the
//                                                           B operand of dgrad-as-fprop
//   FILTER_WBACK : dWk [O][T][C] fp32 (wgrad output)     -> dW' [O][C][T] fp32  (the layout the factor-gradient
//                                                           kernels index)
enum { FILTER_FPROP = 0, FILTER_DGRAD = 1, FILTER_WBACK = 2 };
constexpr int FR_CCHUNK = 128;   // channels per block for the per-output-channel modes
constexpr int FR_MAX_T = 25;     // up to 5x5 filters

// FILTER_FPROP / FILTER_WBACK: grid = (ceil(C / FR_CCHUNK), O), block = 256.  All global traffic is 16-byte
// vectors (C % 8 == 0 for 16-bit, C % 4 == 0 for fp32; 16-byte aligned bases); the transposition happens in the
// shared-memory tile, kept in [c][t] order.
template <typename T, bool BACK>
__global__ void __launch_bounds__(256) filter_row_relayout_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                   int C, int taps) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  constexpr int V = 16 / sizeof(T);
  extern __shared__ uint4 fr_smem4[];
  T* tile = reinterpret_cast<T*>(fr_smem4);  // [cn][taps]
  const int o = blockIdx.y, c0 = blockIdx.x * FR_CCHUNK;
  const int cn = min(FR_CCHUNK, C - c0);
  const size_t row = static_cast<size_t>(o) * C * taps;
  const int nvec = cn * taps / V, groups = cn / V, items = groups * taps;
  if (!BACK) {
    const uint4* src = reinterpret_cast<const uint4*>(in + row + static_cast<size_t>(c0) * taps);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) fr_smem4[i] = src[i];
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int t = i / groups, cg = i - t * groups;
      T v[V];
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = tile[(cg * V + j) * taps + t];
      *reinterpret_cast<uint4*>(out + row + static_cast<size_t>(t) * C + c0 + cg * V) = *reinterpret_cast<uint4*>(v);
    }
  } else {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int t = i / groups, cg = i - t * groups;
      T v[V];
      *reinterpret_cast<uint4*>(v) =
          *reinterpret_cast<const uint4*>(in + row + static_cast<size_t>(t) * C + c0 + cg * V);
#pragma unroll
      for (int j = 0; j < V; ++j) tile[(cg * V + j) * taps + t] = v[j];
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + row + static_cast<size_t>(c0) * taps);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = fr_smem4[i];
  }
}

// FILTER_DGRAD: grid = (ceil(C / 32), ceil(O / 32)), block = 256; tile [32 o][32 c * taps (+8 pad)], 16-byte
// vectors on both sides (C % 8 == 0, O % 8 == 0).
__global__ void __launch_bounds__(256) filter_dgrad_relayout_kernel(const uint16_t* __restrict__ in,
                                                                     uint16_t* __restrict__ out, int O, int C,
                                                                     int taps) {
  pdl_trigger();  // a tensor-core kernel behind this one may start its prologue (pdl.cuh)
  extern __shared__ uint4 fr_smem4[];
  uint16_t* tile = reinterpret_cast<uint16_t*>(fr_smem4);
  const int o0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int on = min(32, O - o0), cn = min(32, C - c0);
  const int run = cn * taps, rvec = run / 8, pitch = 32 * taps + 8;
  for (int i = threadIdx.x; i < on * rvec; i += blockDim.x) {
    const int oo = i / rvec, k = i - oo * rvec;
    reinterpret_cast<uint4*>(tile + oo * pitch)[k] =
        reinterpret_cast<const uint4*>(in + (static_cast<size_t>(o0 + oo) * C + c0) * taps)[k];
  }
  __syncthreads();
  // out[c][t'][o]: for fixed (c, t') the output channels are contiguous; 4 lanes write one 64-byte run
  const int og = on / 8;
  for (int i = threadIdx.x; i < run * og; i += blockDim.x) {
    const int g = i % og, ct = i / og;
    const int c = ct / taps, tp = ct - c * taps;
    uint16_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = tile[(g * 8 + j) * pitch + c * taps + (taps - 1 - tp)];
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(c0 + c) * taps + tp) * O + o0 + g * 8) =
        *reinterpret_cast<uint4*>(v);
  }
}

}  // namespace lyco
