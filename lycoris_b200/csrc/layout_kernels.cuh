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
  __shared__ uint16_t tile[TR_TILE][TR_TILE + 2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c0 = blockIdx.x * TR_TILE, r0 = blockIdx.y * TR_TILE;
  const size_t plane = static_cast<size_t>(rows) * cols;
  const SRC* s = src + blockIdx.z * plane;
  uint16_t* d = dst + blockIdx.z * plane;
  const bool vec_in = (cols % 2 == 0), vec_out = (rows % 2 == 0);

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

}  // namespace lyco
