// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias)      16-bit operands, fp32 accumulation in TMEM
//
// 256 threads per CTA, one CTA per SM:
//   warp 0      TMA producer  (one elected lane)      global -> smem ring (SWIZZLE_128B boxes)
//   warp 1      MMA issuer    (one elected lane)      tcgen05.mma, D in TMEM
//   warp 2      TMEM allocator / deallocator
//   warps 4..7  epilogue: tcgen05.ld -> registers -> (bias, cast) -> swizzled smem -> TMA store
//               (fp32 outputs: the same staging, TMA store or TMA reduce-add for split-K partials)
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two accumulator
// stages so the epilogue of tile i overlaps the main loop of tile i+1), and the static persistent
// tile schedule.  BLOCK_N is a RUNTIME multiple of 32 (32..256) picked per problem by the host so
// that the last wave of tiles is as full as possible.
//
// PAIR = true: two CTAs of one TPC (cluster 2x1x1) share a 256 x BLOCK_N tile with
// tcgen05.mma.cta_group::2 — each stages its own 128 rows of A and HALF of the B tile, which cuts the
// L2 -> SM feed per CTA from (128 + BLOCK_N) to (128 + BLOCK_N/2) rows per k-block (the single-CTA
// kernel is feed-bound at 128x256).  The leader CTA issues the MMAs; TMA completion of both CTAs is
// signalled on the leader's full barrier; tcgen05.commit multicasts to both CTAs.
//
// Either operand may be K-major (reduction index contiguous in memory) or MN-major (output index
// contiguous): the three contractions of a wrapped layer are all served without transposes
//   forward  Y  = X  * W'^T        A = X  [M,K]  K-major    B = W' [N,K]  K-major
//   dgrad    dX = dY * W'          A = dY [M,N]  K-major    B = W' [N,K]  MN-major (as [K_out, N_red])
//   wgrad    dW'= dY^T * X         A = dY [M,N]  MN-major   B = X  [M,K]  MN-major, split over M
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "pdl.cuh"
#include "ptx_sm100.cuh"

namespace lyco {

constexpr int GEMM_BLOCK_M = 128;   // rows per CTA (a pair covers 256)
constexpr int GEMM_BLOCK_K = 64;    // 64 x 16-bit = 128 B = one swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KiB
constexpr int GEMM_ATOM_BYTES = 64 * GEMM_BLOCK_K * 2;         // one MN-major 64x64 box, 8 KiB
constexpr int GEMM_RING_BYTES = 192 * 1024;
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_EPI_BYTES = 4 * 2 * 2048;   // 4 epilogue warps x 2 staging buffers x (32 rows x 64 B)
constexpr int GEMM_TMEM_COLS = 512;            // two accumulator stages at columns 0 and 256
constexpr int GEMM_SMEM_BYTES = GEMM_RING_BYTES + GEMM_EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum { EPI_STORE16 = 0, EPI_STORE_F32 = 1, EPI_ATOMIC_F32 = 2, EPI_STORE16_NCHW = 3, EPI_STORE_F32_NCHW = 4 };
__host__ __device__ constexpr bool epi_uses_tma(int epi) { return epi != EPI_STORE_F32_NCHW; }

struct GemmParams {
  void* C;
  const void* bias;
  int64_t ldc;
  int M, N, K;
  int m_tiles, n_tiles, splits, k_blocks;
  int k_blocks1;    // k-blocks [0, k_blocks1) read operand pair (A, B); [k_blocks1, k_blocks) read the SECOND pair
                    // (A2, B2) into the same accumulator: C = A·Bᵀ + A2·B2ᵀ (the rank-r side path of a low-rank adapter)
  int block_n;      // columns per tile (multiple of 32; MN-major B: of 64, pairs: of 128)
  int stages;       // smem ring depth for this block_n
  int fmt;          // operand / 16-bit output format: 0 = f16, 1 = bf16
  int bias_dtype;   // LYCO_BF16 / LYCO_F16 / LYCO_F32
  int epi_pq;       // EPI_STORE16_NCHW: output pixels per image (rows of C are (image, pixel); multiple of 32)
#ifdef LYCO_GEMM_TRACE
  unsigned long long* trace;  // debug build only (tools/gemm_trace.py): 64 clock64 slots per CTA
#endif
};

// Timeline probes of the debug build: slot 0 entry, 1 globaltimer at entry, 2 setup done, 3 first TMA issued,
// 4 first operand stage landed, 5 last TMA issued, 6 last MMA commit issued, 8+2i / 9+2i accumulator i ready / drained
// (epilogue warp 4), 40 exit, 41 globaltimer at exit, 42 tiles of this CTA.  Compiled out of the product library.
#ifdef LYCO_GEMM_TRACE
#define LYCO_TRACE(slot, value)                                                            \
  do {                                                                                      \
    if (p.trace) p.trace[static_cast<size_t>(blockIdx.x) * 64 + (slot)] = (value);          \
  } while (0)
__device__ __forceinline__ unsigned long long trace_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#else
#define LYCO_TRACE(slot, value) do {} while (0)
#endif

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

// v[0..31] += bias[col0 .. col0+31]  (columns >= n_limit untouched)
__device__ __forceinline__ void add_bias32(float (&v)[32], const GemmParams& p, int col0, int n_limit) {
  if (p.bias == nullptr) return;
  const bool full = col0 + 32 <= n_limit;
  if (full && p.bias_dtype != 2 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
    // 32 x 16-bit values = four 16-byte loads (every lane reads the same addresses: L1 broadcast)
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
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < n_limit) v[j] += load_scalar(p.bias, p.bias_dtype, col0 + j);
  }
}

// fp32 epilogue for one 32-column chunk of this warp's 32 rows (plain store, or the reduction of split-K partials /
// accumulate mode): two 16-column halves, each staged as a SWIZZLE_64B tile (32 rows x 64 B, the layout of the 16-bit
// epilogue) and written by ONE TMA operation — cp.async.bulk.tensor store or cp.reduce.async.bulk.tensor .add.  The
// round-1 epilogue issued one 16-byte red.global per lane and row (32 different lines per warp instruction): the drain
// of a 256-column accumulator took 6.4 us and held the next unit's main loop back (tools/gemm_trace.py, wgrad
// 1280 x 1280 x 8192).  The TMA unit clips at M / N; `n_limit` (the tap boundary of the convolution's weight gradient,
// a multiple of 64 columns) only ever drops whole halves.
template <int EPI>
__device__ __forceinline__ void store_chunk_f32(const uint32_t (&r)[32], int row0, int col0, const GemmParams& p,
                                                int n_limit, const CUtensorMap* tmap_c, uint8_t* stage, int& buf,
                                                int lane) {
  if (row0 >= p.M) return;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c0 = col0 + 16 * h;
    if (c0 >= n_limit) break;  // warp-uniform; a skipped half commits no bulk group (see store_chunk_tma)
    if (lane == 0) ptx::tma_store_wait_read<1>();
    __syncwarp();
    uint8_t* tile = stage + buf * 2048;
    const uint32_t base = ptx::smem_u32(tile) + lane * 64;
    const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t a = base + ((q ^ sw) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[16 * h + 4 * q]),
                   "r"(r[16 * h + 4 * q + 1]), "r"(r[16 * h + 4 * q + 2]), "r"(r[16 * h + 4 * q + 3])
                   : "memory");
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (EPI == EPI_ATOMIC_F32) ptx::tma_reduce_add_2d(tmap_c, tile, c0, row0);
      else ptx::tma_store_2d(tmap_c, tile, c0, row0);
      ptx::tma_store_commit();
    }
    buf ^= 1;
  }
}

// fp32 NCHW epilogue (the input gradient of a convolution whose input was fp32 under autocast): lane = pixel, so
// for each channel the warp writes 32 consecutive floats — one full 128-byte line per store, no staging needed.
__device__ __forceinline__ void store_chunk_f32_nchw(const uint32_t (&r)[32], int row0, int col0, const GemmParams& p,
                                                     int n_limit, int lane) {
  if (col0 >= n_limit || row0 >= p.M) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  add_bias32(v, p, col0, n_limit);
  const int img = row0 / p.epi_pq;
  float* base = reinterpret_cast<float*>(p.C) +
                (static_cast<int64_t>(img) * p.N + col0) * p.epi_pq + (row0 - img * p.epi_pq) + lane;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (col0 + j < n_limit) base[static_cast<int64_t>(j) * p.epi_pq] = v[j];
}

// 16-bit epilogue for one 32-column chunk of this warp's 32 rows: registers -> SWIZZLE_64B staging
// tile (32 rows x 64 B) -> one TMA store (full 64-byte row segments, asynchronous, clipped at M / N).
// `stage` is this warp's pair of 2 KiB buffers; at most one store is left in flight when a buffer is reused.
__device__ __forceinline__ void store_chunk_tma(const uint32_t (&r)[32], int row0, int col0, const GemmParams& p,
                                                int n_limit, const CUtensorMap* tmap_c, uint8_t* stage, int& buf,
                                                int lane) {
  // Nothing to write for this chunk (tile padding past N or M): leave the staging buffers alone — a
  // skipped chunk commits no bulk group, so wait_group.read<1> below would not protect the buffer of
  // the last real store from being overwritten (warp-uniform condition).
  if (col0 >= n_limit || row0 >= p.M) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  add_bias32(v, p, col0, n_limit);
  if (lane == 0) ptx::tma_store_wait_read<1>();  // the buffer written two chunks ago has been read out
  __syncwarp();
  uint8_t* tile = stage + buf * 2048;
  const uint32_t base = ptx::smem_u32(tile) + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;  // Swizzle<2,4,3>: 16-byte chunk index ^= address bits [7,9)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t a = base + ((q ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack16(v[8 * q + 0], v[8 * q + 1], p.fmt)),
                 "r"(pack16(v[8 * q + 2], v[8 * q + 3], p.fmt)), "r"(pack16(v[8 * q + 4], v[8 * q + 5], p.fmt)),
                 "r"(pack16(v[8 * q + 6], v[8 * q + 7], p.fmt))
                 : "memory");
  }
  ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_2d(tmap_c, tile, col0, row0);
    ptx::tma_store_commit();
  }
  buf ^= 1;
}

// Same chunk written channel-major: the convolution's output as NCHW.  Rows of the tile are output pixels of ONE
// image (epi_pq % 32 == 0), so the staging tile is [32 channels][32 pixels] (64 B per channel) and one 3-D TMA
// store (pixel, channel, image) writes 32 contiguous pixels of each channel.
__device__ __forceinline__ void store_chunk_tma_nchw(const uint32_t (&r)[32], int row0, int col0, const GemmParams& p,
                                                     int n_limit, const CUtensorMap* tmap_c, uint8_t* stage, int& buf,
                                                     int lane) {
  if (col0 >= n_limit || row0 >= p.M) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  add_bias32(v, p, col0, n_limit);
  if (lane == 0) ptx::tma_store_wait_read<1>();
  __syncwarp();
  uint8_t* tile = stage + buf * 2048;
  const uint32_t base = ptx::smem_u32(tile) + lane * 2;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const uint32_t pk = pack16(v[j], v[j + 1], p.fmt);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(base + j * 64), "h"(static_cast<uint16_t>(pk & 0xffffu)) : "memory");
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(base + (j + 1) * 64), "h"(static_cast<uint16_t>(pk >> 16)) : "memory");
  }
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    const int img = row0 / p.epi_pq;
    ptx::tma_store_3d(tmap_c, tile, row0 - img * p.epi_pq, col0, img);
    ptx::tma_store_commit();
  }
  buf ^= 1;
}

// Work distribution, identical in the three roles of a CTA (pair): units are (tile, reduction split) pairs handed
// out round-robin, splits of one tile on neighbouring workers, n fastest across tiles (neighbours share A rows in
// L2).  Every unit of a launch has the same number of k-blocks (+-1), which keeps the fp32-atomic epilogue of one
// unit hidden behind the main loop of the next (a flattened equal-range scheduler with ragged units was measured
// 5-10 % slower on the wgrad shapes: short units expose their epilogue).
struct WorkIter {
  int tile, kb0, kb1;  // current unit
  int w, step, total, splits, k_blocks;
  __device__ __forceinline__ WorkIter(const GemmParams& p, int worker, int workers)
      : tile(0), kb0(0), kb1(0), w(worker), step(workers), total(p.m_tiles * p.n_tiles * p.splits), splits(p.splits),
        k_blocks(p.k_blocks) {}
  __device__ __forceinline__ bool next() {
    if (w >= total) return false;
    const int split = w % splits;
    tile = w / splits;
    kb0 = static_cast<int>(static_cast<int64_t>(split) * k_blocks / splits);
    kb1 = static_cast<int>(static_cast<int64_t>(split + 1) * k_blocks / splits);
    w += step;
    return true;
  }
  // true when the unit next() just returned is this worker's last one
  __device__ __forceinline__ bool last() const { return w >= total; }
};

// The LAST unit of a CTA has nothing to hide its drain behind (2.5 us of a 27 us launch at 8192 x 1280 x 1280,
// tools/gemm_trace.py), and by then warps 0..3 (producer, MMA issuer, TMEM allocator, spare) are idle: they take the
// upper half of the accumulator's column chunks — warp w reads the same TMEM lanes as epilogue warp 4 + w — staging
// through the operand ring, which is dead once the last accumulator is complete.  `tlast_bar` is committed by the MMA
// issuer together with the last tfull barrier; it completes exactly once, so the helpers need no phase bookkeeping.
__device__ __forceinline__ int main_chunks_of_last_unit(int block_n) { return ((block_n >> 5) + 1) >> 1; }

template <bool PAIR>
__device__ __forceinline__ void gemm_setup(uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tfull_bar,
                                           uint64_t* tempty_bar, uint64_t* tlast_bar, uint32_t* tmem_slot, int stages,
                                           int warp, int lane) {
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(tlast_bar, 1);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], PAIR ? 2 : 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs)
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) ptx::tmem_alloc_pair(tmem_slot, GEMM_TMEM_COLS);
    else ptx::tmem_alloc(tmem_slot, GEMM_TMEM_COLS);
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync();  // barriers of both CTAs live before any remote arrive / multicast commit
  else __syncthreads();
  ptx::tc_fence_after();
}

template <bool PAIR>
__device__ __forceinline__ void gemm_teardown(uint32_t tmem_base, int warp) {
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync();  // neither CTA may exit (or free TMEM) while its partner still uses it
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, GEMM_TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

// MMA issue for one k-block (4 x K=16) + slot release
template <bool PAIR, bool A_MN, bool B_MN>
__device__ __forceinline__ void gemm_issue_kblock(uint32_t sa, uint32_t sb, uint32_t d_tmem, uint32_t idesc,
                                                  bool first, uint64_t* empty_slot) {
#pragma unroll
  for (int kk = 0; kk < GEMM_BLOCK_K / GEMM_UMMA_K; ++kk) {
    // K-major : advance 16 elements (32 B) inside the 128 B swizzle row
    // MN-major: advance 16 k-rows of 128 B
    const uint64_t da = A_MN ? ptx::make_smem_desc(sa + kk * 2048, GEMM_ATOM_BYTES, 1024)
                             : ptx::make_smem_desc(sa + kk * 32, 16, 1024);
    const uint64_t db = B_MN ? ptx::make_smem_desc(sb + kk * 2048, GEMM_ATOM_BYTES, 1024)
                             : ptx::make_smem_desc(sb + kk * 32, 16, 1024);
    const uint32_t acc = (!first || kk > 0) ? 1u : 0u;
    if (PAIR) ptx::umma_f16_pair(d_tmem, da, db, idesc, acc);
    else ptx::umma_f16(d_tmem, da, db, idesc, acc);
  }
  if (PAIR) ptx::umma_commit_pair(empty_slot, 0b11);  // frees this slot in BOTH CTAs once the MMAs retire
  else ptx::umma_commit(empty_slot);
}

// Drain one accumulator (this warp's 32 rows x block_n columns) from TMEM and write it out.
// Chunks [c_begin, c_end) of 32 columns; `tempty_slot` != nullptr: arrive on it after this warp's last TMEM read.
template <bool PAIR, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(uint32_t t_row, int row0, int col_base, int n_limit, int c_begin,
                                                   int c_end, const GemmParams& p, const CUtensorMap* tmap_c,
                                                   uint8_t* stage, int& buf, uint64_t* tempty_slot, int lane) {
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(t_row + c * 32, r);
    ptx::tmem_ld_wait();
    if (tempty_slot != nullptr && c == c_end - 1) {
      // all of this warp's TMEM reads are done: hand the accumulator back to the MMA thread
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(tempty_slot), 0));
        else ptx::mbar_arrive(tempty_slot);
      }
    }
    const int col0 = col_base + c * 32;
    if (EPI == EPI_STORE16) store_chunk_tma(r, row0, col0, p, n_limit, tmap_c, stage, buf, lane);
    else if (EPI == EPI_STORE16_NCHW) store_chunk_tma_nchw(r, row0, col0, p, n_limit, tmap_c, stage, buf, lane);
    else if (EPI == EPI_STORE_F32_NCHW) store_chunk_f32_nchw(r, row0, col0, p, n_limit, lane);
    else store_chunk_f32<EPI>(r, row0, col0, p, n_limit, tmap_c, stage, buf, lane);
  }
}

// Warps 0..3 after their roles: upper column chunks of the CTA's last unit (see main_chunks_of_last_unit).
template <bool PAIR, int EPI>
__device__ __forceinline__ void gemm_last_unit_helper(uint32_t tmem_base, int acc, int row0, int col_base, int n_limit,
                                                      int block_n, const GemmParams& p, const CUtensorMap* tmap_c,
                                                      uint8_t* ring, uint64_t* tlast_bar, int warp, int lane) {
  const int chunks = block_n >> 5, c0 = main_chunks_of_last_unit(block_n);
  if (c0 >= chunks) return;
  ptx::mbar_wait_parked(tlast_bar, 0, lane);
  ptx::tc_fence_after();
  const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(warp * 32) << 16);
  int buf = 0;
  gemm_epilogue_tile<PAIR, EPI>(t_row, row0, col_base, n_limit, c0, chunks, p, tmap_c, ring + warp * 4096, buf,
                                nullptr, lane);
  if (epi_uses_tma(EPI) && lane == 0) ptx::tma_store_wait_read<0>();
}

template <bool PAIR, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_a2,
                  const __grid_constant__ CUtensorMap tmap_b2, const GemmParams p) {
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
  const int b_rows = PAIR ? block_n / 2 : block_n;  // B rows (N columns) this CTA stages
  const int stage_bytes = GEMM_A_BYTES + b_rows * GEMM_BLOCK_K * 2;
  const int stages = p.stages;
  const int rows_per_tile = PAIR ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (epi_uses_tma(EPI)) ptx::prefetch_tmap(&tmap_c);
    if (p.k_blocks1 < p.k_blocks) {
      ptx::prefetch_tmap(&tmap_a2);
      ptx::prefetch_tmap(&tmap_b2);
    }
  }
#ifdef LYCO_GEMM_TRACE
  if (threadIdx.x == 0) {
    LYCO_TRACE(0, clock64());
    LYCO_TRACE(1, trace_globaltimer());
  }
#endif
  gemm_setup<PAIR>(full_bar, empty_bar, tfull_bar, tempty_bar, tlast_bar, tmem_slot, stages, warp, lane);
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // everything above overlapped the previous kernel; its results are visible from here (pdl.cuh)
  pdl_trigger();
  if (threadIdx.x == 0) LYCO_TRACE(2, clock64());

  const int workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    LYCO_TRACE(3, clock64());
    for (WorkIter it(p, worker, workers); it.next();) {
      const int n_idx = it.tile % p.n_tiles;
      const int m_idx = it.tile / p.n_tiles;
      const int kb0 = it.kb0, kb1 = it.kb1;
      const int m0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M;
      const int n0 = n_idx * block_n + static_cast<int>(rank) * b_rows;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + GEMM_A_BYTES;
        // second operand pair (side path): same tile coordinates, its own reduction index
        const bool second = kb >= p.k_blocks1;
        const CUtensorMap* ma = second ? &tmap_a2 : &tmap_a;
        const CUtensorMap* mb = second ? &tmap_b2 : &tmap_b;
        const int kc = (second ? kb - p.k_blocks1 : kb) * GEMM_BLOCK_K;
        if (PAIR) {
          const uint32_t full_leader = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
          ptx::mbar_expect_tx_cluster(full_leader, stage_bytes);
          if (!A_MN) {
            ptx::tma_load_2d_pair(sa, ma, full_leader, kc, m0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
              ptx::tma_load_2d_pair(sa + j * GEMM_ATOM_BYTES, ma, full_leader, m0 + j * 64, kc);
          }
          if (!B_MN) {
            ptx::tma_load_2d_pair(sb, mb, full_leader, kc, n0);
          } else {
            for (int j = 0; j < b_rows / 64; ++j)
              ptx::tma_load_2d_pair(sb + j * GEMM_ATOM_BYTES, mb, full_leader, n0 + j * 64, kc);
          }
        } else {
          ptx::mbar_expect_tx(&full_bar[stage], stage_bytes);
          if (!A_MN) {
            ptx::tma_load_2d(sa, ma, &full_bar[stage], kc, m0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)
              ptx::tma_load_2d(sa + j * GEMM_ATOM_BYTES, ma, &full_bar[stage], m0 + j * 64, kc);
          }
          if (!B_MN) {
            ptx::tma_load_2d(sb, mb, &full_bar[stage], kc, n0);
          } else {
            for (int j = 0; j < b_rows / 64; ++j)
              ptx::tma_load_2d(sb + j * GEMM_ATOM_BYTES, mb, &full_bar[stage], n0 + j * 64, kc);
          }
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    LYCO_TRACE(5, clock64());
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ------------------------------------------------- MMA issuer (leader CTA only for a pair)
    const uint32_t idesc = ptx::make_idesc_f16(p.fmt, rows_per_tile, block_n, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
#ifdef LYCO_GEMM_TRACE
    bool trace_first = false;
#endif
    for (WorkIter it(p, worker, workers); it.next();) {
      const int kb0 = it.kb0, kb1 = it.kb1;
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
#ifdef LYCO_GEMM_TRACE
        if (!trace_first) { trace_first = true; LYCO_TRACE(4, clock64()); }
#endif
        const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
        gemm_issue_kblock<PAIR, A_MN, B_MN>(sa, sa + GEMM_A_BYTES, d_tmem, idesc, kb == kb0, &empty_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (PAIR) ptx::umma_commit_pair(&tfull_bar[acc], 0b11);  // accumulator halves complete in both CTAs
      else ptx::umma_commit(&tfull_bar[acc]);
      if (it.last()) {
        if (PAIR) ptx::umma_commit_pair(tlast_bar, 0b11);
        else ptx::umma_commit(tlast_bar);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    LYCO_TRACE(6, clock64());
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew + 32)
    uint8_t* stage_buf = epi_smem + ew * 4096;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
#ifdef LYCO_GEMM_TRACE
    int trace_tile = 0;
#endif
    for (WorkIter it(p, worker, workers); it.next();) {
      const int n_idx = it.tile % p.n_tiles;
      const int m_idx = it.tile / p.n_tiles;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
#ifdef LYCO_GEMM_TRACE
      if (ew == 0 && lane == 0 && trace_tile < 16) LYCO_TRACE(8 + 2 * trace_tile, clock64());
#endif
      const int row0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M + ew * 32;
      const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(ew * 32) << 16);
      const int chunks = it.last() ? main_chunks_of_last_unit(block_n) : (block_n >> 5);
      gemm_epilogue_tile<PAIR, EPI>(t_row, row0, n_idx * block_n, p.N, 0, chunks, p, &tmap_c, stage_buf, buf,
                                    &tempty_bar[acc], lane);
#ifdef LYCO_GEMM_TRACE
      if (ew == 0 && lane == 0 && trace_tile < 16) LYCO_TRACE(9 + 2 * trace_tile, clock64());
      ++trace_tile;
#endif
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (epi_uses_tma(EPI) && lane == 0) ptx::tma_store_wait_read<0>();  // staging smem must outlive the stores
#ifdef LYCO_GEMM_TRACE
    if (ew == 0 && lane == 0) LYCO_TRACE(42, static_cast<unsigned long long>(trace_tile));
#endif
  }

  if (warp < 4) {
    __syncwarp();
    const int total = p.m_tiles * p.n_tiles * p.splits;
    if (worker < total) {
      const int u = (total - worker + workers - 1) / workers - 1;  // this worker's last unit
      const int tile = (worker + u * workers) / p.splits;
      const int n_idx = tile % p.n_tiles, m_idx = tile / p.n_tiles;
      const int row0 = m_idx * rows_per_tile + static_cast<int>(rank) * GEMM_BLOCK_M + warp * 32;
      gemm_last_unit_helper<PAIR, EPI>(tmem_base, u & 1, row0, n_idx * block_n, p.N, block_n, p, &tmap_c, smem, tlast_bar,
                                       warp, lane);
    }
  }

  gemm_teardown<PAIR>(tmem_base, warp);
#ifdef LYCO_GEMM_TRACE
  if (threadIdx.x == 0) {
    LYCO_TRACE(40, clock64());
    LYCO_TRACE(41, trace_globaltimer());
  }
#endif
}

}  // namespace lyco
