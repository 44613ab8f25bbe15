// C-ABI entry points (include/lyco_b200.h) and the host-side launch logic.
// No torch types cross this boundary; nothing here synchronises the host with the device.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "../../include/lyco_b200.h"
#include "conv_sm100.cuh"
#include "gemm_sm100.cuh"
#include "weight_kernels.cuh"
#include "layout_kernels.cuh"
#include "lokr_struct_kernels.cuh"
#include "dora_kernels.cuh"
#include "hada_sm100.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define LYCO_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct DeviceInfo {
  int device = -1;
  int sms = 0;
  int cc_major = 0;
};
// per-device cache; the current device is queried on every call (cheap) so that multi-GPU
// processes and device switches stay correct
int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  // The tensor-map encoders are DRIVER entry points and need a current context on the calling thread.  A thread that
  // has made no runtime call yet (autograd's backward worker when an engine node is the FIRST node it runs) has none:
  // cuTensorMapEncodeTiled then fails with CUDA_ERROR_INVALID_CONTEXT (201).  cudaFree(0) binds the primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    LYCO_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  int dev = 0;
  LYCO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail("device index %d out of range", dev);
  if (cache[dev].device != dev) {
    DeviceInfo d;
    d.device = dev;
    LYCO_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    LYCO_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    cache[dev] = d;
  }
  *out = cache[dev];
  if (out->cc_major != 10)
    return fail("lycoris_b200 kernels need compute capability 10.x (B200); device %d is %d.x", dev,
                out->cc_major);
  return 0;
}

// 2-D tiled tensor map over a row-major [outer, inner] array of 16-bit elements, 128B swizzle
int make_tmap(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
              uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu ld=%llu box=%ux%u",
                static_cast<int>(r), base, (unsigned long long)inner, (unsigned long long)outer,
                (unsigned long long)ld_elems, box_inner, box_outer);
  return 0;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeIm2colFn encode_im2col() {
  static EncodeIm2colFn fn = []() -> EncodeIm2colFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeIm2colFn>(p);
  }();
  return fn;
}

// im2col tensor map over an NHWC tensor: boxes of `pixels` output positions x 64 channels, 128B swizzle
int make_tmap_im2col(CUtensorMap* m, const void* base, int Nb, int H, int W, int C, int R, int S, int pad_h,
                     int pad_w, int stride, uint32_t pixels) {
  EncodeIm2colFn enc = encode_im2col();
  if (!enc) return fail("cuTensorMapEncodeIm2col is not available from the driver");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(Nb)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                        static_cast<cuuint64_t>(H) * W * C * 2};
  int lower[2] = {-pad_w, -pad_h};                      // {W, H} order (as CUTLASS passes them)
  int upper[2] = {pad_w - (S - 1), pad_h - (R - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, lower, upper, 64,
                   pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d R=%d S=%d pad=%d,%d stride=%d pixels=%u",
                static_cast<int>(r), Nb, H, W, C, R, S, pad_h, pad_w, stride, pixels);
  return 0;
}

// C tensor map for the TMA-store epilogue: [M rows, N inner] 16-bit, boxes of 32 rows x 32 columns, 64B swizzle
int make_tmap_c(CUtensorMap* m, const void* base, uint64_t N, uint64_t M, uint64_t ld_elems) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[2] = {N, M};
  cuuint64_t gstr[1] = {ld_elems * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled (C) failed (%d): base=%p N=%llu M=%llu ld=%llu", static_cast<int>(r), base,
                (unsigned long long)N, (unsigned long long)M, (unsigned long long)ld_elems);
  return 0;
}

// fp32 C tensor map (TMA store / TMA reduce-add epilogue): boxes of 32 rows x 16 columns = 64-byte rows, 64B swizzle
int make_tmap_c_f32(CUtensorMap* m, const void* base, uint64_t N, uint64_t M, uint64_t ld_elems) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_elems & 3))
    return fail("fp32 output needs a 16-byte aligned base and a row pitch that is a multiple of 4 elements");
  cuuint64_t gdim[2] = {N, M};
  cuuint64_t gstr[1] = {ld_elems * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled (C, fp32) failed (%d): base=%p N=%llu M=%llu ld=%llu", static_cast<int>(r), base,
                (unsigned long long)N, (unsigned long long)M, (unsigned long long)ld_elems);
  return 0;
}

// C tensor map for the NCHW epilogue of the convolution: (pixel, channel, image), boxes of 32 pixels x 32 channels
int make_tmap_c_nchw(CUtensorMap* m, const void* base, uint64_t PQ, uint64_t O, uint64_t Nb) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[3] = {PQ, O, Nb};
  cuuint64_t gstr[2] = {PQ * 2, O * PQ * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled (C, NCHW) failed (%d): base=%p PQ=%llu O=%llu Nb=%llu", static_cast<int>(r),
                base, (unsigned long long)PQ, (unsigned long long)O, (unsigned long long)Nb);
  return 0;
}

inline bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("LYCO_PDL"); return !(e && *e == '0'); }();
  return on;
}

// cluster shape + programmatic dependent launch (csrc/pdl.cuh) for the tensor-core kernels
inline void persistent_attrs(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, bool pair) {
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg->numAttrs = 2;
  }
}

template <typename Kern, typename Params>
int launch_persistent(Kern kern, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                      const Params& p, int grid, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(lyco::GEMM_THREADS);
  cfg.dynamicSmemBytes = lyco::GEMM_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  persistent_attrs(&cfg, attr, pair);
  LYCO_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <typename Kern>
int ensure_smem(Kern kern, bool* configured) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    LYCO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lyco::GEMM_SMEM_BYTES));
    configured[dev & 63] = true;
  }
  return 0;
}

// second operand pair of the dual-operand GEMM (C = A·Bᵀ + A2·B2ᵀ); plain GEMMs pass the first pair again (unused)
struct SidePair {
  const CUtensorMap* ta2;
  const CUtensorMap* tb2;
};

template <bool PAIR, bool A_MN, bool B_MN, int EPI>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const lyco::GemmParams& p,
                int grid, cudaStream_t stream, SidePair side) {
  auto kern = lyco::gemm_sm100_kernel<PAIR, A_MN, B_MN, EPI>;
  static bool configured[64] = {};
  if (ensure_smem(kern, configured)) return 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(lyco::GEMM_THREADS);
  cfg.dynamicSmemBytes = lyco::GEMM_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  persistent_attrs(&cfg, attr, PAIR);
  LYCO_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, side.ta2 ? *side.ta2 : ta, side.tb2 ? *side.tb2 : tb, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <bool PAIR>
int dispatch_gemm(bool a_mn, bool b_mn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                  const lyco::GemmParams& p, int grid, cudaStream_t s, SidePair side = SidePair{nullptr, nullptr}) {
  using namespace lyco;
#define LYCO_EPI(AM, BM)                                                                                   \
  do {                                                                                                     \
    if (epi == EPI_STORE16) return launch_gemm<PAIR, AM, BM, EPI_STORE16>(ta, tb, tc, p, grid, s, side);   \
    if (epi == EPI_STORE_F32) return launch_gemm<PAIR, AM, BM, EPI_STORE_F32>(ta, tb, tc, p, grid, s, side); \
    return launch_gemm<PAIR, AM, BM, EPI_ATOMIC_F32>(ta, tb, tc, p, grid, s, side);                        \
  } while (0)
  if (!a_mn && !b_mn) LYCO_EPI(false, false);
  if (!a_mn && b_mn) LYCO_EPI(false, true);
  if (a_mn && b_mn) LYCO_EPI(true, true);
  LYCO_EPI(true, false);
#undef LYCO_EPI
}

template <bool PAIR, int MODE, int EPI>
int launch_conv_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const lyco::ConvParams& cp,
                  int grid, cudaStream_t stream) {
  auto kern = lyco::conv_sm100_kernel<PAIR, MODE, EPI>;
  static bool configured[64] = {};
  if (ensure_smem(kern, configured)) return 1;
  return launch_persistent(kern, PAIR, ta, tb, tc, cp, grid, stream);
}

// `workers` = CTAs (single) or CTA pairs (pair) to launch
template <int MODE, int EPI>
int launch_conv(bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                const lyco::ConvParams& cp, int workers, cudaStream_t stream) {
  if (pair) return launch_conv_t<true, MODE, EPI>(ta, tb, tc, cp, 2 * workers, stream);
  return launch_conv_t<false, MODE, EPI>(ta, tb, tc, cp, workers, stream);
}

inline bool conv_pair_enabled() {
  // LYCO_CONV_PAIR=0 keeps the convolutions on single-CTA tiles (A/B comparisons)
  static const bool on = []() { const char* e = getenv("LYCO_CONV_PAIR"); return !(e && *e == '0'); }();
  return on;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#ifdef LYCO_GEMM_TRACE
unsigned long long* g_trace = nullptr;
#endif

// Relative time of one k-block (4 x tcgen05.mma K = 16) of a tile, per worker.  Measured with tools/gemm_trace.py
// (8192 x 1280 x 1280, tile widths forced): a CTA pair spends 6.58 / 6.61 / 6.70 / 6.72 / 6.85 us on a 20-k-block tile
// at bn = 128 / 160 / 192 / 224 / 256 — the instruction time is nearly FLAT in N (the A operand, 128 rows per CTA, is
// re-read from shared memory for every instruction whatever N is), so a narrower tile buys no time, only more tiles.
// (Round 1 modelled the cost as proportional to bn and picked 224-wide tiles where 256 needs a wave less.)
// A single CTA stages all of B itself: L2 -> SM feed ~(128 + bn) rows per k-block, measured 27 % over the pair at
// 128 x 256.
struct TileChoice {
  bool pair;
  int bn;
};

inline double tile_cost(bool pair, int bn) {
  const double mma = 1.0 + 0.05 * (bn - 128) / 128.0;
  const double feed = pair ? 0.0 : 0.0033 * (128 + bn);
  return (mma > feed ? mma : feed) + 0.05;  // + fixed per-tile overhead (pipeline fill, epilogue hand-off)
}

inline int stages_for(bool pair, int bn) {
  const int stage_bytes = lyco::GEMM_A_BYTES + (pair ? bn / 2 : bn) * lyco::GEMM_BLOCK_K * 2;
  const int s = lyco::GEMM_RING_BYTES / stage_bytes;
  return s > lyco::GEMM_MAX_STAGES ? lyco::GEMM_MAX_STAGES : s;
}

inline bool pair_enabled() {
  // CTA-pair kernels are the default; LYCO_GEMM_PAIR=0 restricts the choice to single-CTA tiles
  static const bool on = []() { const char* e = getenv("LYCO_GEMM_PAIR"); return !(e && *e == '0'); }();
  return on;
}

inline bool tile_valid(bool pair, int bn, bool b_mn) {
  if (bn < 32 || bn > 256 || bn % 32) return false;
  if (b_mn) return pair ? (bn % 128 == 0) : (bn % 64 == 0);  // MN-major B comes in 64-column atoms per CTA
  return !pair || bn >= 64;
}

inline int force_choice(TileChoice* c, bool b_mn) {
  // LYCO_GEMM_FORCE=pair<bn>|single<bn> (experiments / tests)
  static const char* env = getenv("LYCO_GEMM_FORCE");
  if (!env || !*env) return 0;
  TileChoice t;
  if (!strncmp(env, "pair", 4)) { t.pair = true; t.bn = atoi(env + 4); }
  else if (!strncmp(env, "single", 6)) { t.pair = false; t.bn = atoi(env + 6); }
  else return 0;
  if (!tile_valid(t.pair, t.bn, b_mn)) return 0;
  *c = t;
  return 1;
}

// Minimise (waves of tiles) x (cost of one tile step); ties go to the larger tile.
TileChoice pick_tile(int M, int N, int sms, int splits, bool b_mn) {
  TileChoice best{false, b_mn ? 64 : 32};
  if (force_choice(&best, b_mn)) return best;
  double best_cost = 1e30;
  for (int pair = 1; pair >= 0; --pair) {
    if (pair && (M <= 128 || !pair_enabled())) continue;
    for (int bn = 256; bn >= 32; bn -= 32) {
      if (!tile_valid(pair, bn, b_mn)) continue;
      if (bn > 32 && bn - 32 >= N) continue;  // narrower tile already covers N
      const long tiles = static_cast<long>(cdiv(M, pair ? 256 : 128)) * cdiv(N, bn) * (splits > 0 ? splits : 1);
      const long slots = pair ? sms / 2 : sms;
      const long waves = (tiles + slots - 1) / slots;
      const double cost = waves * tile_cost(pair, bn);
      if (cost < best_cost * 0.995) { best_cost = cost; best = TileChoice{pair != 0, bn}; }
    }
  }
  return best;
}

int pick_splits(long tiles, int k_blocks, int sms) {
  int best = 1;
  double best_cost = 1e30;
  const int smax = k_blocks / 4 > 0 ? (k_blocks / 4 < 64 ? k_blocks / 4 : 64) : 1;
  for (int s = 1; s <= smax; ++s) {
    const long units = tiles * s;
    const long waves = (units + sms - 1) / sms;
    // time ~ waves * (k per unit) + epilogue/atomic overhead per unit
    const double cost = waves * (static_cast<double>(k_blocks) / s + 6.0) + 0.5 * s;
    if (cost < best_cost * 0.98) { best_cost = cost; best = s; }
  }
  return best;
}

int check_desc(const lyco_delta_desc_t* d, bool allow_f32_weight = false) {
  if (!d) return fail("null delta descriptor");
  if (d->out_dim <= 0 || d->in_dim <= 0) return fail("bad weight shape %d x %d", d->out_dim, d->in_dim);
  if (d->w_dtype != LYCO_BF16 && d->w_dtype != LYCO_F16 && !(allow_f32_weight && d->w_dtype == LYCO_F32))
    return fail("weight dtype must be bf16 or f16 (got %d)", d->w_dtype);
  switch (d->algo) {
    case LYCO_ALGO_LOCON:
    case LYCO_ALGO_DYLORA:
      if (d->rank <= 0 || !d->f0 || !d->f1) return fail("locon/dylora need rank>0 and f0,f1");
      break;
    case LYCO_ALGO_LOHA:
      if (d->rank <= 0 || !d->f0 || !d->f1 || !d->f2 || !d->f3) return fail("loha needs rank>0 and f0..f3");
      break;
    case LYCO_ALGO_LOKR:
      if (!d->f0 || !d->f1 || d->up <= 0 || d->uq <= 0 || d->vp <= 0 || d->vq <= 0)
        return fail("lokr needs f0,f1 and positive up,uq,vp,vq");
      if (static_cast<int64_t>(d->up) * d->vp != d->out_dim || static_cast<int64_t>(d->uq) * d->vq != d->in_dim)
        return fail("lokr factor shapes (%d,%d)x(%d,%d) do not tile %d x %d", d->up, d->uq, d->vp, d->vq,
                    d->out_dim, d->in_dim);
      break;
    case LYCO_ALGO_RAW:
      if (!d->f0) return fail("raw merge needs f0");
      if (d->f_dtype != LYCO_BF16 && d->f_dtype != LYCO_F16) return fail("raw products must be 16-bit");
      if ((static_cast<int64_t>(d->out_dim) * d->in_dim) % 8) return fail("raw merge needs N*K' %% 8 == 0");
      break;
    case LYCO_ALGO_IA3:
      if (!d->f0) return fail("ia3 needs f0");
      if (d->on_input && (d->ia3_group <= 0 || d->in_dim % d->ia3_group))
        return fail("ia3 on_input: bad group %d for in_dim %d", d->ia3_group, d->in_dim);
      break;
    default:
      return fail("unknown algo %d", d->algo);
  }
  return 0;
}

}  // namespace

extern "C" {

#ifdef LYCO_GEMM_TRACE
// debug build only (tools/gemm_trace.py): device buffer of 64 x u64 per CTA that the next lyco_gemm launches fill
void lyco_debug_set_trace(void* buf) { g_trace = static_cast<unsigned long long*>(buf); }
#endif

int lyco_abi_version(void) { return LYCO_ABI_VERSION; }
const char* lyco_last_error(void) { return g_err; }
uint64_t lyco_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int lyco_device_check(int device) {
  int n = 0;
  LYCO_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail("no CUDA device %d (count %d)", device, n);
  int major = 0;
  LYCO_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail("device %d has compute capability %d.x, need 10.x (B200)", device, major);
  return 0;
}

int lyco_gemm(const void* A, int a_mn_major, int64_t lda, const void* B, int b_mn_major, int64_t ldb,
              void* C, int c_dtype, int64_t ldc, const void* bias, int bias_dtype, int M, int N, int K,
              int ab_dtype, int split_k, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (accumulate && c_dtype != LYCO_F32) return fail("lyco_gemm: accumulate needs an fp32 C");
  if (M <= 0 || N <= 0 || K <= 0) return fail("lyco_gemm: empty problem %d x %d x %d", M, N, K);
  if (!A || !B || !C) return fail("lyco_gemm: null operand");
  if (ab_dtype != LYCO_BF16 && ab_dtype != LYCO_F16) return fail("lyco_gemm: operands must be bf16/f16");
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15)
    return fail("lyco_gemm: operand base pointers must be 16-byte aligned");
  if ((lda | ldb | ldc) & 7) return fail("lyco_gemm: lda/ldb/ldc must be multiples of 8 (got %lld %lld %lld)",
                                          (long long)lda, (long long)ldb, (long long)ldc);
  if (c_dtype != LYCO_F32 && c_dtype != ab_dtype) return fail("lyco_gemm: C must be f32 or the operand dtype");
  if (bias && c_dtype == LYCO_F32) return fail("lyco_gemm: bias only with 16-bit output");
  DeviceInfo di;
  if (device_info(&di)) return 1;

  const int k_blocks = cdiv(K, lyco::GEMM_BLOCK_K);
  const bool a_mn = a_mn_major != 0, b_mn = b_mn_major != 0;
  int splits = 1;
  TileChoice tc = pick_tile(M, N, di.sms, 1, b_mn);
  if (c_dtype == LYCO_F32) {
    // wgrad-like: few output tiles, long reduction -> widest tile, split the reduction across CTAs
    TileChoice forced;
    if (force_choice(&forced, b_mn)) {
      tc = forced;
    } else {
      // widest tile that covers N, as a CTA pair (256 rows) or a single CTA (128 rows): take the variant that pads
      // the [M, N] output less per unit of tile throughput — a 160 x 160 output (the structured LoKr g_w2) is 39 %
      // useful MMA work in a 256 x 256 pair tile, 52 % in two 128 x 192 single-CTA tiles, which also split the
      // reduction over all 148 SMs instead of 74 pairs
      double best = 1e30;
      for (int pr = 1; pr >= 0; --pr) {
        if (pr && (M <= 128 || !pair_enabled())) continue;
        const int step = b_mn ? (pr ? 128 : 64) : 32;
        int bn = 256;
        while (bn > step && bn - step >= N) bn -= step;
        if (!tile_valid(pr, bn, b_mn)) continue;
        const int rows = pr ? 256 : 128;
        const double padded = static_cast<double>(cdiv(M, rows)) * rows * cdiv(N, bn) * bn;
        const double cost = padded * tile_cost(pr, bn) / (128.0 * bn);
        if (cost < best * 0.97) { best = cost; tc.pair = pr != 0; tc.bn = bn; }
      }
      if (best > 1e29) { tc.pair = false; tc.bn = b_mn ? 64 : 32; }
    }
    const long tiles = static_cast<long>(cdiv(M, tc.pair ? 256 : 128)) * cdiv(N, tc.bn);
    splits = split_k > 0 ? split_k : pick_splits(tiles, k_blocks, tc.pair ? di.sms / 2 : di.sms);
    if (splits > k_blocks) splits = k_blocks;
  }
  const int bn = tc.bn;
  const int m_tiles = cdiv(M, tc.pair ? 256 : 128);
  const int n_tiles = cdiv(N, bn);

  CUtensorMap ta, tb, tcm;
  memset(&tcm, 0, sizeof(tcm));
  if (!a_mn) { if (make_tmap(&ta, A, K, M, lda, 64, 128)) return 1; }
  else       { if (make_tmap(&ta, A, M, K, lda, 64, 64)) return 1; }
  if (!b_mn) { if (make_tmap(&tb, B, K, N, ldb, 64, tc.pair ? bn / 2 : bn)) return 1; }
  else       { if (make_tmap(&tb, B, N, K, ldb, 64, 64)) return 1; }
  if (c_dtype != LYCO_F32) { if (make_tmap_c(&tcm, C, N, M, ldc)) return 1; }
  else if (make_tmap_c_f32(&tcm, C, N, M, ldc)) return 1;

  lyco::GemmParams p;
  p.C = C; p.bias = bias; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.splits = splits; p.k_blocks = k_blocks; p.k_blocks1 = k_blocks;
  p.block_n = bn; p.stages = stages_for(tc.pair, bn);
#ifdef LYCO_GEMM_TRACE
  p.trace = g_trace;
#endif
  p.fmt = (ab_dtype == LYCO_BF16) ? 1 : 0;
  p.bias_dtype = bias_dtype;
  p.epi_pq = 0;

  int epi = lyco::EPI_STORE16;
  if (c_dtype == LYCO_F32) {
    epi = (splits > 1 || accumulate) ? lyco::EPI_ATOMIC_F32 : lyco::EPI_STORE_F32;
    if (splits > 1 && !accumulate) {
      if (ldc == N) LYCO_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * static_cast<size_t>(M) * N, stream));
      else LYCO_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, N * sizeof(float), M, stream));
    }
  }
  const long total = static_cast<long>(m_tiles) * n_tiles * splits;
  const long slots = tc.pair ? di.sms / 2 : di.sms;
  const int workers = static_cast<int>(total < slots ? total : slots);
  if (tc.pair) return dispatch_gemm<true>(a_mn, b_mn, epi, ta, tb, tcm, p, 2 * workers, stream);
  return dispatch_gemm<false>(a_mn, b_mn, epi, ta, tb, tcm, p, workers, stream);
}

int lyco_gemm_dual(const void* A, int64_t lda, const void* B, int b_mn_major, int64_t ldb, int K, const void* A2,
                   int64_t lda2, const void* B2, int64_t ldb2, int K2, void* C, int64_t ldc, const void* bias,
                   int bias_dtype, int M, int N, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0 || K2 <= 0) return fail("lyco_gemm_dual: empty problem %d x %d x (%d + %d)", M, N, K, K2);
  if (!A || !B || !A2 || !B2 || !C) return fail("lyco_gemm_dual: null operand");
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_gemm_dual: operands must be bf16/f16");
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(A2) |
       reinterpret_cast<uintptr_t>(B2) | reinterpret_cast<uintptr_t>(C)) & 15)
    return fail("lyco_gemm_dual: operand base pointers must be 16-byte aligned");
  if ((lda | ldb | lda2 | ldb2 | ldc) & 7) return fail("lyco_gemm_dual: leading dimensions must be multiples of 8");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const bool b_mn = b_mn_major != 0;
  const int kb1 = cdiv(K, lyco::GEMM_BLOCK_K), kb2 = cdiv(K2, lyco::GEMM_BLOCK_K);
  const TileChoice tc = pick_tile(M, N, di.sms, 1, b_mn);
  const int bn = tc.bn;
  const int m_tiles = cdiv(M, tc.pair ? 256 : 128), n_tiles = cdiv(N, bn);
  CUtensorMap ta, tb, ta2, tb2, tcm;
  if (make_tmap(&ta, A, K, M, lda, 64, 128)) return 1;
  if (make_tmap(&ta2, A2, K2, M, lda2, 64, 128)) return 1;
  if (!b_mn) {
    if (make_tmap(&tb, B, K, N, ldb, 64, tc.pair ? bn / 2 : bn)) return 1;
    if (make_tmap(&tb2, B2, K2, N, ldb2, 64, tc.pair ? bn / 2 : bn)) return 1;
  } else {
    if (make_tmap(&tb, B, N, K, ldb, 64, 64)) return 1;
    if (make_tmap(&tb2, B2, N, K2, ldb2, 64, 64)) return 1;
  }
  if (make_tmap_c(&tcm, C, N, M, ldc)) return 1;
  lyco::GemmParams p;
  p.C = C; p.bias = bias; p.ldc = ldc; p.M = M; p.N = N; p.K = K + K2;
  p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.splits = 1; p.k_blocks = kb1 + kb2; p.k_blocks1 = kb1;
  p.block_n = bn; p.stages = stages_for(tc.pair, bn);
#ifdef LYCO_GEMM_TRACE
  p.trace = nullptr;
#endif
  p.fmt = (dtype == LYCO_BF16) ? 1 : 0;
  p.bias_dtype = bias_dtype;
  p.epi_pq = 0;
  const long total = static_cast<long>(m_tiles) * n_tiles;
  const long slots = tc.pair ? di.sms / 2 : di.sms;
  const int workers = static_cast<int>(total < slots ? total : slots);
  const SidePair side{&ta2, &tb2};
  if (tc.pair) return dispatch_gemm<true>(false, b_mn, lyco::EPI_STORE16, ta, tb, tcm, p, 2 * workers, stream, side);
  return dispatch_gemm<false>(false, b_mn, lyco::EPI_STORE16, ta, tb, tcm, p, workers, stream, side);
}

int lyco_conv2d_fprop(const void* X, const void* Wk, void* Y, const void* bias, int bias_dtype, int Nb, int H,
                      int W, int C, int O, int R, int S, int pad_h, int pad_w, int stride, int dtype,
                      int y_layout, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !Wk || !Y) return fail("lyco_conv2d_fprop: null operand");
  if (y_layout != LYCO_NHWC && y_layout != LYCO_NCHW && y_layout != LYCO_NCHW_F32)
    return fail("lyco_conv2d_fprop: bad y_layout %d", y_layout);
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_conv2d_fprop: operands must be bf16/f16");
  if (C % 64 || O % 8) return fail("lyco_conv2d_fprop: needs C %% 64 == 0 and O %% 8 == 0 (C=%d O=%d)", C, O);
  if (stride < 1 || stride > 8 || R < 1 || S < 1) return fail("lyco_conv2d_fprop: bad geometry");
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wk) | reinterpret_cast<uintptr_t>(Y)) & 15)
    return fail("lyco_conv2d_fprop: pointers must be 16-byte aligned");
  const int P = (H + 2 * pad_h - R) / stride + 1, Q = (W + 2 * pad_w - S) / stride + 1;
  if (P <= 0 || Q <= 0) return fail("lyco_conv2d_fprop: empty output");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int64_t M64 = static_cast<int64_t>(Nb) * P * Q;
  if (M64 > (1ll << 30)) return fail("lyco_conv2d_fprop: too many output pixels");
  const int M = static_cast<int>(M64), K = R * S * C;
  TileChoice tc = pick_tile(M, O, di.sms, 1, false);
  {
    double best = 1e30;
    for (int pair = 1; pair >= 0; --pair) {
      if (pair && (M <= 128 || !conv_pair_enabled())) continue;
      for (int bn = 256; bn >= 32; bn -= 32) {
        if (!tile_valid(pair, bn, false)) continue;
        if (bn > 32 && bn - 32 >= O) continue;
        const long tiles = static_cast<long>(cdiv(M, pair ? 256 : 128)) * cdiv(O, bn);
        const long slots = pair ? di.sms / 2 : di.sms;
        const long waves = (tiles + slots - 1) / slots;
        const double cost = waves * tile_cost(pair, bn);
        if (cost < best * 0.995) { best = cost; tc.pair = pair != 0; tc.bn = bn; }
      }
    }
  }
  if (const char* e = getenv("LYCO_CONV_BN")) {
    const int f = atoi(e);
    if (tile_valid(tc.pair, f, false)) tc.bn = f;
  }
  const int bn = tc.bn;
  const bool pair = tc.pair;
  CUtensorMap ta, tb, tcm;
  if (make_tmap_im2col(&ta, X, Nb, H, W, C, R, S, pad_h, pad_w, stride, 128)) return 1;
  if (make_tmap(&tb, Wk, K, O, K, 64, pair ? bn / 2 : bn)) return 1;
  if (y_layout != LYCO_NHWC && (P * Q) % 32)
    return fail("lyco_conv2d_fprop: NCHW output needs P*Q %% 32 == 0 (P*Q=%d)", P * Q);
  if (y_layout == LYCO_NCHW) {
    if (make_tmap_c_nchw(&tcm, Y, static_cast<uint64_t>(P) * Q, O, Nb)) return 1;
  } else if (y_layout == LYCO_NCHW_F32) {
    memset(&tcm, 0, sizeof(tcm));  // written with plain coalesced stores
  } else if (make_tmap_c(&tcm, Y, O, M, O)) return 1;
  lyco::ConvParams cp;
  cp.g.epi_pq = P * Q;
  cp.g.C = Y; cp.g.bias = bias; cp.g.ldc = O; cp.g.M = M; cp.g.N = O; cp.g.K = K;
  cp.g.m_tiles = cdiv(M, pair ? 256 : 128); cp.g.n_tiles = cdiv(O, bn); cp.g.splits = 1;
  cp.g.k_blocks = R * S * (C / 64); cp.g.k_blocks1 = cp.g.k_blocks;
  cp.g.block_n = bn; cp.g.stages = stages_for(pair, bn);
#ifdef LYCO_GEMM_TRACE
  cp.g.trace = nullptr;
#endif
  cp.g.fmt = (dtype == LYCO_BF16) ? 1 : 0; cp.g.bias_dtype = bias_dtype;
  cp.PQ = P * Q; cp.Q = Q; cp.C = C; cp.CB = C / 64; cp.S = S; cp.stride = stride;
  cp.low_w = -pad_w; cp.low_h = -pad_h; cp.tiles_per_tap = 1;
  const long total = static_cast<long>(cp.g.m_tiles) * cp.g.n_tiles;
  const long slots = pair ? di.sms / 2 : di.sms;
  const int workers = static_cast<int>(total < slots ? total : slots);
  if (y_layout == LYCO_NCHW)
    return launch_conv<lyco::MODE_FPROP, lyco::EPI_STORE16_NCHW>(pair, ta, tb, tcm, cp, workers, stream);
  if (y_layout == LYCO_NCHW_F32)
    return launch_conv<lyco::MODE_FPROP, lyco::EPI_STORE_F32_NCHW>(pair, ta, tb, tcm, cp, workers, stream);
  return launch_conv<lyco::MODE_FPROP, lyco::EPI_STORE16>(pair, ta, tb, tcm, cp, workers, stream);
}

int lyco_filter_relayout(const void* in, void* out, int O, int C, int taps, int mode, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out) return fail("lyco_filter_relayout: null operand");
  if (O < 1 || C < 1 || taps < 1 || taps > 9) return fail("lyco_filter_relayout: needs 1 <= R*S <= 9 (got %d)", taps);
  if (O > 65535 * 32) return fail("lyco_filter_relayout: too many output channels");
  if (C % 8 || O % 8) return fail("lyco_filter_relayout: needs C %% 8 == 0 and O %% 8 == 0 (C=%d O=%d)", C, O);
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail("lyco_filter_relayout: arrays must be 16-byte aligned");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  if (mode == LYCO_FILTER_WBACK) {
    if (dtype != LYCO_F32) return fail("lyco_filter_relayout: the weight-gradient re-layout is fp32");
    if (O > 65535) return fail("lyco_filter_relayout: too many output channels");
    const dim3 grid(cdiv(C, lyco::FR_CCHUNK), O);
    lyco::filter_row_relayout_kernel<float, true><<<grid, 256, lyco::FR_CCHUNK * taps * 4, stream>>>(
        static_cast<const float*>(in), static_cast<float*>(out), C, taps);
  } else if (mode == LYCO_FILTER_FPROP) {
    if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_filter_relayout: filters are bf16/f16");
    if (O > 65535) return fail("lyco_filter_relayout: too many output channels");
    const dim3 grid(cdiv(C, lyco::FR_CCHUNK), O);
    lyco::filter_row_relayout_kernel<uint16_t, false><<<grid, 256, lyco::FR_CCHUNK * taps * 2, stream>>>(
        static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), C, taps);
  } else if (mode == LYCO_FILTER_DGRAD) {
    if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_filter_relayout: filters are bf16/f16");
    const dim3 grid(cdiv(C, 32), cdiv(O, 32));
    lyco::filter_dgrad_relayout_kernel<<<grid, 256, 32 * (32 * taps + 8) * 2, stream>>>(
        static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), O, C, taps);
  } else {
    return fail("lyco_filter_relayout: bad mode %d", mode);
  }
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_transpose_cast(const void* src, void* dst, int batch, int rows, int cols, int src_dtype, int dst_dtype,
                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst) return fail("lyco_transpose_cast: null operand");
  if (batch < 1 || rows < 1 || cols < 1) return fail("lyco_transpose_cast: empty tensor");
  if (dst_dtype != LYCO_BF16 && dst_dtype != LYCO_F16) return fail("lyco_transpose_cast: dst must be bf16/f16");
  if (src_dtype != LYCO_F32 && src_dtype != dst_dtype)
    return fail("lyco_transpose_cast: src must be f32 or the dst dtype (src=%d dst=%d)", src_dtype, dst_dtype);
  if ((reinterpret_cast<uintptr_t>(src) & 7) || (reinterpret_cast<uintptr_t>(dst) & 3))
    return fail("lyco_transpose_cast: pointers must be 8-byte (src) / 4-byte (dst) aligned");
  const int gx = cdiv(cols, lyco::TR_TILE), gy = cdiv(rows, lyco::TR_TILE);
  if (gy > 65535 || batch > 65535) return fail("lyco_transpose_cast: rows / batch too large for one launch");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const dim3 grid(gx, gy, batch), block(32, 8);
  const int fmt = (dst_dtype == LYCO_BF16) ? 1 : 0;
  if (src_dtype == LYCO_F32)
    lyco::transpose_cast_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(src),
                                                                  static_cast<uint16_t*>(dst), rows, cols, fmt);
  else
    lyco::transpose_cast_kernel<uint16_t><<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(src),
                                                                     static_cast<uint16_t*>(dst), rows, cols, fmt);
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_conv2d_wgrad(const void* X, const void* dY, float* dW, int Nb, int H, int W, int C, int O, int R, int S,
                      int pad_h, int pad_w, int stride, int dtype, int split_k, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !dY || !dW) return fail("lyco_conv2d_wgrad: null operand");
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_conv2d_wgrad: operands must be bf16/f16");
  if (C % 64 || O % 8) return fail("lyco_conv2d_wgrad: needs C %% 64 == 0 and O %% 8 == 0 (C=%d O=%d)", C, O);
  const int P = (H + 2 * pad_h - R) / stride + 1, Q = (W + 2 * pad_w - S) / stride + 1;
  if (P <= 0 || Q <= 0) return fail("lyco_conv2d_wgrad: empty output");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int64_t M64 = static_cast<int64_t>(Nb) * P * Q;
  if (M64 > (1ll << 30)) return fail("lyco_conv2d_wgrad: too many output pixels");
  const int Mpix = static_cast<int>(M64);
  // Tile: N in 64-column im2col atoms inside one filter tap, M = 128 (one CTA) or 256 (CTA pair) output channels;
  // least padding (channels past C in the last N tile of a tap, rows past O in the last M tile) per unit of tile cost.
  int bn = 64;
  bool pair = false;
  double best = 1e30;
  for (int pr = 1; pr >= 0; --pr) {
    if (pr && (O <= 128 || !conv_pair_enabled())) continue;
    for (int c = 256; c >= 64; c -= 64) {
      if (!tile_valid(pr, c, true)) continue;
      const int rows = pr ? 256 : 128;
      const double waste = static_cast<double>(cdiv(C, c) * c) / C * (static_cast<double>(cdiv(O, rows) * rows) / O);
      const double cost = waste * tile_cost(pr, c) / c;
      if (cost < best * 0.995) { best = cost; bn = c; pair = pr != 0; }
    }
  }
  const int tiles_per_tap = cdiv(C, bn);
  const int taps = R * S;
  const int k_blocks = cdiv(Mpix, 64);
  const int m_tiles = cdiv(O, pair ? 256 : 128);
  const long tiles = static_cast<long>(m_tiles) * taps * tiles_per_tap;
  const long slots = pair ? di.sms / 2 : di.sms;
  int splits = split_k > 0 ? split_k : pick_splits(tiles, k_blocks, static_cast<int>(slots));
  if (splits > k_blocks) splits = k_blocks;
  CUtensorMap ta, tb, tcm;
  memset(&tcm, 0, sizeof(tcm));
  if (make_tmap(&ta, dY, O, Mpix, O, 64, 64)) return 1;  // dY [Mpix, O] consumed MN-major
  if (make_tmap_im2col(&tb, X, Nb, H, W, C, R, S, pad_h, pad_w, stride, 64)) return 1;
  lyco::ConvParams cp;
  const int ldw = taps * C;
  if (make_tmap_c_f32(&tcm, dW, ldw, O, ldw)) return 1;
  cp.g.C = dW; cp.g.bias = nullptr; cp.g.ldc = ldw; cp.g.M = O; cp.g.N = ldw; cp.g.K = Mpix;
  cp.g.m_tiles = m_tiles; cp.g.n_tiles = taps * tiles_per_tap; cp.g.splits = splits; cp.g.k_blocks = k_blocks;
  cp.g.k_blocks1 = k_blocks;
  cp.g.block_n = bn; cp.g.stages = stages_for(pair, bn);
#ifdef LYCO_GEMM_TRACE
  cp.g.trace = nullptr;
#endif
  cp.g.fmt = (dtype == LYCO_BF16) ? 1 : 0; cp.g.bias_dtype = 0;
  cp.PQ = P * Q; cp.Q = Q; cp.C = C; cp.CB = C / 64; cp.S = S; cp.stride = stride;
  cp.low_w = -pad_w; cp.low_h = -pad_h; cp.tiles_per_tap = tiles_per_tap;
  if (splits > 1) LYCO_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * static_cast<size_t>(O) * ldw, stream));
  const long total = tiles * splits;
  const int workers = static_cast<int>(total < slots ? total : slots);
  if (splits > 1) return launch_conv<lyco::MODE_WGRAD, lyco::EPI_ATOMIC_F32>(pair, ta, tb, tcm, cp, workers, stream);
  return launch_conv<lyco::MODE_WGRAD, lyco::EPI_STORE_F32>(pair, ta, tb, tcm, cp, workers, stream);
}

int lyco_merge_weight(const lyco_delta_desc_t* d, const void* W, void* W_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (check_desc(d)) return 1;
  if (!W || !W_out) return fail("lyco_merge_weight: null weight pointer");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const uint16_t* w = static_cast<const uint16_t*>(W);
  uint16_t* wo = static_cast<uint16_t*>(W_out);
  const int64_t total = static_cast<int64_t>(d->out_dim) * d->in_dim;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(W_out)) & 15) == 0;
  switch (d->algo) {
    case LYCO_ALGO_LOCON:
    case LYCO_ALGO_DYLORA:
    case LYCO_ALGO_LOHA: {
      if (!aligned16) return fail("lyco_merge_weight: W / W_out must be 16-byte aligned");
      const int grid = cdiv(d->out_dim, lyco::LR_ROWS) * cdiv(d->in_dim, lyco::LR_COLS);
      if (d->algo == LYCO_ALGO_LOHA) lyco::merge_lowrank_kernel<true><<<grid, 256, 0, stream>>>(*d, w, wo);
      else lyco::merge_lowrank_kernel<false><<<grid, 256, 0, stream>>>(*d, w, wo);
      break;
    }
    case LYCO_ALGO_LOKR: {
      const bool vec = aligned16 && (d->vq % 8 == 0);
      const int64_t work = vec ? total / 8 : total;
      int grid = static_cast<int>((work + 255) / 256);
      const int cap = di.sms * 16;
      if (grid > cap) grid = cap;
      // the two regimes of kohya training get dtype-specialised instantiations (same source, constants folded);
      // LYCO_MERGE_GENERIC=1 keeps everything on the run-time-dtype kernel for A/B and bit-exactness checks
      static const bool generic_only = []() { const char* e = getenv("LYCO_MERGE_GENERIC"); return e && *e == '1'; }();
      const bool f32_bf16 = !generic_only && d->f_dtype == LYCO_F32 && d->w_dtype == LYCO_BF16 && !d->pre_round;
      const bool all_bf16 = !generic_only && d->f_dtype == LYCO_BF16 && d->w_dtype == LYCO_BF16 && d->pre_round &&
                            d->pre_dtype == LYCO_BF16;
      // row-blocked kernels (LYCO_MERGE_ROWS=0 keeps the element-indexed ones for the bit-exactness A/B)
      static const bool rows_on = []() { const char* e = getenv("LYCO_MERGE_ROWS"); return !(e && *e == '0'); }();
      const bool rows_ok = rows_on && vec && d->in_dim % 8 == 0 &&
                           (reinterpret_cast<uintptr_t>(d->f1) & 15) == 0;
      int rgrid = (d->out_dim + 7) / 8;  // 8 warps (rows) per CTA
      if (rgrid > cap) rgrid = cap;
      if (rows_ok && f32_bf16)
        lyco::merge_lokr_rows_kernel<LYCO_F32, LYCO_BF16, lyco::PD_NONE><<<rgrid, 256, 0, stream>>>(*d, w, wo);
      else if (rows_ok && all_bf16)
        lyco::merge_lokr_rows_kernel<LYCO_BF16, LYCO_BF16, LYCO_BF16><<<rgrid, 256, 0, stream>>>(*d, w, wo);
      else if (vec && f32_bf16)
        lyco::merge_lokr_kernel<8, LYCO_F32, LYCO_BF16, lyco::PD_NONE><<<grid, 256, 0, stream>>>(*d, w, wo);
      else if (vec && all_bf16)
        lyco::merge_lokr_kernel<8, LYCO_BF16, LYCO_BF16, LYCO_BF16><<<grid, 256, 0, stream>>>(*d, w, wo);
      else if (vec) lyco::merge_lokr_kernel<8><<<grid, 256, 0, stream>>>(*d, w, wo);
      else lyco::merge_lokr_kernel<1><<<grid, 256, 0, stream>>>(*d, w, wo);
      break;
    }
    case LYCO_ALGO_IA3: {
      int grid = static_cast<int>((total + 255) / 256);
      const int cap = di.sms * 16;
      if (grid > cap) grid = cap;
      if (aligned16 && d->in_dim % 8 == 0) {
        int wgrid = cdiv(d->out_dim, 8);
        if (wgrid > cap) wgrid = cap;
        lyco::merge_ia3_vec_kernel<<<wgrid, 256, 0, stream>>>(*d, w, wo);
      } else {
        lyco::merge_ia3_kernel<<<grid, 256, 0, stream>>>(*d, w, wo);
      }
      break;
    }
    case LYCO_ALGO_RAW: {
      if (!aligned16 || (reinterpret_cast<uintptr_t>(d->f0) & 15) || (d->f1 && (reinterpret_cast<uintptr_t>(d->f1) & 15)))
        return fail("lyco_merge_weight(raw): arrays must be 16-byte aligned");
      int grid = static_cast<int>((total / 8 + 255) / 256);
      const int cap = di.sms * 16;
      if (grid > cap) grid = cap;
      static const bool generic_only = []() { const char* e = getenv("LYCO_MERGE_GENERIC"); return e && *e == '1'; }();
      if (!generic_only && d->f_dtype == LYCO_BF16 && d->w_dtype == LYCO_BF16 && d->pre_round &&
          d->pre_dtype == LYCO_BF16)
        lyco::merge_raw_kernel<LYCO_BF16><<<grid, 256, 0, stream>>>(*d, w, wo);
      else
        lyco::merge_raw_kernel<><<<grid, 256, 0, stream>>>(*d, w, wo);
      break;
    }
  }
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Zero up to four gradient arrays: ONE memset when the caller laid them out back to back (each array starting at
// the end of the previous one rounded up to 64 floats — what engine/kernels.py:factor_grads allocates), else one each.
int zero_grads(float* const* g, const size_t* n, int count, cudaStream_t stream) {
  bool packed = true;
  for (int i = 1; i < count; ++i) packed = packed && g[i] == g[i - 1] + ((n[i - 1] + 63) / 64) * 64;
  if (packed && count > 1) {
    LYCO_CUDA(cudaMemsetAsync(g[0], 0, sizeof(float) * static_cast<size_t>(g[count - 1] + n[count - 1] - g[0]), stream));
    return 0;
  }
  for (int i = 0; i < count; ++i) LYCO_CUDA(cudaMemsetAsync(g[i], 0, sizeof(float) * n[i], stream));
  return 0;
}

int lyco_factor_grads(const lyco_delta_desc_t* d, const float* dW, const void* W, float* g0, float* g1,
                      float* g2, float* g3, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (check_desc(d)) return 1;
  if (!dW) return fail("lyco_factor_grads: null dW");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int N = d->out_dim, K = d->in_dim, r = d->rank;
  switch (d->algo) {
    case LYCO_ALGO_LOCON:
    case LYCO_ALGO_DYLORA:
    case LYCO_ALGO_LOHA: {
      const bool loha = d->algo == LYCO_ALGO_LOHA;
      if (!g0 || !g1 || (loha && (!g2 || !g3))) return fail("lyco_factor_grads: missing gradient buffers");
      float* const gs[4] = {g0, g1, g2, g3};
      const size_t ns[4] = {static_cast<size_t>(N) * r, static_cast<size_t>(r) * K, static_cast<size_t>(N) * r,
                            static_cast<size_t>(r) * K};
      if (zero_grads(gs, ns, loha ? 4 : 2, stream)) return 1;
      const int grid = cdiv(N, lyco::LR_ROWS) * cdiv(K, lyco::LR_COLS);
      if (loha) lyco::grad_lowrank_kernel<true><<<grid, 256, 0, stream>>>(*d, dW, g0, g1, g2, g3);
      else lyco::grad_lowrank_kernel<false><<<grid, 256, 0, stream>>>(*d, dW, g0, g1, g2, g3);
      break;
    }
    case LYCO_ALGO_LOKR: {
      if (!g0 || !g1) return fail("lyco_factor_grads: missing gradient buffers");
      const size_t n_w1 = static_cast<size_t>(d->up) * d->uq;
      const int64_t plane = static_cast<int64_t>(d->vp) * d->vq;
      float* const gs[2] = {g0, g1};
      const size_t ns[2] = {n_w1, static_cast<size_t>(plane)};
      if (zero_grads(gs, ns, 2, stream)) return 1;
      if (d->up > 65535) return fail("lyco_factor_grads: lokr up=%d too large", d->up);
      const bool vec4 = (d->vq % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
      const size_t smem = sizeof(float) * static_cast<size_t>(d->uq);
      if (smem > 40 * 1024) return fail("lyco_factor_grads: lokr uq=%d too large", d->uq);
      if (vec4) {
        dim3 grid(static_cast<unsigned>((plane / 4 + 255) / 256), static_cast<unsigned>(d->up));
        lyco::grad_lokr_kernel<4><<<grid, 256, smem, stream>>>(*d, dW, g0, g1);
      } else {
        dim3 grid(static_cast<unsigned>((plane + 255) / 256), static_cast<unsigned>(d->up));
        lyco::grad_lokr_kernel<1><<<grid, 256, smem, stream>>>(*d, dW, g0, g1);
      }
      break;
    }
    case LYCO_ALGO_RAW:
      return fail("lyco_factor_grads: LYCO_ALGO_RAW has no factor gradients (use lyco_grad_prep + lyco_gemm)");
    case LYCO_ALGO_IA3: {
      if (!g0 || !W) return fail("lyco_factor_grads: ia3 needs g0 and W");
      const uint16_t* w = static_cast<const uint16_t*>(W);
      const bool vec = K % 8 == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(dW)) & 15) == 0;
      if (vec && !d->on_input) {
        // g_w[n] = mult * sum_k dW'[n,k] W[n,k]: the same row reduction as DoRA's backward, 16-byte loads, no atomics
        int wgrid = cdiv(N, 8);
        if (wgrid > di.sms * 16) wgrid = di.sms * 16;
        lyco::dora_reduce_rows_vec_kernel<1><<<wgrid, 256, 0, stream>>>(dW, w, g0, N, K, d->w_dtype, d->m_post2);
      } else if (vec) {
        LYCO_CUDA(cudaMemsetAsync(g0, 0, sizeof(float) * static_cast<size_t>(K / d->ia3_group), stream));
        const dim3 cgrid(cdiv(K / 8, 256), cdiv(N, 32));
        lyco::dora_reduce_cols_vec_kernel<1><<<cgrid, 256, 0, stream>>>(dW, w, g0, N, K, d->ia3_group, d->w_dtype, d->m_post2);
      } else if (!d->on_input) {
        lyco::grad_ia3_kernel<<<cdiv(N, 8), 256, 0, stream>>>(*d, dW, w, g0);
      } else {
        LYCO_CUDA(cudaMemsetAsync(g0, 0, sizeof(float) * static_cast<size_t>(K / d->ia3_group), stream));
        dim3 grid(cdiv(K, 256), cdiv(N, 64));
        lyco::grad_ia3_kernel<<<grid, 256, 0, stream>>>(*d, dW, w, g0);
      }
      break;
    }
  }
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_grad_prep(const float* dW, const void* P, void* G, int64_t n, float gscale, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dW || !G || n <= 0 || (n % 8)) return fail("lyco_grad_prep: need dW, G and n %% 8 == 0");
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_grad_prep: output dtype must be 16-bit");
  if ((reinterpret_cast<uintptr_t>(dW) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(P)) & 15)
    return fail("lyco_grad_prep: arrays must be 16-byte aligned");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  int grid = static_cast<int>((n / 8 + 255) / 256);
  const int cap = di.sms * 16;
  if (grid > cap) grid = cap;
  lyco::grad_prep_kernel<<<grid, 256, 0, stream>>>(dW, static_cast<const uint16_t*>(P), static_cast<uint16_t*>(G), n / 8,
                                                   gscale, dtype);
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_hada(int mode, const void* w1a, const void* w1b, const void* w2a, const void* w2b, const void* W, void* out0,
              void* out1, int N, int K, int rank, int dtype, int w_dtype, float m_pre, float m_post1, float m_post2,
              float gscale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (mode != 0 && mode != 1) return fail("lyco_hada: mode must be 0 (merge) or 1 (gradient operands)");
  if (!w1a || !w1b || !w2a || !w2b || !W || !out0 || (mode == 1 && !out1)) return fail("lyco_hada: null operand");
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_hada: factors must be bf16/f16");
  if (mode == 0 && w_dtype != LYCO_BF16 && w_dtype != LYCO_F16) return fail("lyco_hada: weights must be bf16/f16");
  if (N <= 0 || K <= 0 || rank < 8 || rank > 64 || rank % 8 || K % 8)
    return fail("lyco_hada: needs N, K > 0, K %% 8 == 0, rank in 8..64 and a multiple of 8 (N=%d K=%d rank=%d)", N, K, rank);
  if ((reinterpret_cast<uintptr_t>(w1a) | reinterpret_cast<uintptr_t>(w1b) | reinterpret_cast<uintptr_t>(w2a) |
       reinterpret_cast<uintptr_t>(w2b) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out0) |
       reinterpret_cast<uintptr_t>(out1)) & 15)
    return fail("lyco_hada: arrays must be 16-byte aligned");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  CUtensorMap ta1, tb1, ta2, tb2;
  if (make_tmap(&ta1, w1a, rank, N, rank, 64, lyco::HADA_BM)) return 1;   // [N, r]  K-major A
  if (make_tmap(&ta2, w2a, rank, N, rank, 64, lyco::HADA_BM)) return 1;
  if (make_tmap(&tb1, w1b, K, rank, K, 64, 64)) return 1;                 // [r, K'] MN-major B
  if (make_tmap(&tb2, w2b, K, rank, K, 64, 64)) return 1;
  lyco::HadaParams p;
  p.W = W; p.out0 = out0; p.out1 = out1; p.N = N; p.K = K; p.rank = rank;
  p.fmt = dtype == LYCO_BF16 ? 1 : 0; p.w_dtype = w_dtype;
  p.m_pre = m_pre; p.m_post1 = m_post1; p.m_post2 = m_post2; p.gscale = gscale;
  const int grid = cdiv(N, lyco::HADA_BM) * cdiv(K, lyco::HADA_BN);
  static bool configured[2][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (mode == 0) {
    auto kern = lyco::hada_sm100_kernel<0>;
    if (!configured[0][dev & 63]) {
      LYCO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lyco::HADA_SMEM_BYTES));
      configured[0][dev & 63] = true;
    }
    kern<<<grid, lyco::HADA_THREADS, lyco::HADA_SMEM_BYTES, stream>>>(ta1, tb1, ta2, tb2, p);
  } else {
    auto kern = lyco::hada_sm100_kernel<1>;
    if (!configured[1][dev & 63]) {
      LYCO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lyco::HADA_SMEM_BYTES));
      configured[1][dev & 63] = true;
    }
    kern<<<grid, lyco::HADA_THREADS, lyco::HADA_SMEM_BYTES, stream>>>(ta1, tb1, ta2, tb2, p);
  }
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_lokr_mix(const void* in, void* out, const void* w, int w_dtype, int ldw, int transpose, int64_t M, int na,
                  int nb, int nc, int dtype, float* zero_buf, int64_t zero_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || !w) return fail("lyco_lokr_mix: null operand");
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_lokr_mix: activations must be bf16/f16");
  if (M <= 0 || na < 1 || nb < 1 || na > 8 || nb > 8 || nc < 8 || nc % 8)
    return fail("lyco_lokr_mix: needs M > 0, 1 <= na, nb <= 8, nc %% 8 == 0 (M=%lld na=%d nb=%d nc=%d)", (long long)M, na, nb, nc);
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail("lyco_lokr_mix: arrays must be 16-byte aligned");
  if (zero_n < 0 || (zero_n > 0 && !zero_buf)) return fail("lyco_lokr_mix: bad zero-fill request");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int nc8 = nc / 8;
  const int64_t work = M * nc8;
  int64_t grid = (work + 255) / 256;
  const int64_t cap = static_cast<int64_t>(di.sms) * 16;
  if (grid > cap) grid = cap;
  const int fmt = dtype == LYCO_BF16 ? 1 : 0;
  const uint16_t* src = static_cast<const uint16_t*>(in);
  uint16_t* dst = static_cast<uint16_t*>(out);
  if (na <= 4 && nb <= 4)
    lyco::lokr_mix_kernel<4, 4><<<static_cast<int>(grid), 256, 0, stream>>>(src, dst, w, w_dtype, ldw, transpose, M, na, nb, nc8, fmt, zero_buf, zero_n);
  else
    lyco::lokr_mix_kernel<8, 8><<<static_cast<int>(grid), 256, 0, stream>>>(src, dst, w, w_dtype, ldw, transpose, M, na, nb, nc8, fmt, zero_buf, zero_n);
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_lokr_w1grad(const void* P, const void* R, float* g_w1, int64_t M, int na, int nb, int nc, float gscale,
                     int dtype, int zero_fill, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!P || !R || !g_w1) return fail("lyco_lokr_w1grad: null operand");
  if (dtype != LYCO_BF16 && dtype != LYCO_F16) return fail("lyco_lokr_w1grad: activations must be bf16/f16");
  if (M <= 0 || na < 1 || nb < 1 || na > 8 || nb > 8 || nc < 8 || nc % 8)
    return fail("lyco_lokr_w1grad: needs M > 0, 1 <= na, nb <= 8, nc %% 8 == 0 (M=%lld na=%d nb=%d nc=%d)", (long long)M, na, nb, nc);
  if ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(R)) & 15)
    return fail("lyco_lokr_w1grad: arrays must be 16-byte aligned");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  if (zero_fill) LYCO_CUDA(cudaMemsetAsync(g_w1, 0, sizeof(float) * static_cast<size_t>(na) * nb, stream));
  const int nc8 = nc / 8;
  const int64_t work = M * nc8;
  int64_t grid = (work + 255) / 256;
  const int64_t cap = di.sms;  // one resident CTA per SM (150 registers per thread), each thread loops over its items:
                               // the butterfly + 64 global atomics per CTA are paid once per SM, not once per wave
  if (grid > cap) grid = cap;
  const int fmt = dtype == LYCO_BF16 ? 1 : 0;
  const uint16_t* p = static_cast<const uint16_t*>(P);
  const uint16_t* r = static_cast<const uint16_t*>(R);
  if (na <= 4 && nb <= 4)
    lyco::lokr_w1grad_kernel<4, 4><<<static_cast<int>(grid), 256, 0, stream>>>(p, r, g_w1, M, na, nb, nc8, gscale, fmt);
  else
    lyco::lokr_w1grad_kernel<8, 8><<<static_cast<int>(grid), 256, 0, stream>>>(p, r, g_w1, M, na, nb, nc8, gscale, fmt);
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int lyco_delta_weight(const lyco_delta_desc_t* d, const void* W, void* dW_out, int out_dtype, float* norm_sq,
                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (check_desc(d, /*allow_f32_weight=*/true)) return 1;
  if (d->algo == LYCO_ALGO_RAW) return fail("lyco_delta_weight: LYCO_ALGO_RAW is a merge-only descriptor");
  if (d->algo == LYCO_ALGO_IA3 && !W) return fail("lyco_delta_weight: ia3 needs W");
  if (!dW_out && !norm_sq) return fail("lyco_delta_weight: nothing to compute (no output, no norm)");
  if (out_dtype != LYCO_BF16 && out_dtype != LYCO_F16 && out_dtype != LYCO_F32)
    return fail("lyco_delta_weight: bad out_dtype %d", out_dtype);
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int64_t total = static_cast<int64_t>(d->out_dim) * d->in_dim;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(di.sms) * 16;
  if (grid > cap) grid = cap;
  lyco::delta_weight_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(*d, W, dW_out, out_dtype, norm_sq);
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

static int dora_check(const char* who, const void* Wm, int N, int K, int on_out, int taps, int w_dtype) {
  if (!Wm) return fail("%s: null weight", who);
  if (N <= 0 || K <= 0 || taps <= 0 || (!on_out && K % taps)) return fail("%s: bad shape N=%d K=%d taps=%d", who, N, K, taps);
  if (w_dtype != LYCO_BF16 && w_dtype != LYCO_F16) return fail("%s: weights must be bf16/f16", who);
  if ((N + 63) / 64 > 65535) return fail("%s: too many rows", who);
  return 0;
}

int lyco_dora_fwd(const void* Wm, void* W_out, const float* dora_scale, float* sumsq, int N, int K, int on_out,
                  int taps, float mult, float eps, int w_dtype, int scale_dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (dora_check("lyco_dora_fwd", Wm, N, K, on_out, taps, w_dtype)) return 1;
  if (!W_out || !dora_scale || !sumsq) return fail("lyco_dora_fwd: null operand");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int groups = on_out ? N : K / taps;
  const uint16_t* wm = static_cast<const uint16_t*>(Wm);
  const int64_t total = static_cast<int64_t>(N) * K;
  const int64_t cap = static_cast<int64_t>(di.sms) * 16;
  const bool vec = K % 8 == 0 && ((reinterpret_cast<uintptr_t>(Wm) | reinterpret_cast<uintptr_t>(W_out)) & 15) == 0;
  if (vec) {
    int wgrid = static_cast<int>(std::min<int64_t>(cap, cdiv(N, 8)));  // 8 warps = 8 rows per CTA
    if (on_out) {
      lyco::dora_reduce_rows_vec_kernel<0><<<wgrid, 256, 0, stream>>>(nullptr, wm, sumsq, N, K, w_dtype);
    } else {
      LYCO_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * groups, stream));
      const dim3 cgrid(cdiv(K / 8, 256), cdiv(N, 32));
      lyco::dora_reduce_cols_vec_kernel<0><<<cgrid, 256, 0, stream>>>(nullptr, wm, sumsq, N, K, taps, w_dtype);
    }
    lyco::dora_apply_fwd_vec_kernel<<<wgrid, 256, 0, stream>>>(wm, static_cast<uint16_t*>(W_out), sumsq, dora_scale, N, K,
                                                               on_out, taps, mult, eps, w_dtype, scale_dtype);
  } else {
    LYCO_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * groups, stream));
    const dim3 rgrid(cdiv(K, 256), cdiv(N, 64));
    lyco::dora_reduce_kernel<0><<<rgrid, 256, 0, stream>>>(nullptr, wm, sumsq, N, K, on_out, taps, w_dtype);
    int64_t grid = (total + 255) / 256;
    if (grid > cap) grid = cap;
    lyco::dora_apply_fwd_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(wm, static_cast<uint16_t*>(W_out), sumsq, dora_scale,
                                                                           N, K, on_out, taps, mult, eps, w_dtype, scale_dtype);
  }
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

int lyco_dora_bwd(float* dW, const void* Wm, const float* dora_scale, const float* sumsq, float* t, float* g_scale,
                  int N, int K, int on_out, int taps, float mult, float eps, int w_dtype, int scale_dtype,
                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (dora_check("lyco_dora_bwd", Wm, N, K, on_out, taps, w_dtype)) return 1;
  if (!dW || !dora_scale || !sumsq || !t) return fail("lyco_dora_bwd: null operand");
  DeviceInfo di;
  if (device_info(&di)) return 1;
  const int groups = on_out ? N : K / taps;
  const uint16_t* wm = static_cast<const uint16_t*>(Wm);
  const int64_t total = static_cast<int64_t>(N) * K;
  const int64_t cap = static_cast<int64_t>(di.sms) * 16;
  const bool vec = K % 8 == 0 && ((reinterpret_cast<uintptr_t>(Wm) | reinterpret_cast<uintptr_t>(dW)) & 15) == 0;
  if (vec) {
    int wgrid = static_cast<int>(std::min<int64_t>(cap, cdiv(N, 8)));
    if (on_out) {
      lyco::dora_reduce_rows_vec_kernel<1><<<wgrid, 256, 0, stream>>>(dW, wm, t, N, K, w_dtype);
    } else {
      LYCO_CUDA(cudaMemsetAsync(t, 0, sizeof(float) * groups, stream));
      const dim3 cgrid(cdiv(K / 8, 256), cdiv(N, 32));
      lyco::dora_reduce_cols_vec_kernel<1><<<cgrid, 256, 0, stream>>>(dW, wm, t, N, K, taps, w_dtype);
    }
    lyco::dora_apply_bwd_vec_kernel<<<wgrid, 256, 0, stream>>>(dW, wm, sumsq, dora_scale, t, g_scale, N, K, on_out, taps, mult,
                                                               eps, w_dtype, groups, scale_dtype);
  } else {
    LYCO_CUDA(cudaMemsetAsync(t, 0, sizeof(float) * groups, stream));
    const dim3 rgrid(cdiv(K, 256), cdiv(N, 64));
    lyco::dora_reduce_kernel<1><<<rgrid, 256, 0, stream>>>(dW, wm, t, N, K, on_out, taps, w_dtype);
    int64_t grid = (total + 255) / 256;
    if (grid > cap) grid = cap;
    if (grid * 256 < groups) grid = (groups + 255) / 256;  // the first `groups` threads also write g_scale
    lyco::dora_apply_bwd_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(dW, wm, sumsq, dora_scale, t, g_scale, N, K, on_out,
                                                                           taps, mult, eps, w_dtype, groups, scale_dtype);
  }
  LYCO_CUDA(cudaGetLastError());
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
