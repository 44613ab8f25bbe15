/*
 * lyco_b200.h — C ABI of the B200-native LyCORIS adapter-layer engine.
 *
 * Drop-in boundary for ONE hot path of KohakuBlueleaf/LyCORIS: the per-layer
 * delta-weight application (LoCon / LoHa / LoKr / (IA)^3 / DyLoRA) on a wrapped
 * nn.Linear / nn.Conv2d, forward + backward.  The reference has no FFI for this
 * path (it is 100 % PyTorch eager); each entry point below names the reference
 * Python call sites whose ATen launches it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the text of the
 *     last failure on the calling thread is returned by lyco_last_error();
 *   - pointers are raw CUDA device pointers owned by the caller; no tensor
 *     ownership crosses the boundary; the library never allocates device memory
 *     that outlives a call;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no
 *     host synchronisation happens inside any call (CUDA-graph capturable);
 *   - matrices are row-major; "ld" arguments are leading dimensions in ELEMENTS;
 *   - there is NO CPU implementation behind this ABI: calls fail when the device
 *     is not sm_100.
 */
#ifndef LYCO_B200_H_
#define LYCO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LYCO_ABI_VERSION 3

/* element types */
enum { LYCO_BF16 = 0, LYCO_F16 = 1, LYCO_F32 = 2 };
enum { LYCO_NHWC = 0, LYCO_NCHW = 1, LYCO_NCHW_F32 = 2 }; /* layout (and fp32 variant) of the convolution output */
enum { LYCO_FILTER_FPROP = 0, LYCO_FILTER_DGRAD = 1, LYCO_FILTER_WBACK = 2 };

/* adapter algorithms — lycoris/wrapper.py:45-55 network_module_dict keys */
enum {
  LYCO_ALGO_LOCON = 0, /* lycoris/modules/locon.py  */
  LYCO_ALGO_LOHA = 1,  /* lycoris/modules/loha.py   */
  LYCO_ALGO_LOKR = 2,  /* lycoris/modules/lokr.py   */
  LYCO_ALGO_IA3 = 3,   /* lycoris/modules/ia3.py    */
  LYCO_ALGO_DYLORA = 4, /* lycoris/modules/dylora.py (uses the LOCON kernels) */
  LYCO_ALGO_RAW = 5     /* rank-r products already formed by lyco_gemm (K = r): f0 = raw1, f1 = raw2 or NULL,
                           16-bit [N, K'] arrays in f_dtype; merge only (tensor-core path of LOCON/DYLORA/LOHA) */
};

/* ------------------------------------------------------------------------- */
/* library / device                                                           */
/* ------------------------------------------------------------------------- */

int lyco_abi_version(void);
/* thread-local text of the last error (never NULL) */
const char* lyco_last_error(void);
/* 0 when `device` is a compute-capability-10.x part the kernels can run on */
int lyco_device_check(int device);
/* number of kernels this library has launched since load (bench gpu_launches) */
uint64_t lyco_launch_count(void);

/* ------------------------------------------------------------------------- */
/* dense contraction (tcgen05 + TMA)                                          */
/* ------------------------------------------------------------------------- */

/*
 * C[M,N] (+)= A · Bᵀ (+ bias), bf16 (or fp16) operands, fp32 accumulation in TMEM.
 *
 *   A is the [M x K] operand, B the [N x K] operand (K = reduction length).
 *   a_mn_major = 0: A stored row-major [M, K]  (reduction index contiguous)
 *   a_mn_major = 1: A stored row-major [K, M]  (output-row index contiguous)
 *   likewise b_mn_major for B ([N, K] or [K, N]).
 *   lda / ldb: leading dimension (elements) of the array as stored.
 *   c_dtype: LYCO_BF16 / LYCO_F16 (store, optional bias[N]) or LYCO_F32
 *            (store; with split_k > 1 the partials are added into C — fp32 reductions
 *             performed by the TMA unit, in no fixed order — which the call zero-fills first).
 *   ab_dtype: LYCO_BF16 or LYCO_F16.
 *   bias / bias_dtype: optional [N] vector added in the epilogue (NULL = none).
 *   split_k: 0 = choose automatically; otherwise number of reduction splits.
 *   accumulate: fp32 C only — C += A·Bᵀ (fp32 reductions as above, no zero fill).  Lets fp32 layers be contracted
 *            as three bf16 products (hi·hi + hi·lo + lo·hi) into one fp32 accumulator.
 *
 * Replaces, per wrapped layer and step, the ATen library calls at
 *   forward   base + delta contraction  lycoris/modules/locon.py:317,331
 *                                        (same in loha.py:309,321 lokr.py:551,565
 *                                         ia3.py:136,143 dylora.py:150,156)
 *   backward  dX (both branches) and dδ = dYᵀ·X produced by autograd for those.
 * Alignment: base pointers 16-byte aligned, lda/ldb/ldc multiples of 8.
 */
int lyco_gemm(const void* A, int a_mn_major, int64_t lda,
              const void* B, int b_mn_major, int64_t ldb,
              void* C, int c_dtype, int64_t ldc,
              const void* bias, int bias_dtype,
              int M, int N, int K,
              int ab_dtype, int split_k, int accumulate, void* stream);

/*
 * C[M,N] = A·Bᵀ + A2·B2ᵀ (+ bias): two operand pairs accumulated into ONE TMEM accumulator by the same kernel —
 * the skinny rank-r side path of a low-rank adapter next to the dense base contraction, with no merged weight.
 *   A  [M, K]  row-major (K-major),   B  [N, K] (b_mn_major = 0) or [K, N] (b_mn_major = 1)
 *   A2 [M, K2] row-major,             B2 [N, K2] or [K2, N] with the SAME major-ness as B
 *   C 16-bit, same dtype as the operands; K2 is padded to the 64-element k-block by TMA zero fill.
 * LoCon / DyLoRA on nn.Linear:
 *   forward   T = X·downᵀ [M,r] (lyco_gemm),   Y  = X·Wᵀ + T·(s·up)ᵀ + b        (A2 = T,  B2 = s·up [N, r])
 *   backward  U = dY·(s·up) [M,r] (lyco_gemm), dX = dY·W + U·down                (A2 = U,  B2 = down [r, K], MN-major)
 *             g_up = s·dYᵀ·T,  g_down = Uᵀ·X   (skinny lyco_gemm calls) — no dW' = dYᵀ·X, no W' = W + dW pass.
 * Replaces `org_forward(x)` + `make_weight` + `W + dW` + `self.op(x, delta_weight)` + add (lycoris/modules/locon.py:
 * 309-332) in the contraction ORDER of the reference's own bypass path (locon.py:273-307: lora_down, then lora_up).
 */
int lyco_gemm_dual(const void* A, int64_t lda, const void* B, int b_mn_major, int64_t ldb, int K,
                   const void* A2, int64_t lda2, const void* B2, int64_t ldb2, int K2,
                   void* C, int64_t ldc, const void* bias, int bias_dtype, int M, int N, int dtype, void* stream);

/* ------------------------------------------------------------------------- */
/* implicit-GEMM convolution (tcgen05 + TMA im2col)                           */
/* ------------------------------------------------------------------------- */

/*
 * Y[n,p,q,o] = sum_{r,s,c} X[n, p*stride - pad_h + r, q*stride - pad_w + s, c] * Wk[o, r, s, c]  (+ bias[o])
 *
 * X is NHWC (torch channels_last storage) [Nb, H, W, C]; Wk is the merged filter re-laid as
 * [O, R, S, C]; Y is NHWC [Nb, P, Q, O] with P = (H + 2*pad_h - R)/stride + 1 (same for Q).
 * Needs C % 64 == 0 and O % 8 == 0, dilation 1, groups 1.
 * The input-gradient (stride 1) is the same call on dY with the flipped/transposed filter
 * [C, R, S, O] and pad' = R-1-pad.
 * Replaces F.conv2d on base and delta weights lycoris/modules/locon.py:317,331 (conv branch of
 * LycorisBaseModule, base.py:110-124) and autograd's dgrad for both.
 */
int lyco_conv2d_fprop(const void* X, const void* Wk, void* Y, const void* bias, int bias_dtype,
                      int Nb, int H, int W, int C, int O, int R, int S, int pad_h, int pad_w,
                      int stride, int dtype, int y_layout, void* stream);
/*   y_layout: LYCO_NHWC, or LYCO_NCHW — the epilogue writes Y as [Nb, O, P, Q] (PyTorch's default
 *   layout, what F.conv2d returns for an NCHW input) with channel-major TMA stores; needs P*Q % 32 == 0. */

/*   LYCO_NCHW_F32: as LYCO_NCHW with an fp32 Y — the input gradient of a layer whose input was fp32 under
 *   autocast comes back in the input's dtype without a separate cast pass. */

/*
 * Filter re-layouts between PyTorch's [O, C, R, S] and the operand layouts of the convolution kernels
 * (taps = R*S <= 9):
 *   LYCO_FILTER_FPROP  W' [O][C][taps] (16-bit)     -> Wk [O][taps][C]              B operand of lyco_conv2d_fprop
 *   LYCO_FILTER_DGRAD  W' [O][C][taps] (16-bit)     -> Wd [C][taps flipped][O]      B operand of the dgrad call
 *   LYCO_FILTER_WBACK  dWk [O][taps][C] (fp32)      -> dW' [O][C][taps]             what lyco_factor_grads indexes
 * Replaces the permute/flip copies a PyTorch host would make around those calls (no counterpart in the
 * reference: cuDNN consumes [O,C,R,S] directly, F.conv2d at lycoris/modules/locon.py:317,331).
 */
int lyco_filter_relayout(const void* in, void* out, int O, int C, int taps, int mode, int dtype,
                         void* stream);

/*
 * dst[b][c][r] = cast(src[b][r][c]) for b < batch: the activation layout pass in front of the im2col
 * producer.  NCHW -> NHWC is (rows = C, cols = H*W), NHWC -> NCHW is (rows = H*W, cols = C).
 * src_dtype: LYCO_F32 (fused autocast cast, round-to-nearest-even like Tensor.to) or == dst_dtype;
 * dst_dtype: LYCO_BF16 / LYCO_F16.
 * Replaces the `input.to(autocast dtype)` copy autocast inserts in front of F.conv2d
 * (lycoris/modules/locon.py:317,331 run under torch.autocast in kohya training).
 */
int lyco_transpose_cast(const void* src, void* dst, int batch, int rows, int cols, int src_dtype,
                        int dst_dtype, void* stream);

/*
 * dW[o, r, s, c] (fp32, [O, R*S*C]) = sum_{n,p,q} dY[n,p,q,o] * X[n, p*stride - pad_h + r, q*stride - pad_w + s, c]
 * dY is NHWC [Nb, P, Q, O].  Needs C % 64 == 0, O % 8 == 0, dW 16-byte aligned.  split_k as in lyco_gemm.
 * Replaces autograd's weight-gradient of the delta convolution (d(delta_weight), locon.py:331).
 */
int lyco_conv2d_wgrad(const void* X, const void* dY, float* dW, int Nb, int H, int W, int C, int O,
                      int R, int S, int pad_h, int pad_w, int stride, int dtype, int split_k,
                      void* stream);

/* ------------------------------------------------------------------------- */
/* weight-side kernels (HBM-bound)                                            */
/* ------------------------------------------------------------------------- */

/*
 * Describes ΔW of one adapter on one layer.  Weights are handled as row-major
 * [out_dim, in_dim] with in_dim = in_channels*kh*kw for convolutions (the
 * reference flattens the same way: locon.py:207, loha.py:76, lokr.py:131-135).
 *
 * Rounding chain (reproduces the reference's rounding points):
 *   raw   = factor product in fp32
 *   raw   = rnd_pre(raw)                      if pre_round (product dtype = 16-bit)
 *   d     = rnd_pre(raw * m_pre)
 *   d     = rnd_w(d)                          cast to the base-weight dtype
 *   d     = rnd_w(d * m_post1); d = rnd_w(d * m_post2)
 *   W'    = rnd_w(W + d)
 * rnd_pre rounds to `pre_dtype` when pre_round != 0, rnd_w rounds to w_dtype.
 */
typedef struct lyco_delta_desc {
  int32_t algo;      /* LYCO_ALGO_*                                          */
  int32_t out_dim;   /* N                                                     */
  int32_t in_dim;    /* K' = in_channels * kh * kw                            */
  int32_t rank;      /* r  (LOCON/LOHA/DYLORA)                                */
  int32_t up, uq;    /* LOKR: w1 is [up, uq]   (lokr.py:161-163)              */
  int32_t vp, vq;    /* LOKR: w2 is [vp, vq]   (vq includes kh*kw)            */
  int32_t on_input;  /* IA3: scale input channels (ia3.py:92-97)              */
  int32_t ia3_group; /* IA3 on_input: elements per input channel (kh*kw)      */
  int32_t f_dtype;   /* dtype of the factor arrays (LYCO_BF16/F16/F32)        */
  int32_t w_dtype;   /* dtype of W and W' (LYCO_BF16/F16)                     */
  int32_t pre_round; /* see rounding chain                                    */
  int32_t pre_dtype; /* dtype rnd_pre rounds to                               */
  float m_in;        /* DYLORA: multiplier folded into `down` (dylora.py:117) */
  float m_pre, m_post1, m_post2;
  /* factor arrays:
   *   LOCON/DYLORA f0 = up   [N, r]        f1 = down [r, K']
   *   LOHA         f0 = w1a  [N, r]        f1 = w1b  [r, K']
   *                f2 = w2a  [N, r]        f3 = w2b  [r, K']
   *   LOKR         f0 = w1   [up, uq]      f1 = w2   [vp, vq]
   *   IA3          f0 = w    [N] or [K'/ia3_group]                            */
  const void* f0;
  const void* f1;
  const void* f2;
  const void* f3;
} lyco_delta_desc_t;

/*
 * W_out[N,K'] = W + ΔW  (rounded as described above).  One pass over W.
 * Replaces make_weight/get_weight + `.to(dtype)*scale` + `W + ΔW*mult`
 *   lycoris/modules/locon.py:198-219,321-328   loha.py:194-226,310-318
 *   lycoris/modules/lokr.py:358-381,552-562 (+ functional/lokr.py:11-20 kron)
 *   lycoris/modules/ia3.py:91-102,137-141      dylora.py:97-117,295-298
 * and removes `delta_weight = new_weight - base_weight` (locon.py:330 …).
 */
int lyco_merge_weight(const lyco_delta_desc_t* d, const void* W, void* W_out,
                      void* stream);

/*
 * Factor gradients from dW' (fp32 [N,K'], = dYᵀ·X of the merged contraction).
 * g0..g3 receive fp32 gradients with the shapes of f0..f3 (unused ones NULL);
 * they are zero-filled by the call (one memset when the arrays lie back to back, each starting at the end of
 * the previous one rounded up to 64 floats; otherwise one per array).  `W` is needed by IA3 only.
 * Replaces autograd through the chain listed at lyco_merge_weight, incl.
 * HadaWeight.backward lycoris/functional/loha.py:18-30 and torch.kron backward.
 */
int lyco_factor_grads(const lyco_delta_desc_t* d, const float* dW, const void* W,
                      float* g0, float* g1, float* g2, float* g3, void* stream);

/*
 * G[i] = rnd16(gscale * dW[i] * (P ? P[i] : 1)) for i < n (n % 8 == 0): the 16-bit operand of the skinny
 * gradient contractions on the tensor-core path (g_up = G·downᵀ, g_down = upᵀ·G; LoHa: P = the other
 * Hadamard factor, recomputed — lycoris/functional/loha.py:18-30).
 */
int lyco_grad_prep(const float* dW, const void* P, void* G, int64_t n, float gscale, int dtype, void* stream);

/* ------------------------------------------------------------------------- */
/* LoHa: factor products, Hadamard product and merge in one tensor-core kernel */
/* ------------------------------------------------------------------------- */

/*
 * P1 = w1a [N, r] · w1b [r, K'],  P2 = w2a [N, r] · w2b [r, K']  per 128 x 128 weight tile on tcgen05 (two TMEM
 * accumulators; the factors are staged by TMA, r <= 64 is one k-block), then in the epilogue
 *   mode 0:  W_out = rnd_w(W + chain(rnd(P1) * rnd(P2)))      W, out0 16-bit [N, K'] (w_dtype); chain multipliers
 *            m_pre, m_post1, m_post2 as in lyco_delta_desc_t with the product domain = `dtype`
 *   mode 1:  out0 = G1 = rnd(gscale * dW * rnd(P2)),  out1 = G2 = rnd(gscale * dW * rnd(P1))      W = dW' fp32 [N, K']
 * so the [N, K'] products never travel through HBM (forward 4 instead of 12 bytes per weight element, backward 8
 * instead of 20).  Factors in `dtype` (bf16 / f16), r % 8 == 0, 8 <= r <= 64, K' % 8 == 0, 16-byte aligned arrays.
 * Replaces HadaWeight.forward and the re-products of HadaWeight.backward, lycoris/functional/loha.py:10-30, plus
 * `.to(dtype) * scalar`, `W + dW * mult` of lycoris/modules/loha.py:310-318.
 */
int lyco_hada(int mode, const void* w1a, const void* w1b, const void* w2a, const void* w2b, const void* W,
              void* out0, void* out1, int N, int K, int rank, int dtype, int w_dtype, float m_pre, float m_post1,
              float m_post2, float gscale, void* stream);

/* ------------------------------------------------------------------------- */
/* structured LoKr factor gradients (no dense dW')                            */
/* ------------------------------------------------------------------------- */

/*
 * out[m, a, c] = sum_b Wm(a, b) * in[m, b, c]   for m < M;  in is [M, nb, nc], out is [M, na, nc] (16-bit, `dtype`),
 * Wm(a, b) = w[a*ldw + b] (transpose = 0) or w[b*ldw + a] (transpose = 1), w in w_dtype (bf16 / f16 / f32).
 * nc % 8 == 0, na, nb <= 8, 16-byte aligned arrays.
 *
 * With w = lokr_w1 [up, uq] this forms Xt[m,pu,v] = sum_u w1[pu,u] X[m,u,v] from X [M, uq*vq] (transpose = 0) or
 * Z[m,u,pv] = sum_pu w1[pu,u] dY[m,pu,pv] from dY [M, up*vp] (transpose = 1): the channel-group mix of the
 * reference's structured Kronecker contraction, lycoris/modules/lokr.py:517-530 (F.linear over the group axis).
 * Then  g_w2 = lyco_gemm(dY as [M*up, vp]ᵀ, Xt as [M*up, vq])  — one contraction with 1/uq of the FLOPs of the
 * dense dW' = dYᵀ·X that autograd runs for `delta_weight` (lokr.py:565) — and g_w1 comes from lyco_lokr_w1grad.
 * zero_buf / zero_n: optional fp32 array the kernel also zero-fills (the gradient buffers the following
 * accumulate-mode lyco_gemm and lyco_lokr_w1grad reduce into), so no memset node sits between the kernels.
 */
int lyco_lokr_mix(const void* in, void* out, const void* w, int w_dtype, int ldw, int transpose,
                  int64_t M, int na, int nb, int nc, int dtype, float* zero_buf, int64_t zero_n, void* stream);

/*
 * g_w1[a, b] = gscale * sum_{m, c} P[m, a, c] * R[m, b, c];  P is [M, na, nc], R is [M, nb, nc] (16-bit, `dtype`),
 * g_w1 fp32 [na, nb], zero-filled by the call when zero_fill != 0 (else it must already be zero: the sums are
 * added with atomics).  nc % 8 == 0, na, nb <= 8.
 * With Q = dY2·w2 ([M*up, vq], one lyco_gemm) and X [M, uq, vq]:  g_w1[pu,u] = sum_{m,v} Q[m,pu,v] X[m,u,v]
 * (or P = dY, R = H = X2·w2ᵀ when dY is the mixed side).  Replaces torch.kron's backward reduction over the dense
 * d(delta_weight) for lokr_w1 (lycoris/functional/lokr.py:11-20 under autograd).
 */
int lyco_lokr_w1grad(const void* P, const void* R, float* g_w1, int64_t M, int na, int nb, int nc,
                     float gscale, int dtype, int zero_fill, void* stream);

/* ------------------------------------------------------------------------- */
/* delta weight / DoRA (merge_to, onfly_merge, apply_max_norm, dora_wd)       */
/* ------------------------------------------------------------------------- */

/*
 * dW[N,K'] = the adapter's delta weight (the rounding chain above with W left out), written in out_dtype
 * (LYCO_BF16 / LYCO_F16 / LYCO_F32; dW_out may be NULL when only the norm is wanted), and, when norm_sq != NULL,
 * *norm_sq += sum dW^2 (fp32; the caller zero-fills it).  Here d->w_dtype may be LYCO_F32: no 16-bit rounding
 * points, i.e. the arithmetic of the reference's get_diff_weight on fp32 parameters.  `W` is read by IA3 only
 * (dW = W * w * mult), in d->w_dtype.
 * Replaces get_diff_weight / make_weight(...).norm() on the merge and max-norm paths:
 *   lycoris/modules/base.py:326-374 (merge_to / onfly_merge), locon.py:221-237,262-271, loha.py:228-260,
 *   lokr.py:383-397,442-466, ia3.py:91-111, kohya.py:589-613 (apply_max_norm_regularization).
 */
int lyco_delta_weight(const lyco_delta_desc_t* d, const void* W, void* dW_out, int out_dtype, float* norm_sq,
                      void* stream);

/*
 * DoRA (dora_wd) around the merged weight Wm = W + dW (16-bit [N,K'], from lyco_merge_weight):
 *   sumsq[g] = sum over group g of Wm^2      groups: output rows (on_out = 1) or input channels (K'/taps of them)
 *   s[g]     = dora_scale[g] / (sqrt(sumsq[g]) + eps);   s <- mult*(s - 1) + 1 when mult != 1
 *   W_out    = rnd_w(Wm * s[group])
 * scale_dtype: the dtype the reference evaluates the norm / quotient / product in (dora_scale's dtype: LYCO_F32
 * under autocast; LYCO_BF16 / LYCO_F16 for a 16-bit adapter, whose roundings of the norm and of the scale are then
 * reproduced); eps is that dtype's machine epsilon, passed by the caller.
 * sumsq (fp32 [groups]) is zero-filled and written by the call and kept by the caller for the backward.
 * Replaces apply_weight_decompose, lycoris/modules/locon.py:239-260 (copies in loha.py / lokr.py).
 */
int lyco_dora_fwd(const void* Wm, void* W_out, const float* dora_scale, float* sumsq, int N, int K, int on_out,
                  int taps, float mult, float eps, int w_dtype, int scale_dtype, void* stream);

/*
 * Backward of lyco_dora_fwd, in place: dW (fp32 [N,K'], = dYᵀ·X of the contraction with W_out) becomes dWm, and
 * g_scale[g] (fp32 [groups]) receives the gradient of dora_scale:
 *   t[g] = sum_g dW*Wm;  n = sqrt(sumsq), ne = n + eps
 *   g_scale[g] = t * mult / ne;     dWm = s * dW - (mult * dora_scale * t / (ne^2 * n)) * Wm
 * `t` is fp32 scratch [groups] (zero-filled by the call).  Replaces autograd through apply_weight_decompose.
 */
int lyco_dora_bwd(float* dW, const void* Wm, const float* dora_scale, const float* sumsq, float* t,
                  float* g_scale, int N, int K, int on_out, int taps, float mult, float eps, int w_dtype,
                  int scale_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LYCO_B200_H_ */
