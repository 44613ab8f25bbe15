"""The fused adapter layer: autograd functions that replace the reference's rebuild-mode forward
(lycoris/modules/locon.py:317-332 and its siblings) with

    forward   W' = merge(W, factors)                 one HBM pass      (lyco_merge_weight)
              Y  = X · W'ᵀ + b                        one contraction   (lyco_gemm, tcgen05)
    backward  dX  = dY · W'                           one contraction
              dW' = dYᵀ · X   (fp32, split over M)    one contraction
              d(factors) from dW'                     one HBM pass      (lyco_factor_grads)

i.e. 1 + 2 dense contractions where the reference runs 2 + 3, no N×K temporaries besides W'
itself, and ~6 launches per layer-step instead of 35–60 ATen ops.
"""

from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from ..logging import warning_once
from . import kernels as K
from ._lib import EngineUnavailable

_HALF = (torch.bfloat16, torch.float16)


@dataclass
class NativeSpec:
    """What the merge / gradient kernels need to know about one adapter (see lyco_delta_desc_t)."""

    algo: int
    factors: tuple
    rank: int = 0
    up: int = 0
    uq: int = 0
    vp: int = 0
    vq: int = 0
    on_input: int = 0
    ia3_group: int = 1
    matmul_product: bool = True  # factor product is an autocast-eligible matmul (locon/loha/dylora)
    m_in: float = 1.0
    m_pre: float = 1.0
    m_post1: float = 1.0
    m_post2: float = 1.0
    grad_mask: tuple = field(default_factory=tuple)  # optional per-factor "needs grad" override
    # DoRA (dora_wd): (dora_scale parameter, wd_on_out, multiplier) — the merged weight is rescaled by lyco_dora_fwd;
    # the merge itself then runs WITHOUT the multiplier (m_post2 = 1), which DoRA applies to the scale instead
    dora: tuple = None


def _autocast_dtype():
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return None


def _build_desc(spec: NativeSpec, factors, w_dtype, ac_dtype, out_dim, in_dim):
    fdt = factors[0].dtype
    prod_dtype = ac_dtype if (ac_dtype is not None and spec.matmul_product) else fdt
    pre_round = prod_dtype in _HALF
    return K.make_desc(
        spec.algo, out_dim, in_dim, factors=factors, w_dtype=w_dtype, rank=spec.rank,
        up=spec.up, uq=spec.uq, vp=spec.vp, vq=spec.vq, on_input=spec.on_input, ia3_group=spec.ia3_group,
        pre_round=int(pre_round), pre_dtype=prod_dtype if pre_round else w_dtype,
        m_in=spec.m_in, m_pre=spec.m_pre, m_post1=spec.m_post1, m_post2=spec.m_post2,
    )


def _uniform_factors(factors):
    """The kernels take one factor dtype per layer; mixed dtypes promote to fp32 (exact)."""
    dts = {f.dtype for f in factors}
    if len(dts) == 1:
        return tuple(f.contiguous() for f in factors)
    return tuple(f.float().contiguous() for f in factors)


# --------------------------------------------------------------------------- dense
def _dense_nt(a, b, bias=None):
    """a[M,K] · b[N,K]ᵀ (+ bias) — forward contraction.  The output [M, N] has row pitch N, so N must be
    TMA-addressable too (lyco_gemm rejects ldc % 8 != 0); an empty batch has nothing to launch."""
    if a.shape[0] > 0 and b.shape[0] % 8 == 0 and K.gemm_supported(a, b):
        return K.gemm(a, b, bias=bias)
    if a.shape[0] > 0:
        warning_once("lycoris_b200: a layer shape is not TMA-addressable (needs multiples of 8); "
                     "that layer's contraction uses the library GEMM")
    return F.linear(a, b, bias)


def _dense_nn(a, b):
    """a[M,N] · b[N,K] — dgrad contraction (b consumed MN-major, no transpose copy)."""
    if a.shape[0] > 0 and b.shape[1] % 8 == 0 and K.gemm_supported(a, b):
        return K.gemm(a, b, b_mn=True)
    return a @ b


def _dense_tn_f32(a, b):
    """a[M,N]ᵀ · b[M,K] -> fp32 [N,K] — wgrad contraction, reduction split across CTAs."""
    if a.shape[0] > 0 and b.shape[1] % 8 == 0 and K.gemm_supported(a, b):
        return K.gemm(a, b, a_mn=True, b_mn=True, out_dtype=torch.float32)
    return (a.t() @ b).float()


# fp32 layers (no autocast): every operand is split x = hi + lo into two bf16 tensors and the contraction is
# hi·hi + hi·lo + lo·hi accumulated in ONE fp32 C on the tcgen05 GEMM (the lo·lo term is below fp32 rounding of
# the sum); relative error ~2^-16 per product instead of bf16's 2^-8.
def _split_bf16(t):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def _dense3(a, b, *, a_mn=False, b_mn=False, init=None):
    """fp32 C = A·Bᵀ (+ init) from fp32 A, B as three bf16 products; shapes as in K.gemm."""
    ah, al = _split_bf16(a)
    bh, bl = _split_bf16(b)
    M = a.shape[1] if a_mn else a.shape[0]
    N = b.shape[1] if b_mn else b.shape[0]
    out = torch.zeros((M, N), device=a.device, dtype=torch.float32) if init is None else init
    for x, y in ((ah, bh), (ah, bl), (al, bh)):
        K.gemm(x, y, a_mn=a_mn, b_mn=b_mn, out=out, out_dtype=torch.float32, accumulate=True)
    return out


def _f32_supported(*mats):
    """fp32 Linear on the engine: both dims of W are a row pitch of some operand or output
    (W[N,K]: K for X / W, N for Y / dY), so both must be multiples of 8."""
    return all(t.dim() == 2 and t.is_contiguous() and t.shape[0] % 8 == 0 and t.shape[1] % 8 == 0
               and t.dtype == torch.float32 for t in mats)


class _MergedContractionF32(torch.autograd.Function):
    """fp32 Linear on the engine: y = x·Wm^T + b with Wm assembled in fp32 by PyTorch ops."""

    @staticmethod
    def forward(ctx, x, Wm, bias):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        Wc = Wm.contiguous()
        init = None if bias is None else bias.float().expand(x2.shape[0], Wc.shape[0]).contiguous()
        y = _dense3(x2, Wc, init=init).view(*x.shape[:-1], Wc.shape[0])
        ctx.save_for_backward(x, Wm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wm = ctx.saved_tensors
        dy2 = dy.reshape(-1, Wm.shape[0]).float().contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _dense3(dy2, Wm.contiguous(), b_mn=True).view(x.shape)
        if ctx.needs_input_grad[1]:
            dw = _dense3(dy2, x.reshape(-1, x.shape[-1]).contiguous(), a_mn=True, b_mn=True)
        return dx, dw, None


# LYCO_CONV_IMPL=cudnn routes every k>1 convolution through aten/cuDNN (on the merged weight) instead of
# the implicit-GEMM kernels; the default uses the engine whenever the layer is eligible (NHWC, C % 64 == 0).
_CONV_ENGINE = os.environ.get("LYCO_CONV_IMPL", "engine") != "cudnn"


def _conv_engine_ok(x, w, cp, dtype=None):
    return (_CONV_ENGINE and w.dim() == 4 and not isinstance(cp["padding"], str)
            and K.conv2d_supported(x, w.shape, cp["stride"], cp["padding"], cp["dilation"], cp["groups"], dtype))


def _conv_forward(x, w, bias, cp):
    """(y, x_saved, info) — on the engine the input goes through ONE layout pass (NCHW -> NHWC fused with the
    autocast cast; nothing if it already is NHWC) and the output comes back in the input's layout, written
    directly by the epilogue.  ``x_saved`` (NHWC, 16-bit) is what the weight-gradient kernel reads again."""
    if _conv_engine_ok(x, w, cp, w.dtype):
        O, C, R, S = w.shape
        nchw_in = not x.is_contiguous(memory_format=torch.channels_last)
        xs = K.as_nhwc(x, w.dtype)
        wk = K.filter_relayout(w.contiguous(), K.FILTER_FPROP)  # [O, R*S*C]: taps outermost, channels contiguous
        y = K.conv2d_fprop(xs, wk, bias, R, S, cp["padding"], cp["stride"][0], out_nchw=nchw_in)
        return y, xs, (True, nchw_in, x.dtype)
    xd = x.dtype
    if x.dtype != w.dtype:
        x = x.to(w.dtype)
    y = torch.ops.aten.convolution(x, w, bias, cp["stride"], cp["padding"], cp["dilation"], False,
                                   [0] * len(cp["stride"]), cp["groups"])
    return y, x, (False, False, xd)


def _conv_backward(dy, x, w, cp, info, need_x, need_w):
    """(dx, dw) with dw shaped like ``w`` (fp32 from the engine, w.dtype from the library path); dx in the
    layout and dtype of the forward input."""
    engine, nchw_in, x_dtype = info
    dx = dw = None
    if engine:
        O, C, R, S = w.shape
        st, pad = cp["stride"][0], cp["padding"]
        dyc = K.as_nhwc(dy, w.dtype)
        if need_x:
            if st == 1 and O % 64 == 0 and C % 8 == 0 and pad[0] <= R - 1 and pad[1] <= S - 1:
                # input gradient = the same implicit GEMM over dY with the flipped, transposed filter
                wd = K.filter_relayout(w.contiguous(), K.FILTER_DGRAD)  # [C, R*S*O], taps flipped
                dx = K.conv2d_fprop(dyc, wd, None, R, S, (R - 1 - pad[0], S - 1 - pad[1]), 1, out_nchw=nchw_in,
                                    out_dtype=x_dtype)
            else:
                dx = torch.ops.aten.convolution_backward(
                    dyc, x, w, None, cp["stride"], cp["padding"], cp["dilation"], False, [0, 0], cp["groups"],
                    [True, False, False])[0]
        if need_w:
            dwk = K.conv2d_wgrad(x, dyc, R, S, pad, st)  # [O, R*S*C] fp32
            dw = K.filter_relayout((dwk, (O, C, R, S)), K.FILTER_WBACK)
    else:
        dx, dw, _ = torch.ops.aten.convolution_backward(
            dy.contiguous(), x, w, None, cp["stride"], cp["padding"], cp["dilation"], False,
            [0] * len(cp["stride"]), cp["groups"], [need_x, need_w, False])
    if dx is not None and dx.dtype != x_dtype:
        dx = dx.to(x_dtype)
    return dx, dw


def _is_pointwise(cp, w):
    """1x1 / stride 1 / no padding convolution == linear over channels."""
    return (w.dim() == 4 and w.shape[2] == 1 and w.shape[3] == 1 and tuple(cp["stride"]) == (1, 1)
            and tuple(cp["padding"]) == (0, 0) and cp["groups"] == 1)


# ------------------------------------------------------- low-rank products on the tensor cores
# LoCon / DyLoRA / LoHa: when the factor product is a 16-bit matmul in the reference (bf16/fp16 adapters or
# autocast) and r % 8 == 0, the rank-r products and their gradients run on the tcgen05 GEMM (K = r, or
# N = r) instead of the CUDA-core tile kernels; LYCO_LOWRANK=simt forces the latter.
_LOWRANK_TC = os.environ.get("LYCO_LOWRANK", "tc") != "simt"
_LOWRANK_ALGOS = (K.ALGO_LOCON, K.ALGO_DYLORA, K.ALGO_LOHA)


def _lowrank_tc_dtype(spec, factors, out_dim, in_dim, ac_dtype):
    """dtype of the tensor-core product path, or None when the layer has to use the SIMT kernels."""
    if not _LOWRANK_TC or spec.algo not in _LOWRANK_ALGOS:
        return None
    prod = ac_dtype if ac_dtype is not None else factors[0].dtype
    if prod not in _HALF or spec.rank % 8 or out_dim % 8 or in_dim % 8:
        return None
    return prod


def _lowrank_operands(spec, factors, prod):
    """16-bit copies of the factors as the matmul sees them (DyLoRA folds alpha/(b+1)*mult into `down`
    in the parameter dtype first, dylora.py:117)."""
    ops = [f.to(prod) for f in factors]
    if spec.algo == K.ALGO_DYLORA and spec.m_in != 1.0:
        ops[1] = (factors[1] * spec.m_in).to(prod)
    return [o.contiguous() for o in ops]


# LoHa: one tensor-core kernel forms both factor products of a tile, multiplies them and merges (hada_sm100.cuh);
# LYCO_LOHA=split keeps the round-1 sequence (two product GEMMs writing [N, K'] arrays + merge_raw / grad_prep).
_LOHA_FUSED = os.environ.get("LYCO_LOHA", "fused") != "split"


def _hada_ok(spec, f, W):
    return (_LOHA_FUSED and spec.algo == K.ALGO_LOHA and 8 <= spec.rank <= 64 and spec.rank % 8 == 0
            and W.numel() // W.shape[0] % 8 == 0 and f[0].dtype == W.dtype)


def _lowrank_merge(spec, factors, W, prod, out_dim, in_dim, keep=None):
    f = _lowrank_operands(spec, factors, prod)
    if keep is not None:
        keep.extend(f)  # the 16-bit operand copies are reused by backward (4 cast kernels less per LoHa layer-step)
    if _hada_ok(spec, f, W):
        return K.hada_merge(f, W.view(out_dim, in_dim), spec.m_pre, spec.m_post1, spec.m_post2).view(W.shape)
    raws = [K.gemm(f[0], f[1], b_mn=True)]  # [N, r] x [r, K'] -> [N, K'], rounded to `prod` like the reference's matmul
    if spec.algo == K.ALGO_LOHA:
        raws.append(K.gemm(f[2], f[3], b_mn=True))
    desc = K.make_desc(K.ALGO_RAW, out_dim, in_dim, factors=raws, w_dtype=W.dtype, pre_round=1, pre_dtype=prod,
                       m_pre=spec.m_pre, m_post1=spec.m_post1, m_post2=spec.m_post2)
    return K.merge_weight(desc, W)


def _zeroed_views(shapes, device):
    """fp32 tensors of ``shapes`` carved out of ONE zero-filled buffer (64-float aligned): the split-K contractions
    that follow ADD into them (accumulate mode), so a layer pays one fill instead of one memset per gradient."""
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 63) // 64 * 64
    flat = torch.zeros(total, device=device, dtype=torch.float32)
    return [flat[o:o + n].view(s) for o, n, s in zip(offs, sizes, shapes)]


def _lowrank_grads(spec, factors, dWm, prod, f=None):
    """Factor gradients from fp32 dW' with skinny tensor-core contractions (fp32 outputs)."""
    if not f:
        f = _lowrank_operands(spec, factors, prod)
    gscale = spec.m_pre * spec.m_post1 * spec.m_post2
    f32 = torch.float32
    N, r = f[0].shape
    Kp = f[1].shape[1]
    if spec.algo != K.ALGO_LOHA:
        G = K.grad_prep(dWm, None, gscale, prod)
        g_up, g_down = _zeroed_views([(N, r), (r, Kp)], dWm.device)
        K.gemm(G, f[1], out=g_up, out_dtype=f32, accumulate=True)                          # G · downᵀ    [N, r]
        K.gemm(f[0], G, a_mn=True, b_mn=True, out=g_down, out_dtype=f32, accumulate=True)  # upᵀ · G      [r, K']
        if spec.algo == K.ALGO_DYLORA and spec.m_in != 1.0:
            g_down = g_down * spec.m_in
        return [g_up, g_down]
    if _LOHA_FUSED and 8 <= r <= 64 and Kp % 8 == 0:
        G1, G2 = K.hada_grad_operands(f, dWm, gscale)  # P1, P2 re-formed per tile on the tensor cores, never stored
    else:
        P1 = K.gemm(f[0], f[1], b_mn=True)  # recomputed, never cached (functional/loha.py:18-30)
        P2 = K.gemm(f[2], f[3], b_mn=True)
        G1 = K.grad_prep(dWm, P2, gscale, prod)
        G2 = K.grad_prep(dWm, P1, gscale, prod)
    g = _zeroed_views([(N, r), (r, Kp), (N, r), (r, Kp)], dWm.device)
    K.gemm(G1, f[1], out=g[0], out_dtype=f32, accumulate=True)
    K.gemm(f[0], G1, a_mn=True, b_mn=True, out=g[1], out_dtype=f32, accumulate=True)
    K.gemm(G2, f[3], out=g[2], out_dtype=f32, accumulate=True)
    K.gemm(f[2], G2, a_mn=True, b_mn=True, out=g[3], out_dtype=f32, accumulate=True)
    return g


# ------------------------------------------------------- LoCon / DyLoRA side path (no merged weight)
# y = x·Wᵀ + (x·downᵀ)·(s·up)ᵀ + b with BOTH products in one TMEM accumulator (lyco_gemm_dual): the adapter's rank-r
# side path rides in the base contraction, W' is never formed, and backward needs no dense dW' — 2 dense contractions
# per layer-step, the algorithmic minimum, instead of 3.  Contraction order = the reference's own bypass
# path (locon.py:273-307); rebuild-mode rounding differs by the snap of dW onto W's grid.
# LYCO_LOCON=merged keeps the round-1 merged-weight path.
_LOCON_SIDE = os.environ.get("LYCO_LOCON", "side") != "merged"


def _locon_side_ok(spec, conv, delta_only, W, x2, cdt):
    return (_LOCON_SIDE and spec.algo in (K.ALGO_LOCON, K.ALGO_DYLORA) and conv is None and not delta_only
            and spec.dora is None and spec.rank % 8 == 0 and 8 <= spec.rank <= 256
            and W.dim() == 2 and W.shape[0] % 8 == 0 and W.shape[1] % 8 == 0
            and x2.shape[0] > 0 and cdt in _HALF and K.gemm_supported(x2, W))


class _LoconSidePath(torch.autograd.Function):
    """Linear LoCon / DyLoRA without a merged weight (see _LOCON_SIDE)."""

    @staticmethod
    def forward(ctx, x, W, bias, spec, up, down):
        cdt = W.dtype
        s = spec.m_pre * spec.m_post1 * spec.m_post2
        down_c = (down * spec.m_in if (spec.algo == K.ALGO_DYLORA and spec.m_in != 1.0) else down).to(cdt).contiguous()
        up_s = (up * s if s != 1.0 else up).to(cdt).contiguous()            # [N, r]: the scale chain folded into `up`
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        T = K.gemm(x2, down_c)                                               # [M, r] = x·downᵀ
        y = K.gemm_dual(x2, W, T, up_s, bias=bias).view(*x.shape[:-1], W.shape[0])
        ctx.save_for_backward(x, W, T, up_s, down_c)
        ctx.spec, ctx.scale, ctx.f_dtypes = spec, s, (up.dtype, down.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, T, up_s, down_c = ctx.saved_tensors
        spec = ctx.spec
        N = W.shape[0]
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        need_x = ctx.needs_input_grad[0]
        need_up, need_down = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        dx = g_up = g_down = None
        f32 = torch.float32
        U = K.gemm(dy2, up_s, b_mn=True) if (need_x or need_down) else None  # [M, r] = dY·(s·up)
        if need_x:
            dx = K.gemm_dual(dy2, W, U, down_c, b_mn=True).view(x.shape)     # dY·W + U·down
        r = T.shape[1]
        zu, zd = _zeroed_views([(N, r), (r, W.shape[1])], dy2.device) if (need_up or need_down) else (None, None)
        if need_up:
            g_up = K.gemm(dy2, T, a_mn=True, b_mn=True, out=zu, out_dtype=f32, accumulate=True)   # dYᵀ·T   [N, r]
            if ctx.scale != 1.0:
                g_up = g_up * ctx.scale
        if need_down:
            x2 = x.reshape(-1, x.shape[-1])
            if not x2.is_contiguous():
                x2 = x2.contiguous()
            g_down = K.gemm(U, x2, a_mn=True, b_mn=True, out=zd, out_dtype=f32, accumulate=True)  # Uᵀ·X    [r, K]
            if spec.algo == K.ALGO_DYLORA and spec.m_in != 1.0:
                g_down = g_down * spec.m_in
        if g_up is not None and g_up.dtype != ctx.f_dtypes[0]:
            g_up = g_up.to(ctx.f_dtypes[0])
        if g_down is not None and g_down.dtype != ctx.f_dtypes[1]:
            g_down = g_down.to(ctx.f_dtypes[1])
        return dx, None, None, None, g_up, g_down


def _locon_conv_skinny_ok(spec, conv, info, w):
    """LoCon / DyLoRA on an engine-run k x k convolution: factor gradients without the dense dW' (see below)."""
    return (_LOCON_SIDE and conv is not None and spec.algo in (K.ALGO_LOCON, K.ALGO_DYLORA) and spec.dora is None
            and info[0] and w.dim() == 4 and spec.rank % 8 == 0 and 8 <= spec.rank <= 256
            and w.shape[2] * w.shape[3] <= 9 and w.shape[1] % 64 == 0 and w.shape[0] % 8 == 0)


def _locon_conv_skinny_grads(spec, factors, dy, xs, w_shape, conv):
    """[g_up, g_down] (fp32) of a LoCon / DyLoRA convolution from two skinny convolutions and two skinny GEMMs — the
    bypass-order contraction (lora_down k x k conv, then lora_up 1 x 1; locon.py:273-307) applied to the GRADIENTS:
        T = conv(X, down)  [pixels, r]        U = dY . (s up)  [pixels, r]
        g_up = s . dY^T T  [O, r]             g_down = wgrad(X, U)  [r, C k k]
    instead of the dense dW' = wgrad(X, dY) [O, C k k] + one reduction pass over it: 2 r / O of the dense FLOPs."""
    O, C, R, S = w_shape
    up, down = factors                      # [O, r], [r, C*R*S]
    r = up.shape[1]
    cdt = xs.dtype
    s_ = spec.m_pre * spec.m_post1 * spec.m_post2
    down_c = (down * spec.m_in if (spec.algo == K.ALGO_DYLORA and spec.m_in != 1.0) else down).to(cdt)
    up_s = (up * s_ if s_ != 1.0 else up).to(cdt).contiguous()
    st, pad = conv["stride"][0], conv["padding"]
    down_k = K.filter_relayout(down_c.reshape(r, C, R, S).contiguous(), K.FILTER_FPROP)       # [r, R*S*C]
    T = K.conv2d_fprop(xs, down_k, None, R, S, pad, st)                                       # NHWC [Nb, r, P, Q]
    dyc = K.as_nhwc(dy, cdt)
    Nb, _, P, Q = dyc.shape
    dY2 = dyc.permute(0, 2, 3, 1).reshape(Nb * P * Q, O)
    T2 = T.permute(0, 2, 3, 1).reshape(Nb * P * Q, r)
    U2 = K.gemm(dY2, up_s, b_mn=True)                                                        # [pixels, r]
    g_up, = _zeroed_views([(O, r)], dy.device)
    K.gemm(dY2, T2, a_mn=True, b_mn=True, out=g_up, out_dtype=torch.float32, accumulate=True)
    if s_ != 1.0:
        g_up = g_up * s_
    U4 = U2.view(Nb, P, Q, r).permute(0, 3, 1, 2)                                            # NHWC storage
    g_down_k = K.conv2d_wgrad(xs, U4, R, S, pad, st)                                         # [r, R*S*C] fp32
    g_down = K.filter_relayout((g_down_k, (r, C, R, S)), K.FILTER_WBACK).reshape(r, C * R * S)
    if spec.algo == K.ALGO_DYLORA and spec.m_in != 1.0:
        g_down = g_down * spec.m_in
    return [g_up, g_down]


# ------------------------------------------------------- structured LoKr factor gradients
# dW = kron(w1, w2): g_w1 / g_w2 from two skinny contractions (1/uq of the dense FLOPs each) instead of the dense
# fp32 dW' = dYᵀ·X + a reduction pass over it (lokr_struct_kernels.cuh).  The structured form trades FLOPs for HBM
# passes (≈4 over the smaller activation + 1 over the larger), so it is used where that pays — measured on B200:
# layers with max(N, K) >= 4·min(N, K) (the feed-forward projections: 166 -> ~100 us at 10240x1280, M = 8192); square
# layers (1280², 640²) stay on the dense tensor-core wgrad, which is faster there (45 vs ~55 us).
# LYCO_LOKR_GRAD = auto (default) | all (every eligible layer) | dense (round-1 path everywhere).
_LOKR_STRUCT = os.environ.get("LYCO_LOKR_GRAD", "auto")
_LOKR_STRUCT_ASPECT = 4


def _lokr_structured_ok(spec, conv, x2, dy2):
    mode = _LOKR_STRUCT
    if mode in (False, "dense") or spec.algo != K.ALGO_LOKR:
        return False
    if mode not in (True, "all"):
        n_out, n_in = spec.up * spec.vp, spec.uq * spec.vq
        if max(n_out, n_in) < _LOKR_STRUCT_ASPECT * min(n_out, n_in):
            return False
    return (conv is None and spec.dora is None
            and 1 <= spec.up <= 8 and 1 <= spec.uq <= 8 and spec.vp % 8 == 0 and spec.vq % 8 == 0
            and x2.shape[0] > 0 and x2.dtype in _HALF and dy2.dtype == x2.dtype
            and x2.is_contiguous() and dy2.is_contiguous() and x2.data_ptr() % 16 == 0 and dy2.data_ptr() % 16 == 0)


def _lokr_structured_grads(spec, factors, dy2, x2):
    """[g_w1, g_w2] (fp32) for dW = kron(w1 [up,uq], w2 [vp,vq]) from dY [M, up*vp] and X [M, uq*vq]."""
    w1, w2 = factors
    up, uq, vp, vq = spec.up, spec.uq, spec.vp, spec.vq
    M = x2.shape[0]
    cdt = x2.dtype
    gscale = spec.m_pre * spec.m_post1 * spec.m_post2
    w2c = w2 if w2.dtype == cdt else w2.to(cdt)
    f32 = torch.float32
    # one fp32 buffer for both gradients, zero-filled by the mix kernel on the side (no memset nodes in the graph):
    # g_w1 (up*uq floats, padded to 64) followed by g_w2 — the skinny GEMM and the w1 reduction ADD into it
    n1 = (up * uq + 63) // 64 * 64
    flat = torch.empty(n1 + vp * vq, device=x2.device, dtype=f32)
    g_w1 = flat[: up * uq].view(up, uq)
    g_w2 = flat[n1:].view(vp, vq)
    if up * vq <= uq * vp:
        # mix the (smaller) input side:  Xt[m,pu,v] = sum_u w1[pu,u] X[m,u,v]
        Xt = K.lokr_mix(x2, w1, up, uq, vq, transpose=False, zero=flat)
        dY2 = dy2.view(M * up, vp)
        K.gemm(dY2, Xt.view(M * up, vq), a_mn=True, b_mn=True, out=g_w2, out_dtype=f32, accumulate=True)   # [vp, vq]
        Q = K.gemm(dY2, w2c.t().contiguous())                                            # dY2 · w2   [M*up, vq]
        K.lokr_w1grad(Q.view(M, up * vq), x2, up, uq, vq, gscale, out=g_w1)
    else:
        # mix the output-gradient side:  Z[m,u,pv] = sum_pu w1[pu,u] dY[m,pu,pv]
        Z = K.lokr_mix(dy2, w1, uq, up, vp, transpose=True, zero=flat)
        X2 = x2.view(M * uq, vq)
        K.gemm(Z.view(M * uq, vp), X2, a_mn=True, b_mn=True, out=g_w2, out_dtype=f32, accumulate=True)      # [vp, vq]
        H = K.gemm(X2, w2c.contiguous())                                                 # X2 · w2ᵀ   [M*uq, vp]
        K.lokr_w1grad(dy2, H.view(M, uq * vp), up, uq, vp, gscale, out=g_w1)
    if gscale != 1.0:
        g_w2 = g_w2 * gscale
    return [g_w1, g_w2]


# ------------------------------------------------------------------ autograd nodes
class _AdapterContraction(torch.autograd.Function):
    """y = op(x, merge(W, factors), bias) with everything on the sm_100a kernels."""

    @staticmethod
    def _merge(spec, factors_u, W, ac_dtype, out_dim, in_dim, keep=None):
        prod = _lowrank_tc_dtype(spec, factors_u, out_dim, in_dim, ac_dtype)
        if prod is not None:
            return _lowrank_merge(spec, factors_u, W, prod, out_dim, in_dim, keep)
        return K.merge_weight(_build_desc(spec, factors_u, W.dtype, ac_dtype, out_dim, in_dim), W)

    @staticmethod
    def _dora_args(spec, W):
        g, on_out, mult = spec.dora
        taps = 1
        for d_ in W.shape[2:]:
            taps *= d_
        # eps is that of the dtype the reference computes the norm in (= dora_scale's: fp32, or bf16 for a bf16 adapter)
        return (g.detach().reshape(-1).float().contiguous(), bool(on_out), taps, float(mult), float(torch.finfo(g.dtype).eps),
                g.dtype if g.dtype in _HALF else torch.float32)

    @staticmethod
    def forward(ctx, x, W, bias, spec, conv, delta_only, ac_dtype, dora_scale, *factors):
        factors_u = _uniform_factors(factors)
        out_dim = W.shape[0]
        in_dim = W.numel() // out_dim
        ctx.lowrank_ops = []
        Wm = _AdapterContraction._merge(spec, factors_u, W, ac_dtype, out_dim, in_dim, ctx.lowrank_ops)
        sumsq = None
        if spec.dora is not None:
            g32, on_out, taps, mult, eps, sdt = _AdapterContraction._dora_args(spec, W)
            Wm, sumsq = K.dora_fwd(Wm, g32, on_out, taps, mult, eps, sdt)  # the pre-rescale W + dW is recomputed in backward
        if delta_only:
            Wm = Wm - W  # exact on W's grid: this is the reference's `new_weight - base_weight`
        if conv is None:
            x2 = x.reshape(-1, x.shape[-1])
            if not x2.is_contiguous():
                x2 = x2.contiguous()
            y = _dense_nt(x2, Wm, bias).view(*x.shape[:-1], out_dim)
        else:
            y, x, ctx.conv_info = _conv_forward(x, Wm, bias, conv)
        ctx.save_for_backward(x, W, Wm, sumsq, *factors)
        ctx.spec, ctx.conv, ctx.ac_dtype = spec, conv, ac_dtype
        ctx.dims = (out_dim, in_dim)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, Wm, sumsq, *factors = ctx.saved_tensors
        spec, conv = ctx.spec, ctx.conv
        out_dim, in_dim = ctx.dims
        NF = 8  # index of the first factor among forward's arguments
        need_x = ctx.needs_input_grad[0]
        need_f = any(ctx.needs_input_grad[NF:]) or ctx.needs_input_grad[NF - 1]
        dx = None
        dWm = None
        gs = None
        if conv is None:
            dy2 = dy.reshape(-1, out_dim)
            if not dy2.is_contiguous():
                dy2 = dy2.contiguous()
            if need_x:
                dx = _dense_nn(dy2, Wm).view(x.shape)
            if need_f:
                x2 = x.reshape(-1, x.shape[-1])
                if not x2.is_contiguous():
                    x2 = x2.contiguous()
                if _lokr_structured_ok(spec, conv, x2, dy2):
                    gs = _lokr_structured_grads(spec, _uniform_factors(factors), dy2, x2)
                else:
                    dWm = _dense_tn_f32(dy2, x2)
        elif need_f and _locon_conv_skinny_ok(spec, conv, ctx.conv_info, Wm):
            dx, _ = _conv_backward(dy, x, Wm, conv, ctx.conv_info, need_x, False)
            gs = _locon_conv_skinny_grads(spec, _uniform_factors(factors), dy, x, Wm.shape, conv)
        else:
            dx, dw = _conv_backward(dy, x, Wm, conv, ctx.conv_info, need_x, need_f)
            if need_f:
                dWm = dw.reshape(out_dim, in_dim).float()
        grads = [None] * len(factors)
        g_dora = None
        if need_f:
            factors_u = _uniform_factors(factors)
            if gs is None:
                dWm = dWm.contiguous()
                if spec.dora is not None:
                    # gradient of W'' -> gradient of W + dW (in place) and of dora_scale
                    g32, on_out, taps, mult, eps, sdt = _AdapterContraction._dora_args(spec, W)
                    Wpre = _AdapterContraction._merge(spec, factors_u, W, ctx.ac_dtype, out_dim, in_dim)
                    g_dora = K.dora_bwd(dWm, Wpre, g32, sumsq, on_out, taps, mult, eps,
                                        want_scale_grad=ctx.needs_input_grad[NF - 1], scale_dtype=sdt)
                    if g_dora is not None:
                        ds = spec.dora[0]
                        g_dora = g_dora.view(ds.shape).to(ds.dtype)
                prod = _lowrank_tc_dtype(spec, factors_u, out_dim, in_dim, ctx.ac_dtype)
                if prod is not None:
                    gs = _lowrank_grads(spec, factors_u, dWm, prod, ctx.lowrank_ops)
                else:
                    desc = _build_desc(spec, factors_u, W.dtype, ctx.ac_dtype, out_dim, in_dim)
                    gs = K.factor_grads(desc, dWm, W, [f.shape for f in factors_u])
            for i, (g, f) in enumerate(zip(gs, factors)):
                if ctx.needs_input_grad[NF + i]:
                    grads[i] = g.to(f.dtype) if g.dtype != f.dtype else g
        return (dx, None, None, None, None, None, None, g_dora, *grads)


class _MergedContraction(torch.autograd.Function):
    """y = op(x, Wm, bias) where Wm was assembled by PyTorch ops (DoRA / Tucker / scalar / dropout
    variants): only the three dense contractions run in the engine; dWm goes back to autograd."""

    @staticmethod
    def forward(ctx, x, Wm, bias, conv):
        if conv is None:
            x2 = x.reshape(-1, x.shape[-1])
            if not x2.is_contiguous():
                x2 = x2.contiguous()
            y = _dense_nt(x2, Wm.contiguous(), bias).view(*x.shape[:-1], Wm.shape[0])
        else:
            y, x, ctx.conv_info = _conv_forward(x, Wm, bias, conv)
        ctx.save_for_backward(x, Wm)
        ctx.conv = conv
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wm = ctx.saved_tensors
        conv = ctx.conv
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = dw = None
        if conv is None:
            dy2 = dy.reshape(-1, Wm.shape[0])
            if not dy2.is_contiguous():
                dy2 = dy2.contiguous()
            if need_x:
                dx = _dense_nn(dy2, Wm.contiguous()).view(x.shape)
            if need_w:
                x2 = x.reshape(-1, x.shape[-1])
                if not x2.is_contiguous():
                    x2 = x2.contiguous()
                dw = _dense_tn_f32(dy2, x2).to(Wm.dtype)
        else:
            dx, dw = _conv_backward(dy, x, Wm, conv, ctx.conv_info, need_x, need_w)
            if dw is not None and dw.dtype != Wm.dtype:
                dw = dw.to(Wm.dtype)
        return dx, dw, None, None


# ------------------------------------------------------------------- cold paths
def delta_weight(spec, m_pre, m_post1, m_post2, shape, want_out=True, want_norm=False):
    """``(dW, sum dW^2)`` of one adapter through lyco_delta_weight, in the FACTOR dtype with the rounding points of
    the reference's get_diff_weight / get_weight (torch ops in the parameter dtype): merge_to, onfly_merge,
    apply_max_norm (reference base.py:326-374, lokr.py:383-397, 442-466) without assembling kron / up@down / the
    Hadamard product with PyTorch ops.  The multipliers replace the training chain's (m_pre, m_post1, m_post2)."""
    factors = _uniform_factors(tuple(f.detach() for f in spec.factors))
    fdt = factors[0].dtype
    out_dim = shape[0]
    in_dim = 1
    for d_ in shape[1:]:
        in_dim *= d_
    half = fdt in _HALF
    desc = K.make_desc(
        spec.algo, out_dim, in_dim, factors=factors, w_dtype=fdt, rank=spec.rank, up=spec.up, uq=spec.uq, vp=spec.vp,
        vq=spec.vq, on_input=spec.on_input, ia3_group=spec.ia3_group, pre_round=int(half), pre_dtype=fdt,
        m_in=spec.m_in, m_pre=m_pre, m_post1=m_post1, m_post2=m_post2)
    with torch.cuda.device(factors[0].device):
        return K.delta_weight(desc, tuple(shape), fdt, None, want_out, want_norm)


# --------------------------------------------------------------------- dispatcher
def _conv_params(module):
    kw = module.kw_dict
    return {
        "stride": list(kw["stride"]),
        "padding": list(kw["padding"]) if not isinstance(kw["padding"], str) else kw["padding"],
        "dilation": list(kw["dilation"]),
        "groups": kw["groups"],
    }


def _base_trainable(org):
    """The base layer itself is being fine-tuned (joint training, a trainable bias): its gradient has to come
    from ``org_forward`` like in the reference, so the engine must not swallow the base contraction."""
    return org.weight.requires_grad or (org.bias is not None and org.bias.requires_grad)


def adapter_forward(module, x, args, kwargs, native_spec, assemble_fallback):
    """Rebuild-mode forward of one adapter on the engine (called from ``Module.forward``)."""
    if not x.is_cuda:
        raise EngineUnavailable(
            "lycoris_b200: the adapter forward/backward hot path exists only as sm_100a CUDA kernels; "
            f"got a {x.device.type} input for {module.lora_name!r}. There is no CPU fallback — "
            "use the reference implementation for CPU runs."
        )
    if x.device.index != torch.cuda.current_device():
        # the C-ABI launches on the CURRENT device and stream: enter the tensor's device first
        with torch.cuda.device(x.device):
            return adapter_forward(module, x, args, kwargs, native_spec, assemble_fallback)
    org = module.org_module[0]
    W = org.weight.detach()
    bias = None if org.bias is None else org.bias.detach()
    ac = _autocast_dtype()
    cdt = ac if ac is not None else W.dtype
    if cdt == torch.float32:
        return _adapter_forward_f32(module, x, args, kwargs, W, bias, assemble_fallback)
    if cdt not in _HALF:
        raise NotImplementedError(f"lycoris_b200: compute dtype {cdt} is not supported (bf16 / fp16 / fp32)")
    if W.dtype != cdt:
        W = W.to(cdt)
    if bias is not None and bias.dtype != cdt:
        bias = bias.to(cdt)
    if not W.is_contiguous():
        W = W.contiguous()

    is_conv = module.module_type.startswith("conv")
    conv = _conv_params(module) if is_conv else None
    if is_conv and isinstance(conv["padding"], str):
        raise NotImplementedError("lycoris_b200: string padding modes are not supported")
    if x.dtype != cdt:
        if ac is None:
            raise RuntimeError(f"lycoris_b200: input dtype {x.dtype} != weight dtype {cdt} (no autocast active)")
        pointwise_nhwc = (is_conv and x.dim() == 4 and _is_pointwise(conv, W)
                          and x.is_contiguous(memory_format=torch.channels_last))
        if pointwise_nhwc or not (is_conv and x.dim() == 4 and _conv_engine_ok(x, W, conv, cdt)):
            x = x.to(cdt)
        # else: the convolution node casts and transposes to NHWC in one engine pass (lyco_transpose_cast)
    if is_conv and _is_pointwise(conv, W) and x.is_contiguous(memory_format=torch.channels_last) and x.dim() == 4:
        # NHWC 1x1 convolution is a plain linear over channels: run it on the tcgen05 GEMM
        y = adapter_forward_pointwise(module, x, W, bias, ac, native_spec, assemble_fallback, args, kwargs)
        return y

    plain = module._is_outermost_on_plain_forward() and not args and not kwargs and not _base_trainable(org)
    base = None
    if not plain:
        # another wrapper sits below us (stacking) or the base forward takes extra arguments:
        # keep its output and add only our delta contraction, like the reference does.
        base = module.org_forward(x, *args, **kwargs)

    spec = native_spec()
    if spec is not None and plain and _locon_side_ok(spec, conv, False, W, x.reshape(-1, x.shape[-1]), cdt):
        up, down = spec.factors
        y = _LoconSidePath.apply(x, W, bias, spec, up, down)
        return y
    if spec is not None:
        y = _AdapterContraction.apply(x, W, None if not plain else bias, spec, conv, not plain, ac,
                                      spec.dora[0] if spec.dora is not None else None, *spec.factors)
    else:
        Wm = assemble_fallback(W)
        if not plain:
            Wm = Wm - W
        y = _MergedContraction.apply(x, Wm, None if not plain else bias, conv)
    return y if plain else base + y


def _adapter_forward_f32(module, x, args, kwargs, W, bias, assemble_fallback):
    """fp32 base + fp32 adapter without autocast (the reference's CPU-style regime, BASELINE cfg #1):
    W' is assembled in fp32 by PyTorch ops (no rounding points to reproduce), Linear layers contract as three
    bf16 products on the engine; convolutions and unaligned shapes use the library."""
    if x.dtype != torch.float32:
        raise RuntimeError(f"lycoris_b200: input dtype {x.dtype} != weight dtype torch.float32 (no autocast active)")
    plain = (module._is_outermost_on_plain_forward() and not args and not kwargs
             and not _base_trainable(module.org_module[0]))
    base = None if plain else module.org_forward(x, *args, **kwargs)
    Wm = assemble_fallback(W)
    if not plain:
        Wm = Wm - W
    b = bias if plain else None
    if module.module_type == "linear" and _f32_supported(W) and x.shape[-1] % 8 == 0 and x.numel() > 0:
        y = _MergedContractionF32.apply(x, Wm, b)
    else:
        warning_once("lycoris_b200: fp32 convolutions / unaligned fp32 layers contract through the library")
        y = module.op(x, Wm, b, **module.kw_dict)
    return y if plain else base + y


def adapter_forward_pointwise(module, x, W, bias, ac, native_spec, assemble_fallback, args, kwargs):
    """channels_last 1x1 conv: [B,C,H,W] (NHWC storage) -> [B*H*W, C] linear -> NHWC output."""
    B, C, H, Wd = x.shape
    plain = (module._is_outermost_on_plain_forward() and not args and not kwargs
             and not _base_trainable(module.org_module[0]))
    base = None if plain else module.org_forward(x, *args, **kwargs)
    x2 = x.permute(0, 2, 3, 1).reshape(B * H * Wd, C)
    W2 = W.reshape(W.shape[0], C)
    spec = native_spec()
    if spec is not None:
        y2 = _AdapterContraction.apply(x2, W2, bias if plain else None, spec, None, not plain, ac,
                                       spec.dora[0] if spec.dora is not None else None, *spec.factors)
    else:
        Wm = assemble_fallback(W).reshape(W.shape[0], C)
        if not plain:
            Wm = Wm - W2
        y2 = _MergedContraction.apply(x2, Wm, bias if plain else None, None)
    y = y2.view(B, H, Wd, W.shape[0]).permute(0, 3, 1, 2)
    return y if plain else base + y
