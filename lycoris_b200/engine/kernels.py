"""Tensor-level wrappers over the C-ABI: device pointers, shapes and the current CUDA stream go in,
nothing else.  PyTorch is used for memory and streams only; all arithmetic happens in the
hand-written sm_100a kernels."""

from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import ALGO_DYLORA, ALGO_IA3, ALGO_LOCON, ALGO_LOHA, ALGO_LOKR, ALGO_RAW, BF16, F16, F32, DeltaDesc

_DT = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"lycoris_b200: unsupported dtype {dt}") from None


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "lycoris_b200: the adapter hot path only exists as sm_100a CUDA kernels; "
                f"got a {t.device.type} tensor (no CPU fallback)."
            )


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def gemm_supported(*mats) -> bool:
    """TMA needs 16-byte aligned bases and row pitches that are multiples of 8 elements."""
    for t in mats:
        if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16:
            return False
        if t.dtype not in (torch.bfloat16, torch.float16):
            return False
    return True


_gemm_profile = None  # bench.py: list collecting (start_event, end_event, flops, M, N, K, bytes) per launch


_gemm_profile_refill = (0, 0)  # (every n profiled launches, spin this many GPU cycles first)
_gemm_profile_calls = 0


def set_gemm_profiler(sink, refill_every=0, refill_cycles=0):
    """Record a CUDA-event pair around every GEMM launch into ``sink`` (None disables).

    ``refill_every`` / ``refill_cycles``: before every n-th bracketed launch the stream first spins for that many GPU
    cycles (outside the bracket).  The host gets ahead of the GPU again during the spin, so the event pair brackets the
    kernel on a busy stream and not the GPU waiting for Python to reach the launch (an eager step that is host-bound
    on a slow box would otherwise count host time as kernel time)."""
    global _gemm_profile, _gemm_profile_refill, _gemm_profile_calls
    _gemm_profile = sink
    _gemm_profile_refill = (int(refill_every), int(refill_cycles))
    _gemm_profile_calls = 0


def _bracket_open():
    global _gemm_profile_calls
    every, cycles = _gemm_profile_refill
    if every > 0:
        if _gemm_profile_calls % every == 0:
            torch.cuda._sleep(cycles)
        _gemm_profile_calls += 1
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0, e1


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, out_dtype=None, split_k=0, out=None, accumulate=False):
    """``C[M,N] = A · Bᵀ (+ bias)``.

    ``a`` is stored ``[M, K]`` (``a_mn=False``) or ``[K, M]`` (``a_mn=True``); likewise ``b`` with
    ``N``.  Both must be row-contiguous 2-D tensors of the same 16-bit dtype.
    """
    _require_cuda(a, b, bias)
    lib = _lib.load()
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if K != Kb:
        raise ValueError(f"lycoris_b200.gemm: reduction mismatch {K} vs {Kb}")
    if a.dtype != b.dtype:
        raise TypeError("lycoris_b200.gemm: operand dtypes differ")
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    prof = _gemm_profile
    if prof is not None:
        e0, e1 = _bracket_open()
    rc = lib.lyco_gemm(
        _ptr(a), int(a_mn), a.stride(0),
        _ptr(b), int(b_mn), b.stride(0),
        _ptr(out), dtype_code(out.dtype), out.stride(0),
        _ptr(bias), dtype_code(bias.dtype) if bias is not None else 0,
        M, N, K, dtype_code(a.dtype), int(split_k), int(bool(accumulate)), _stream(),
    )
    _lib.check(rc, "gemm")
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K, M, N, K,
                     float(a.element_size() * (M * K + N * K) + out.element_size() * M * N)))  # + algorithmic bytes
    return out


def gemm_dual(a, b, a2, b2, *, b_mn=False, bias=None):
    """``C[M,N] = A·Bᵀ + A2·B2ᵀ (+ bias)`` in ONE kernel / one TMEM accumulator (lyco_gemm_dual): ``a`` [M,K], ``a2``
    [M,K2] row-major; ``b`` [N,K] and ``b2`` [N,K2] (``b_mn=False``) or ``b`` [K,N] and ``b2`` [K2,N] (``b_mn=True``)."""
    _require_cuda(a, b, a2, b2, bias)
    M, Kd = a.shape
    K2 = a2.shape[1]
    N = b.shape[1] if b_mn else b.shape[0]
    if (b.shape[0] if b_mn else b.shape[1]) != Kd or (b2.shape[0] if b_mn else b2.shape[1]) != K2 or a2.shape[0] != M \
            or (b2.shape[1] if b_mn else b2.shape[0]) != N:
        raise ValueError(f"lycoris_b200.gemm_dual: shape mismatch {tuple(a.shape)} {tuple(b.shape)} {tuple(a2.shape)} {tuple(b2.shape)}")
    if not (a.dtype == b.dtype == a2.dtype == b2.dtype):
        raise TypeError("lycoris_b200.gemm_dual: operand dtypes differ")
    out = torch.empty((M, N), device=a.device, dtype=a.dtype)
    prof = _gemm_profile
    if prof is not None:
        e0, e1 = _bracket_open()
    rc = _lib.load().lyco_gemm_dual(
        _ptr(a), a.stride(0), _ptr(b), int(b_mn), b.stride(0), Kd,
        _ptr(a2), a2.stride(0), _ptr(b2), b2.stride(0), K2,
        _ptr(out), out.stride(0), _ptr(bias), dtype_code(bias.dtype) if bias is not None else 0, M, N,
        dtype_code(a.dtype), _stream())
    _lib.check(rc, "gemm_dual")
    if prof is not None:
        e1.record()
        es = a.element_size()
        prof.append((e0, e1, 2.0 * M * N * (Kd + K2), M, N, Kd + K2, float(es * (M * (Kd + K2) + N * (Kd + K2) + M * N))))
    return out


def conv2d_supported(x, weight_shape, stride, padding, dilation, groups, dtype=None) -> bool:
    """Geometry the implicit-GEMM kernels cover: 2-D, dilation 1, groups 1, C % 64 == 0, O % 8 == 0.
    (Activations must be NHWC; callers convert NCHW tensors with one transpose pass.)"""
    if x.dim() != 4 or len(weight_shape) != 4 or (dtype or x.dtype) not in (torch.bfloat16, torch.float16):
        return False
    if x.numel() == 0:
        return False  # empty batch: nothing to launch, the library handles the bookkeeping
    O, C, R, S = weight_shape
    if groups != 1 or tuple(dilation) != (1, 1) or stride[0] != stride[1] or not (1 <= stride[0] <= 8):
        return False
    return C % 64 == 0 and O % 8 == 0


# debugging knobs: LYCO_LAYOUT_PASS=torch does the NCHW -> NHWC pass with Tensor.to; LYCO_CONV_OUT=nhwc keeps the
# convolution output channels_last even for NCHW inputs
_ENGINE_LAYOUT_PASS = os.environ.get("LYCO_LAYOUT_PASS", "engine") != "torch"
_NCHW_EPILOGUE = os.environ.get("LYCO_CONV_OUT", "auto") != "nhwc"


def as_nhwc(x, dtype=None):
    """``x`` [Nb, C, H, W] in channels_last storage and ``dtype`` (default: its own).  Already-NHWC tensors of
    that dtype pass through; a contiguous NCHW tensor (fp32 or the target dtype) goes through ONE engine pass
    that transposes and casts together (lyco_transpose_cast) — this is the copy autocast would make anyway."""
    dtype = dtype or x.dtype
    if x.dtype == dtype and x.is_contiguous(memory_format=torch.channels_last):
        return x
    if (_ENGINE_LAYOUT_PASS and x.is_cuda and x.is_contiguous() and dtype in (torch.bfloat16, torch.float16)
            and x.dtype in (torch.float32, dtype) and x.numel() > 0 and x.shape[0] < 65536):
        Nb, C, H, W = x.shape
        out = torch.empty((Nb, C, H, W), device=x.device, dtype=dtype, memory_format=torch.channels_last)
        rc = _lib.load().lyco_transpose_cast(_ptr(x), _ptr(out), Nb, C, H * W, dtype_code(x.dtype),
                                             dtype_code(dtype), _stream())
        _lib.check(rc, "transpose_cast")
        return out
    # any other striding (channel slices of a concatenation, expanded tensors, ...): one strided copy into a DENSE
    # NHWC buffer.  Not Tensor.to(memory_format=...): that returns `self` for a non-dense tensor whose strides
    # merely look channels-last.
    out = torch.empty(x.shape, device=x.device, dtype=dtype, memory_format=torch.channels_last)
    return out.copy_(x)


FILTER_FPROP, FILTER_DGRAD, FILTER_WBACK = 0, 1, 2


def filter_relayout(w, mode):
    """Filter re-layouts on the engine (lyco_filter_relayout).  ``w``: [O, C, R, S] contiguous 16-bit for
    FILTER_FPROP -> [O, R*S*C] and FILTER_DGRAD -> [C, R*S*O] (flipped taps); fp32 [O, R*S*C] with the 4-D shape
    passed as ``w = (tensor, (O, C, R, S))`` for FILTER_WBACK -> [O, C, R, S].  Filters larger than 3x3 taps go
    through PyTorch permutes."""
    if mode == FILTER_WBACK:
        t, (O, C, R, S) = w
    else:
        t = w
        O, C, R, S = w.shape
    taps = R * S
    _require_cuda(t)
    if taps > 9 or not t.is_contiguous() or O > 65535 or C % 8 or O % 8:
        if mode == FILTER_FPROP:
            return t.permute(0, 2, 3, 1).reshape(O, taps * C)
        if mode == FILTER_DGRAD:
            return t.flip(2, 3).permute(1, 2, 3, 0).reshape(C, taps * O)
        return t.view(O, R, S, C).permute(0, 3, 1, 2).contiguous()
    if mode == FILTER_FPROP:
        out = torch.empty((O, taps * C), device=t.device, dtype=t.dtype)
    elif mode == FILTER_DGRAD:
        out = torch.empty((C, taps * O), device=t.device, dtype=t.dtype)
    else:
        out = torch.empty((O, C, R, S), device=t.device, dtype=t.dtype)
    rc = _lib.load().lyco_filter_relayout(_ptr(t), _ptr(out), O, C, taps, mode, dtype_code(t.dtype), _stream())
    _lib.check(rc, "filter_relayout")
    return out


def conv2d_fprop(x, wk, bias, R, S, pad, stride, out_nchw=False, out_dtype=None):
    """``x``: [Nb, C, H, W] in channels_last storage; ``wk``: [O, R*S*C] (filter as [O,R,S,C]).
    Returns y [Nb, O, P, Q]: channels_last storage, or plain contiguous NCHW with ``out_nchw`` (written
    channel-major by the epilogue; needs P*Q % 32 == 0, otherwise the NHWC result is returned); with
    ``out_dtype=torch.float32`` the NCHW result is written in fp32 (other layouts come back 16-bit)."""
    _require_cuda(x, wk, bias)
    x = as_nhwc(x)
    Nb, C, H, W = x.shape
    O = wk.shape[0]
    P = (H + 2 * pad[0] - R) // stride + 1
    Q = (W + 2 * pad[1] - S) // stride + 1
    nchw = bool(out_nchw) and _NCHW_EPILOGUE and (P * Q) % 32 == 0
    layout = 0
    if nchw:
        # fp32 output only exists for the NCHW epilogue (plain coalesced stores); otherwise the caller casts
        f32 = out_dtype == torch.float32
        y = torch.empty((Nb, O, P, Q), device=x.device, dtype=torch.float32 if f32 else x.dtype)
        layout = 2 if f32 else 1
    else:
        y = torch.empty((Nb, O, P, Q), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    rc = _lib.load().lyco_conv2d_fprop(
        _ptr(x), _ptr(wk), _ptr(y), _ptr(bias), dtype_code(bias.dtype) if bias is not None else 0,
        Nb, H, W, C, O, R, S, pad[0], pad[1], stride, dtype_code(x.dtype), layout, _stream())
    _lib.check(rc, "conv2d_fprop")
    return y


def conv2d_wgrad(x, dy, R, S, pad, stride, split_k=0):
    """fp32 [O, R*S*C] weight gradient from channels_last ``x`` [Nb,C,H,W] and ``dy`` [Nb,O,P,Q]."""
    _require_cuda(x, dy)
    x, dy = as_nhwc(x), as_nhwc(dy)
    Nb, C, H, W = x.shape
    O = dy.shape[1]
    dw = torch.empty((O, R * S * C), device=x.device, dtype=torch.float32)
    rc = _lib.load().lyco_conv2d_wgrad(
        _ptr(x), _ptr(dy), _ptr(dw), Nb, H, W, C, O, R, S, pad[0], pad[1], stride, dtype_code(x.dtype),
        int(split_k), _stream())
    _lib.check(rc, "conv2d_wgrad")
    return dw


def make_desc(algo, out_dim, in_dim, *, factors, w_dtype, rank=0, up=0, uq=0, vp=0, vq=0, on_input=0,
              ia3_group=1, pre_round=0, pre_dtype=None, m_in=1.0, m_pre=1.0, m_post1=1.0, m_post2=1.0):
    f = list(factors) + [None] * (4 - len(factors))
    fd = None
    for t in factors:
        _require_cuda(t)
        if not t.is_contiguous():
            raise ValueError("lycoris_b200: factor tensors must be contiguous")
        fd = fd or t.dtype
        if t.dtype != fd:
            raise TypeError("lycoris_b200: all factors of a layer must share a dtype")
    d = DeltaDesc()
    d.algo = algo
    d.out_dim, d.in_dim, d.rank = int(out_dim), int(in_dim), int(rank)
    d.up, d.uq, d.vp, d.vq = int(up), int(uq), int(vp), int(vq)
    d.on_input, d.ia3_group = int(on_input), int(ia3_group)
    d.f_dtype = dtype_code(fd)
    d.w_dtype = dtype_code(w_dtype)
    d.pre_round = int(pre_round)
    d.pre_dtype = dtype_code(pre_dtype) if pre_dtype is not None else dtype_code(w_dtype)
    d.m_in, d.m_pre, d.m_post1, d.m_post2 = float(m_in), float(m_pre), float(m_post1), float(m_post2)
    d.f0, d.f1, d.f2, d.f3 = (None if t is None else t.data_ptr() for t in f)
    d._keepalive = tuple(factors)  # keep the storages alive as long as the descriptor
    return d


def merge_weight(desc: DeltaDesc, W: torch.Tensor) -> torch.Tensor:
    """``W' = W + ΔW`` with the reference's rounding points; one pass over ``W``."""
    _require_cuda(W)
    if not W.is_contiguous():
        raise ValueError("lycoris_b200.merge_weight: W must be contiguous")
    out = torch.empty_like(W)
    rc = _lib.load().lyco_merge_weight(ctypes.byref(desc), _ptr(W), _ptr(out), _stream())
    _lib.check(rc, "merge_weight")
    return out


def factor_grads(desc: DeltaDesc, dW: torch.Tensor, W, shapes):
    """fp32 gradients of the factor arrays from fp32 ``dW' = dYᵀ·X``."""
    _require_cuda(dW)
    assert dW.dtype == torch.float32 and dW.is_contiguous()
    # one buffer, arrays back to back at 64-float granularity: the library zero-fills them with ONE memset
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 63) // 64 * 64
    flat = torch.empty(total, device=dW.device, dtype=torch.float32)
    gs = [flat[o:o + n].view(s) for o, n, s in zip(offs, sizes, shapes)]
    g = gs + [None] * (4 - len(gs))
    rc = _lib.load().lyco_factor_grads(
        ctypes.byref(desc), _ptr(dW), _ptr(W), _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), _ptr(g[3]), _stream()
    )
    _lib.check(rc, "factor_grads")
    return gs


def grad_prep(dW: torch.Tensor, P, gscale: float, dtype: torch.dtype) -> torch.Tensor:
    """``G = dtype(gscale * dW * (P or 1))`` — 16-bit operand for the skinny gradient contractions."""
    _require_cuda(dW, P)
    assert dW.dtype == torch.float32 and dW.is_contiguous()
    G = torch.empty(dW.shape, device=dW.device, dtype=dtype)
    rc = _lib.load().lyco_grad_prep(_ptr(dW), _ptr(P), _ptr(G), dW.numel(), float(gscale), dtype_code(dtype), _stream())
    _lib.check(rc, "grad_prep")
    return G


def hada_merge(f, W, m_pre, m_post1, m_post2):
    """LoHa forward merge in one kernel: ``W' = rnd(W + chain((w1a·w1b) ⊙ (w2a·w2b)))`` — ``f`` = the four 16-bit factor
    arrays ``[N,r], [r,K'], [N,r], [r,K']`` (lyco_hada mode 0)."""
    _require_cuda(W, *f)
    N, r = f[0].shape
    Kp = f[1].shape[1]
    out = torch.empty_like(W)
    rc = _lib.load().lyco_hada(0, _ptr(f[0]), _ptr(f[1]), _ptr(f[2]), _ptr(f[3]), _ptr(W), _ptr(out), None, N, Kp, r,
                               dtype_code(f[0].dtype), dtype_code(W.dtype), float(m_pre), float(m_post1), float(m_post2),
                               1.0, _stream())
    _lib.check(rc, "hada_merge")
    return out


def hada_grad_operands(f, dW, gscale):
    """``(G1, G2) = (rnd(g·dW'·P2), rnd(g·dW'·P1))`` with P1, P2 recomputed on the tensor cores per tile (lyco_hada
    mode 1): the 16-bit operands of LoHa's four skinny gradient contractions."""
    _require_cuda(dW, *f)
    assert dW.dtype == torch.float32 and dW.is_contiguous()
    N, r = f[0].shape
    Kp = f[1].shape[1]
    G1 = torch.empty((N, Kp), device=dW.device, dtype=f[0].dtype)
    G2 = torch.empty((N, Kp), device=dW.device, dtype=f[0].dtype)
    rc = _lib.load().lyco_hada(1, _ptr(f[0]), _ptr(f[1]), _ptr(f[2]), _ptr(f[3]), _ptr(dW), _ptr(G1), _ptr(G2), N, Kp, r,
                               dtype_code(f[0].dtype), dtype_code(f[0].dtype), 1.0, 1.0, 1.0, float(gscale), _stream())
    _lib.check(rc, "hada_grad_operands")
    return G1, G2


def lokr_mix(x, w1, na, nb, nc, transpose, zero=None):
    """``out[m, a, c] = sum_b Wm(a, b) * x[m, b, c]`` with ``Wm = w1`` (or ``w1ᵀ``); ``x`` is a contiguous 16-bit
    ``[M, nb*nc]`` array, the result ``[M, na*nc]`` (lyco_lokr_mix).  ``zero``: optional contiguous fp32 tensor the
    kernel zero-fills on the side (gradient buffers of the kernels that follow)."""
    _require_cuda(x, w1, zero)
    M = x.shape[0]
    assert x.is_contiguous() and x.shape[1] == nb * nc and w1.is_contiguous()
    assert zero is None or (zero.dtype == torch.float32 and zero.is_contiguous())
    out = torch.empty((M, na * nc), device=x.device, dtype=x.dtype)
    rc = _lib.load().lyco_lokr_mix(_ptr(x), _ptr(out), _ptr(w1), dtype_code(w1.dtype), w1.stride(0), int(transpose),
                                   M, na, nb, nc, dtype_code(x.dtype), _ptr(zero), 0 if zero is None else zero.numel(),
                                   _stream())
    _lib.check(rc, "lokr_mix")
    return out


def lokr_w1grad(P, R, na, nb, nc, gscale, out=None):
    """fp32 ``g[a, b] = gscale * sum_{m,c} P[m, a, c] * R[m, b, c]`` (lyco_lokr_w1grad).  ``out``: an fp32 ``[na, nb]``
    buffer that is ALREADY zero (e.g. zero-filled by lokr_mix) — the sums are added to it without a memset."""
    _require_cuda(P, R, out)
    M = P.shape[0]
    assert P.is_contiguous() and R.is_contiguous() and P.shape[1] == na * nc and R.shape[1] == nb * nc and R.shape[0] == M
    g = out if out is not None else torch.empty((na, nb), device=P.device, dtype=torch.float32)
    assert g.dtype == torch.float32 and g.is_contiguous() and g.numel() == na * nb
    rc = _lib.load().lyco_lokr_w1grad(_ptr(P), _ptr(R), _ptr(g), M, na, nb, nc, float(gscale), dtype_code(P.dtype),
                                      int(out is None), _stream())
    _lib.check(rc, "lokr_w1grad")
    return g


def delta_weight(desc: DeltaDesc, shape, out_dtype=torch.float32, W=None, want_out=True, want_norm=False):
    """``(dW, norm_sq)``: the adapter's delta weight in ``out_dtype`` (None when ``want_out`` is False) and the fp32
    0-d tensor ``sum dW^2`` (None unless ``want_norm``) — lyco_delta_weight."""
    out = torch.empty(shape, device=torch.device("cuda", torch.cuda.current_device()), dtype=out_dtype) if want_out else None
    dev = out.device if out is not None else torch.device("cuda", torch.cuda.current_device())
    nsq = torch.zeros((), device=dev, dtype=torch.float32) if want_norm else None
    rc = _lib.load().lyco_delta_weight(ctypes.byref(desc), _ptr(W), _ptr(out), dtype_code(out_dtype), _ptr(nsq), _stream())
    _lib.check(rc, "delta_weight")
    return out, nsq


def dora_fwd(Wm, dora_scale, on_out, taps, mult, eps, scale_dtype=torch.float32):
    """``(W'', sumsq)`` — DoRA rescale of the merged 16-bit weight (lyco_dora_fwd)."""
    _require_cuda(Wm, dora_scale)
    N = Wm.shape[0]
    K = Wm.numel() // N
    groups = N if on_out else K // taps
    assert Wm.is_contiguous() and dora_scale.dtype == torch.float32 and dora_scale.numel() == groups
    out = torch.empty_like(Wm)
    sumsq = torch.empty(groups, device=Wm.device, dtype=torch.float32)
    rc = _lib.load().lyco_dora_fwd(_ptr(Wm), _ptr(out), _ptr(dora_scale), _ptr(sumsq), N, K, int(on_out), int(taps),
                                   float(mult), float(eps), dtype_code(Wm.dtype), dtype_code(scale_dtype), _stream())
    _lib.check(rc, "dora_fwd")
    return out, sumsq


def dora_bwd(dW, Wm, dora_scale, sumsq, on_out, taps, mult, eps, want_scale_grad=True, scale_dtype=torch.float32):
    """In place: fp32 ``dW`` (gradient of W'') becomes the gradient of Wm; returns the fp32 gradient of dora_scale."""
    _require_cuda(dW, Wm)
    N = Wm.shape[0]
    K = Wm.numel() // N
    groups = sumsq.numel()
    assert dW.dtype == torch.float32 and dW.is_contiguous() and dW.numel() == Wm.numel()
    t = torch.empty(groups, device=dW.device, dtype=torch.float32)
    g = torch.empty(groups, device=dW.device, dtype=torch.float32) if want_scale_grad else None
    rc = _lib.load().lyco_dora_bwd(_ptr(dW), _ptr(Wm), _ptr(dora_scale), _ptr(sumsq), _ptr(t), _ptr(g), N, K, int(on_out),
                                   int(taps), float(mult), float(eps), dtype_code(Wm.dtype), dtype_code(scale_dtype), _stream())
    _lib.check(rc, "dora_bwd")
    return g


__all__ = [
    "gemm", "gemm_dual", "gemm_supported", "conv2d_supported", "as_nhwc", "conv2d_fprop", "conv2d_wgrad", "make_desc", "merge_weight", "factor_grads", "dtype_code",
    "grad_prep", "hada_merge", "hada_grad_operands", "lokr_mix", "lokr_w1grad", "delta_weight", "dora_fwd", "dora_bwd", "ALGO_LOCON", "ALGO_LOHA", "ALGO_LOKR", "ALGO_IA3", "ALGO_DYLORA", "ALGO_RAW", "BF16", "F16", "F32",
]
