"""ORACLE — test infrastructure, not product code.

A functional restatement of the reference's rebuild-mode hot path (KohakuBlueleaf/LyCORIS @ f49d30d)
in plain PyTorch ops, each step citing the reference file:line it follows.  The reference's
arithmetic lives in PyTorch ATen (un-pinned ``torch`` in its requirements.txt; this image ships
torch 2.11.0+cu128), so the restatement issues the same ATen calls in the same order with the same
rounding points; backward is autograd, exactly as in the reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``lycoris_b200``) never does.

Pinned: ``oracle/gen_golden.py`` imports the real reference from ``/root/reference`` in the build
container, runs identical seeded cases through reference modules and through this oracle, asserts
equality, and writes the reference outputs to ``tests/golden/*.pt`` — the fixtures every parity
test (CPU: oracle vs golden; GPU: CUDA engine vs golden and vs oracle) is anchored on.
"""

from __future__ import annotations

import math
import random

import torch
import torch.nn.functional as F

ALGOS = ("locon", "loha", "lokr", "ia3", "dylora")


# --------------------------------------------------------------------------- integer work
def factorization(dimension: int, factor: int = -1):
    """lycoris/functional/general.py:14-56 — (m, n), m*n == dimension, m <= n."""
    if factor > 0 and (dimension % factor) == 0:
        m, n = factor, dimension // factor
        return (n, m) if m > n else (m, n)
    if factor < 0:
        factor = dimension
    m, n = 1, dimension
    length = m + n
    while m < n:
        new_m = m + 1
        while dimension % new_m != 0:
            new_m += 1
        new_n = dimension // new_m
        if new_m + new_n > length or new_m > factor:
            break
        m, n = new_m, new_n
    return (n, m) if m > n else (m, n)


def lokr_shapes(out_dim, in_dim, k_size, lora_dim, factor=-1, decompose_both=False, full_matrix=False,
                unbalanced=False):
    """lycoris/modules/lokr.py:89-173 — which LoKr parameters exist and their shapes."""
    in_m, in_n = factorization(in_dim, factor)
    out_l, out_k = factorization(out_dim, factor)
    if unbalanced:
        out_l, out_k = out_k, out_l
    shapes = {}
    if decompose_both and lora_dim < max(out_l, in_m) / 2 and not full_matrix:
        shapes["lokr_w1_a"] = (out_l, lora_dim)
        shapes["lokr_w1_b"] = (lora_dim, in_m)
    else:
        shapes["lokr_w1"] = (out_l, in_m)
    if lora_dim >= max(out_k, in_n) / 2 or full_matrix:
        shapes["lokr_w2"] = (out_k, in_n, *k_size)
    else:
        shapes["lokr_w2_a"] = (out_k, lora_dim)
        shapes["lokr_w2_b"] = (lora_dim, in_n * int(math.prod(k_size)))
    return shapes


# ------------------------------------------------------------------------------- delta W
def _op(x, w, b, conv):
    if conv is None:
        return F.linear(x, w, b)
    return F.conv2d(x, w, b, **conv)


def _scalar(p, like):
    """trainable `scalar` parameter when use_scalar, else the constant-1 buffer (locon.py:150-153)."""
    return p["scalar"] if "scalar" in p else torch.tensor(1.0, device=like.device)


def delta_locon(p, shape, cfg):
    """lycoris/modules/locon.py:198-219 make_weight (no dropout): up @ down — or the Tucker rebuild
    einsum("i j ..., i p, j r -> p r ...") of functional/general.py:9-11 — times scalar."""
    wa, wb = p["lora_up.weight"], p["lora_down.weight"]
    if "lora_mid.weight" in p:
        t = p["lora_mid.weight"]
        weight = torch.einsum("i j ..., i p, j r -> p r ...", t, wa.view(wa.size(0), -1).transpose(0, 1),
                              wb.view(wb.size(0), -1))
    else:
        weight = wa.view(wa.size(0), -1) @ wb.view(wb.size(0), -1)
    return weight.view(shape) * _scalar(p, weight)


def apply_weight_decompose(weight, dora_scale, wd_on_out, multiplier=1):
    """DoRA rescale, lycoris/modules/locon.py:239-260 (identical copies in loha.py / lokr.py)."""
    weight = weight.to(dora_scale.dtype)
    dims = weight.dim() - 1
    if wd_on_out:
        norm = weight.reshape(weight.shape[0], -1).norm(dim=1).reshape(weight.shape[0], *[1] * dims)
    else:
        norm = weight.transpose(0, 1).reshape(weight.shape[1], -1).norm(dim=1, keepdim=True)
        norm = norm.reshape(weight.shape[1], *[1] * dims).transpose(0, 1)
    norm = norm + torch.finfo(weight.dtype).eps
    scale = dora_scale.to(weight.device) / norm
    if multiplier != 1:
        scale = multiplier * (scale - 1) + 1
    return weight * scale


class _HadaWeight(torch.autograd.Function):
    """lycoris/functional/loha.py:10-30.  A custom Function matters for parity: its backward runs
    OUTSIDE autocast, so under mixed precision the four factor gradients are formed from fp32
    re-products of the factors, not from bf16 matmuls as plain autograd would."""

    @staticmethod
    def forward(ctx, w1d, w1u, w2d, w2u, scale):
        ctx.save_for_backward(w1d, w1u, w2d, w2u, scale)
        return ((w1u @ w1d) * (w2u @ w2d)) * scale  # loha.py:14

    @staticmethod
    def backward(ctx, grad_out):
        w1d, w1u, w2d, w2u, scale = ctx.saved_tensors
        grad_out = grad_out * scale  # loha.py:20
        temp = grad_out * (w2u @ w2d)  # loha.py:21
        grad_w1u = temp @ w1d.T
        grad_w1d = w1u.T @ temp
        temp = grad_out * (w1u @ w1d)  # loha.py:25
        grad_w2u = temp @ w2d.T
        grad_w2d = w2u.T @ temp
        return grad_w1d, grad_w1u, grad_w2d, grad_w2u, None


_TUCKER_FWD = "i j ..., j r, i p -> p r ..."  # core [r, r, *k], down [r, in], up [r, out] -> [out, in, *k]


class _HadaWeightTucker(torch.autograd.Function):
    """lycoris/functional/loha.py:33-75: Hadamard product of two Tucker rebuilds.  As with _HadaWeight the custom
    backward runs outside autocast, and its contraction ORDER (core x down first, then x up; gradients peeled off
    in the order up -> down -> core) fixes the rounding, so it is restated contraction by contraction."""

    @staticmethod
    def forward(ctx, t1, w1d, w1u, t2, w2d, w2u, scale):
        ctx.save_for_backward(t1, w1d, w1u, t2, w2d, w2u, scale)
        return torch.einsum(_TUCKER_FWD, t1, w1d, w1u) * torch.einsum(_TUCKER_FWD, t2, w2d, w2u) * scale  # loha.py:38-41

    @staticmethod
    def backward(ctx, grad_out):
        t1, w1d, w1u, t2, w2d, w2u, scale = ctx.saved_tensors
        grad_out = grad_out * scale  # loha.py:46
        # NB the reference reuses `temp` = (t2 x w2d) — the OTHER branch's half product — when it forms grad_w1u
        # (loha.py:48,54), and (t1 x w1d) for grad_w2u (loha.py:62,68).  That is what is restated here.
        half2 = torch.einsum("i j ..., j r -> i r ...", t2, w2d)
        gw1 = torch.einsum("i j ..., i r -> r j ...", half2, w2u) * grad_out
        g_w1u = torch.einsum("r j ..., i j ... -> r i", half2, gw1)
        g_half = torch.einsum("i j ..., i r -> r j ...", gw1, w1u.T)
        g_w1d = torch.einsum("i r ..., i j ... -> r j", t1, g_half)
        g_t1 = torch.einsum("i j ..., j r -> i r ...", g_half, w1d.T)
        half1 = torch.einsum("i j ..., j r -> i r ...", t1, w1d)
        gw2 = torch.einsum("i j ..., i r -> r j ...", half1, w1u) * grad_out
        g_w2u = torch.einsum("r j ..., i j ... -> r i", half1, gw2)
        g_half = torch.einsum("i j ..., i r -> r j ...", gw2, w2u.T)
        g_w2d = torch.einsum("i r ..., i j ... -> r j", t2, g_half)
        g_t2 = torch.einsum("i j ..., j r -> i r ...", g_half, w2d.T)
        return g_t1, g_w1d, g_w1u, g_t2, g_w2d, g_w2u, None


def delta_loha(p, shape, cfg):
    """lycoris/modules/loha.py:194-226 get_weight -> functional/loha.py:119-147 diff_weight ->
    HadaWeight (or HadaWeightTucker when the Tucker cores hada_t1 / hada_t2 exist); scale is a 0-dim tensor in the
    factor dtype (loha.py:195-197)."""
    gamma = torch.tensor(cfg["scale"], dtype=p["hada_w1_b"].dtype, device=p["hada_w1_b"].device)
    if "hada_t1" in p:
        w = _HadaWeightTucker.apply(p["hada_t1"], p["hada_w1_b"], p["hada_w1_a"], p["hada_t2"], p["hada_w2_b"],
                                    p["hada_w2_a"], gamma)
    else:
        w = _HadaWeight.apply(p["hada_w1_b"], p["hada_w1_a"], p["hada_w2_b"], p["hada_w2_a"], gamma)
    return w.reshape(shape)


def delta_lokr(p, shape, cfg):
    """lycoris/modules/lokr.py:358-381 get_weight -> functional/lokr.py:11-20 make_kron."""
    w1 = p["lokr_w1"] if "lokr_w1" in p else p["lokr_w1_a"] @ p["lokr_w1_b"]
    if "lokr_w2" in p:
        w2 = p["lokr_w2"]
    elif "lokr_t2" in p:
        # Tucker core for the large block (lokr.py:121-128, 362-366): rebuild_tucker(t2, w2_a, w2_b)
        w2 = torch.einsum("i j ..., i p, j r -> p r ...", p["lokr_t2"], p["lokr_w2_a"], p["lokr_w2_b"])
    else:
        w2 = p["lokr_w2_a"] @ p["lokr_w2_b"]
    for _ in range(w2.dim() - w1.dim()):
        w1 = w1.unsqueeze(-1)
    rebuild = torch.kron(w1, w2.contiguous())
    if cfg["scale"] != 1:
        rebuild = rebuild * cfg["scale"]
    return rebuild.view(shape)


def merged_ia3(p, W, cfg):
    """lycoris/modules/ia3.py:91-102 make_weight(diff=False): W * (w*mult + 1) on out or in channels."""
    weight = p["weight"] * cfg["multiplier"] + 1
    if cfg["train_on_input"]:
        return W * weight
    return (W.transpose(0, 1) * weight).transpose(0, 1)


def delta_dylora(p, shape, cfg, b):
    """lycoris/modules/dylora.py:97-117: blocks < b frozen (.data), block b live,
    up @ (down * (alpha/(b+1) * mult))."""
    down = torch.concat([t.data for t in p["down_list"][:b]] + list(p["down_list"][b : b + 1]))
    up = torch.concat([t.data for t in p["up_list"][:b]] + list(p["up_list"][b : b + 1]), dim=1)
    scale = cfg["alpha"] / (b + 1)
    return (up @ (down * (scale * cfg["multiplier"]))).view(shape)


# ------------------------------------------------------------------------------- forward
def layer_forward(algo, x, W, bias, p, cfg, conv=None):
    """Rebuild-mode forward of one wrapped layer — the sequence at
    lycoris/modules/locon.py:317-332 (loha.py:309-322, lokr.py:551-566, ia3.py:136-144,
    dylora.py:150-157): base op, dW, cast/scale, W + dW*mult, minus W, delta op, add.

    ``cfg``: scale, multiplier, and per-algo extras (train_on_input, alpha, block draw ``b``).
    Run it under ``torch.autocast`` to reproduce the mixed-precision regime.
    """
    base = _op(x, W, bias, conv)
    base_weight = W.detach()
    mult = cfg.get("multiplier", 1.0)
    shape = tuple(W.shape)
    dora = p.get("dora_scale")
    wd_out = cfg.get("wd_on_out", True)
    if algo == "locon":
        diff = delta_locon(p, shape, cfg).to(base_weight.dtype) * cfg["scale"]  # locon.py:322
        if dora is not None:
            new_weight = apply_weight_decompose(base_weight + diff, dora, wd_out, mult)  # locon.py:323-326
        else:
            new_weight = base_weight + diff * mult  # locon.py:328
    elif algo == "loha":
        diff = delta_loha(p, shape, cfg).to(base_weight.dtype) * _scalar(p, base_weight)  # loha.py:311
        if dora is not None:
            new_weight = apply_weight_decompose(base_weight + diff, dora, wd_out, mult)  # loha.py:313-316
        else:
            new_weight = base_weight + diff * mult  # loha.py:318
    elif algo == "lokr":
        diff = delta_lokr(p, shape, cfg).to(base_weight.dtype) * _scalar(p, base_weight)  # lokr.py:553
        if dora is not None:
            new_weight = apply_weight_decompose(base_weight + diff, dora, wd_out, mult)  # lokr.py:555-558
        else:
            new_weight = base_weight + diff if mult == 1 else base_weight + diff * mult  # lokr.py:559-562
    elif algo == "ia3":
        new_weight = merged_ia3(p, W, cfg).to(base_weight.device, dtype=base_weight.dtype)  # ia3.py:137-141
    elif algo == "dylora":
        merged = delta_dylora(p, shape, cfg, cfg["b"]) + W  # dylora.py:128-129 get_merged_weight
        new_weight = merged.to(base_weight.dtype)  # dylora.py:296-298
    else:
        raise KeyError(algo)
    delta_weight = new_weight - base_weight
    delta = _op(x, delta_weight, None, conv)
    return base + delta


def draw_dylora_block(block_count):
    """dylora.py:108-110 — one draw from Python's global ``random`` per forward."""
    return random.randint(0, block_count - 1)


def layer_forward_backward(algo, x, W, bias, p, cfg, dy, conv=None, autocast_dtype=None):
    """Forward + autograd backward; returns (y, dx, {param: grad})."""
    x = x.detach().clone().requires_grad_(True)
    leaves = {}
    for k, v in p.items():
        if isinstance(v, (list, tuple)):
            leaves[k] = [t.detach().clone().requires_grad_(True) for t in v]
        else:
            leaves[k] = v.detach().clone().requires_grad_(True)
    dev = x.device.type
    if autocast_dtype is not None:
        with torch.autocast(dev, dtype=autocast_dtype):
            y = layer_forward(algo, x, W, bias, leaves, cfg, conv)
    else:
        y = layer_forward(algo, x, W, bias, leaves, cfg, conv)
    y.backward(dy.to(y.dtype))
    grads = {}
    for k, v in leaves.items():
        if isinstance(v, list):
            grads[k] = [t.grad for t in v]
        else:
            grads[k] = v.grad
    return y.detach(), x.grad, grads
