"""Functional LoKr (cold path; API per docs/API.md:47-79, reference lycoris/functional/lokr.py).
Weights are passed as ``(w1, w1a, w1b, w2, w2a, w2b, t)``.

Index convention (what the CUDA kernels implement):
``kron(w1, w2)[pu*vp + pv, u*vq + v] = w1[pu, u] * w2[pv, v]`` — the w1 index is the slow one on
both axes; for convolutions w2 carries the kernel taps and w1 is broadcast over them.
"""

import math

import torch
import torch.nn.functional as F

from .general import FUNC_LIST, factorization, rebuild_tucker


def make_kron(w1, w2, scale):
    """``kron(w1, w2) * scale`` with ``w1`` broadcast over any trailing kernel dims of ``w2``."""
    w1 = w1.reshape(*w1.shape, *[1] * (w2.dim() - w1.dim()))
    out = torch.kron(w1, w2.contiguous())
    return out if scale == 1 else out * scale


def _plan(out_dim, in_dim, k, rank, tucker, factor, decompose_both, full_matrix, unbalanced, linear_rule):
    """Which LoKr factors exist for a layer (mirrors the shape logic of modules/lokr.py)."""
    in_m, in_n = factorization(in_dim, factor)
    out_l, out_k = factorization(out_dim, factor)
    if unbalanced:
        out_l, out_k = out_k, out_l
    small_w1 = decompose_both and rank < max(out_l, in_m) / 2 and (linear_rule or not full_matrix)
    if linear_rule:
        full_w2 = not (rank < max(out_k, in_n) / 2)
    else:
        full_w2 = rank >= max(out_k, in_n) / 2 or full_matrix
    return (out_l, out_k), (in_m, in_n), small_w1, full_w2, bool(tucker and k and any(i != 1 for i in k))


def weight_gen(org_weight, rank, tucker=True, factor=-1, decompose_both=False, full_matrix=False,
               unbalanced_factorization=False):
    """Fresh LoKr factors: ``w1, w1a, w1b, w2, w2a, w2b, t2`` (unused entries are None)."""
    out_dim, in_dim, *k = org_weight.shape
    (out_l, out_k), (in_m, in_n), small_w1, full_w2, use_tucker = _plan(
        out_dim, in_dim, k, rank, tucker, factor, decompose_both, full_matrix, unbalanced_factorization,
        linear_rule=not k)
    w1 = w1a = w1b = w2 = w2a = w2b = t2 = None
    if small_w1:
        w1a, w1b = torch.empty(out_l, rank), torch.empty(rank, in_m)
        torch.nn.init.kaiming_uniform_(w1a, a=math.sqrt(5))
        torch.nn.init.kaiming_uniform_(w1b, a=math.sqrt(5))
    else:
        w1 = torch.empty(out_l, in_m)
        torch.nn.init.kaiming_uniform_(w1, a=math.sqrt(5))
    if full_w2:
        w2 = torch.empty(out_k, in_n, *k)
        torch.nn.init.constant_(w2, 0)
    elif k and use_tucker:
        t2, w2a, w2b = torch.empty(rank, rank, *k), torch.empty(rank, out_k), torch.empty(rank, in_n)
        torch.nn.init.kaiming_uniform_(t2, a=math.sqrt(5))
        torch.nn.init.kaiming_uniform_(w2a, a=math.sqrt(5))
        torch.nn.init.constant_(w2b, 0)
    else:
        w2a, w2b = torch.empty(out_k, rank), torch.empty(rank, in_n, *k)
        torch.nn.init.kaiming_uniform_(w2a, a=math.sqrt(5))
        torch.nn.init.constant_(w2b, 0)
    return w1, w1a, w1b, w2, w2a, w2b, t2


def _rank_of(w1a, w2a, gamma):
    if w1a is not None:
        return w1a.shape[1]
    if w2a is not None:
        return w2a.shape[1]
    return gamma


def diff_weight(*weights, gamma=1.0):
    """ΔW = kron(w1 or w1a·w1b, w2 or w2a·w2b or Tucker) · gamma / rank."""
    w1, w1a, w1b, w2, w2a, w2b, t = weights
    scale = gamma / _rank_of(w1a, w2a, gamma)
    if w1 is None:
        w1 = w1a @ w1b
    if w2 is None:
        if t is None:
            r, o, *k = w2b.shape
            w2 = (w2a @ w2b.view(r, -1)).view(-1, o, *k)
        else:
            w2 = rebuild_tucker(t, w2a, w2b)
    return make_kron(w1, w2, scale)


def bypass_forward_diff(h, org_out, *weights, gamma=1.0, extra_args={}):
    """Structured (w1 ⊗ w2)·x without building ΔW: split channels into (uq, vq) groups, contract
    vq with w2 (grouped op), then the uq axis with w1 (cross-group linear)."""
    w1, w1a, w1b, w2, w2a, w2b, t = weights
    dim = t.dim() if t is not None else (w2.dim() if w2 is not None else w2b.dim())
    rank = w1b.size(0) if w1 is None else (w2b.size(0) if w2 is None else gamma)
    scale = gamma / rank
    is_conv = dim > 2
    op = FUNC_LIST[dim]
    kw = extra_args if is_conv else {}
    c = w1 if w1 is not None else w1a @ w1b
    uq = c.size(1)

    if is_conv:
        B, _, *rest = h.shape
        grouped = h.reshape(B * uq, -1, *rest)
    else:
        grouped = h.reshape(*h.shape[:-1], uq, -1)

    ones = [1] * (dim - 2)
    if w2 is not None:
        hb = op(grouped, w2, **kw)
    elif t is not None:
        hb = op(op(op(grouped, w2b.view(*w2b.shape, *ones)), t, **kw), w2a.view(*w2a.shape, *ones))
    elif is_conv:
        hb = op(op(grouped, w2b, **kw), w2a.view(*w2a.shape, *ones))
    else:
        hb = op(op(grouped, w2b, **kw), w2a)

    if is_conv:
        hb = hb.view(B, -1, *hb.shape[1:])
        hc = F.linear(hb.transpose(1, -1), c).transpose(1, -1)
        out = hc.reshape(B, -1, *hc.shape[3:])
    else:
        hc = F.linear(hb.transpose(-1, -2), c).transpose(-1, -2)
        out = hc.reshape(*hc.shape[:-2], -1)
    return out * scale
