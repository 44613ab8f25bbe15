"""Functional LoHa (cold path: merge / tooling; the training forward uses the CUDA merge kernel).

API per the reference's docs/API.md:47-79: ``weight_gen``, ``diff_weight``, ``bypass_forward_diff``
(lycoris/functional/loha.py).  Weights are passed as ``(w1d, w1u, w2d, w2u, t1, t2)``.
"""

import torch
import torch.nn as nn

from .general import FUNC_LIST


class _HadamardOfProducts(torch.autograd.Function):
    """(Wu1·Wd1) ⊙ (Wu2·Wd2) · gamma, keeping only the factors for backward (the two N×K products
    are rebuilt there) — same memory contract as the reference's HadaWeight (loha.py:10-30)."""

    @staticmethod
    def forward(ctx, d1, u1, d2, u2, gamma):
        ctx.save_for_backward(d1, u1, d2, u2, gamma)
        return (u1 @ d1) * (u2 @ d2) * gamma

    @staticmethod
    def backward(ctx, g):
        d1, u1, d2, u2, gamma = ctx.saved_tensors
        g = g * gamma
        side1 = g * (u2 @ d2)
        g_u1, g_d1 = side1 @ d1.T, u1.T @ side1
        side2 = g * (u1 @ d1)
        g_u2, g_d2 = side2 @ d2.T, u2.T @ side2
        return g_d1, g_u1, g_d2, g_u2, None


def _tucker_full(t, d, u):
    return torch.einsum("i j ..., j r, i p -> p r ...", t, d, u)


class _HadamardOfTuckers(torch.autograd.Function):
    """Tucker flavour: each side is core t[i,j,...] expanded by u[i,p] and d[j,r] (loha.py:33-75)."""

    @staticmethod
    def forward(ctx, t1, d1, u1, t2, d2, u2, gamma):
        ctx.save_for_backward(t1, d1, u1, t2, d2, u2, gamma)
        return _tucker_full(t1, d1, u1) * _tucker_full(t2, d2, u2) * gamma

    @staticmethod
    def _side_grads(g_full, t, d, u, half_for_up):
        """Gradients of one Tucker triple from ``g_full`` (gradient w.r.t. its rebuilt [p, r, ...] tensor).
        ``half_for_up`` is the (core x down) product the UP gradient is contracted with — see backward()."""
        g_u = torch.einsum("i r ..., p r ... -> i p", half_for_up, g_full)
        g_half = torch.einsum("p r ..., i p -> i r ...", g_full, u)
        g_d = torch.einsum("i j ..., i r ... -> j r", t, g_half)
        g_t = torch.einsum("i r ..., j r -> i j ...", g_half, d)
        return g_t, g_d, g_u

    @staticmethod
    def backward(ctx, g):
        t1, d1, u1, t2, d2, u2, gamma = ctx.saved_tensors
        g = g * gamma
        half1 = torch.einsum("i j ..., j r -> i r ...", t1, d1)
        half2 = torch.einsum("i j ..., j r -> i r ...", t2, d2)
        g1 = g * torch.einsum("i r ..., i p -> p r ...", half2, u2)
        g2 = g * torch.einsum("i r ..., i p -> p r ...", half1, u1)
        # REFERENCE QUIRK, reproduced (never silently fixed — DESIGN §4): upstream's HadaWeightTucker.backward
        # (lycoris/functional/loha.py:48-54, 62-68) contracts grad_w1u with the OTHER branch's half product
        # (t2 x w2d) and grad_w2u with (t1 x w1d); the mathematically exact gradient would use the branch's own
        # half.  Training dynamics of a LoHa-Tucker adapter depend on it, so parity with the reference's outputs
        # (tests/golden/tucker_*.pt) requires the same contraction.
        gt1, gd1, gu1 = _HadamardOfTuckers._side_grads(g1, t1, d1, u1, half_for_up=half2)
        gt2, gd2, gu2 = _HadamardOfTuckers._side_grads(g2, t2, d2, u2, half_for_up=half1)
        return gt1, gd1, gu1, gt2, gd2, gu2, None


def make_weight(w1d, w1u, w2d, w2u, scale):
    return _HadamardOfProducts.apply(w1d, w1u, w2d, w2u, scale)


def make_weight_tucker(t1, w1d, w1u, t2, w2d, w2u, scale):
    return _HadamardOfTuckers.apply(t1, w1d, w1u, t2, w2d, w2u, scale)


def weight_gen(org_weight, rank, tucker=True):
    """Fresh LoHa factors for ``org_weight``: returns ``w1d, w1u, w2d, w2u, t1, t2``."""
    out_dim, in_dim, *k = org_weight.shape
    if k and tucker:
        w1d, w1u = torch.empty(rank, in_dim), torch.empty(rank, out_dim)
        w2d, w2u = torch.empty(rank, in_dim), torch.empty(rank, out_dim)
        t1, t2 = torch.empty(rank, rank, *k), torch.empty(rank, rank, *k)
        nn.init.normal_(t1, std=0.1)
        nn.init.normal_(t2, std=0.1)
    else:
        w1d, w1u = torch.empty(rank, in_dim), torch.empty(out_dim, rank)
        w2d, w2u = torch.empty(rank, in_dim), torch.empty(out_dim, rank)
        t1 = t2 = None
    nn.init.normal_(w1d, std=1)
    nn.init.constant_(w1u, 0)
    nn.init.normal_(w2d, std=1)
    nn.init.normal_(w2u, std=0.1)
    return w1d, w1u, w2d, w2u, t1, t2


def diff_weight(*weights, gamma=1.0):
    """ΔW of a LoHa factor set, shaped ``[out, in, *k]``."""
    w1d, w1u, w2d, w2u, t1, t2 = weights
    gamma = gamma if isinstance(gamma, torch.Tensor) else torch.tensor(gamma, dtype=w1d.dtype, device=w1d.device)
    if t1 is not None and t2 is not None:
        _, in_dim = w1d.shape
        _, out_dim = w1u.shape
        k = list(t1.shape[2:])
        result = make_weight_tucker(t1, w1d, w1u, t2, w2d, w2u, gamma)
    else:
        _, in_dim, *k = w1d.shape
        out_dim = w1u.shape[0]
        flat = lambda d: d.reshape(d.size(0), -1)  # noqa: E731
        cols = lambda u: u.reshape(-1, u.size(1))  # noqa: E731
        result = make_weight(flat(w1d), cols(w1u), flat(w2d), cols(w2u), gamma)
    return result.reshape(out_dim, in_dim, *k)


def bypass_forward_diff(x, org_out, *weights, gamma=1.0, extra_args={}):
    """LoHa has no activation-side shortcut: rebuild ΔW, then one more op."""
    w1d, w1u, w2d, w2u, t1, t2 = weights
    diff_w = diff_weight(w1d, w1u, w2d, w2u, t1, t2, gamma=gamma)
    return FUNC_LIST[w1d.dim() if t1 is None else t1.dim()](x, diff_w, **extra_args)
