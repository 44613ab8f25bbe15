"""Functional LoCon (cold path; API per docs/API.md:47-79, reference lycoris/functional/locon.py).
Weights are passed as ``(down, up, mid)``."""

import math

import torch
import torch.nn as nn

from .general import FUNC_LIST, rebuild_tucker


def weight_gen(org_weight, rank, tucker=True):
    """Fresh LoCon factors: ``down, up, mid`` (``mid`` only for Tucker on a k>1 convolution)."""
    out_dim, in_dim, *k = org_weight.shape
    ones = [1] * len(k)
    if k and tucker:
        down, up, mid = torch.empty(rank, in_dim, *ones), torch.empty(out_dim, rank, *ones), torch.empty(rank, rank, *k)
        nn.init.kaiming_uniform_(mid, a=math.sqrt(5))
    else:
        down, up, mid = torch.empty(rank, in_dim), torch.empty(out_dim, rank), None
    nn.init.kaiming_uniform_(down, a=math.sqrt(5))
    nn.init.constant_(up, 0)
    return down, up, mid


def diff_weight(*weights, gamma=1.0):
    """ΔW = (gamma·up) · down, or the Tucker rebuild, shaped ``[out, in, *k]``."""
    d, u, m = weights
    _, in_dim, *k = d.shape
    out_dim = u.shape[0]
    u = u * gamma
    if m is None:
        result = u.reshape(-1, u.size(1)) @ d.reshape(d.size(0), -1)
    else:
        k = list(m.shape[2:])
        result = rebuild_tucker(m, u.reshape(u.size(0), -1).transpose(0, 1), d.reshape(d.size(0), -1))
    return result.reshape(out_dim, in_dim, *k)


def bypass_forward_diff(x, org_out, *weights, gamma=1.0, extra_args={}):
    """Activation-side LoCon: down (k×k) → [mid] → up (1×1), scaled by gamma."""
    d, u, m = weights
    op = FUNC_LIST[d.dim()]
    if m is None:
        h = op(x, d, **extra_args)
    else:
        h = op(op(x, d), m, **extra_args)
    return op(h, u) * gamma
