"""Shape / index helpers shared by the adapter algorithms.

``factorization`` is integer work and must agree bit-for-bit with the reference
(lycoris/functional/general.py:14-56); tests/test_factorization.py pins it against a table
generated from the reference itself.
"""

import torch
import torch.nn.functional as F

# index = weight.dim(): which functional op applies a weight of that rank
FUNC_LIST = [None, None, F.linear, F.conv1d, F.conv2d, F.conv3d]


def factorization(dimension: int, factor: int = -1) -> tuple[int, int]:
    """Split ``dimension = m * n`` with ``m <= n``.

    * ``factor > 0`` dividing ``dimension``: ``(factor, dimension // factor)`` sorted;
    * otherwise walk the divisors of ``dimension`` upwards from 1 and keep the largest one that is
      still ``<= factor`` (``factor < 0``: no cap), never stepping further than one divisor past
      ``sqrt(dimension)``.

    In LoKr ``m`` sizes the small "scale" block ``w1`` and ``n`` the large block ``w2``
    (``kron(w1, w2)`` is not commutative).
    """
    if factor > 0 and dimension % factor == 0:
        other = dimension // factor
        return (factor, other) if factor <= other else (other, factor)
    cap = dimension if factor < 0 else factor
    m = 1
    while m * m < dimension:
        nxt = m + 1
        while dimension % nxt:
            nxt += 1
        if nxt > cap:
            break
        m = nxt
    n = dimension // m
    return (m, n) if m <= n else (n, m)


def power2factorization(dimension: int, factor: int = -1):
    """``dimension = m * n`` with ``n`` a power of two and ``m`` even, ``m <= factor``; the last
    admissible ``m`` wins (reference: functional/general.py:59-81; used by BOFT only)."""
    if factor == -1:
        factor = dimension
    n = 0
    m = 0
    while m <= factor:
        m += 2
        while dimension % m != 0 and m < dimension:
            m += 2
        if m > factor:
            break
        q = dimension // m
        if q & (q - 1) == 0 and q > 0:
            n = q
    if n == 0:
        return None, n
    return dimension // n, n


def rebuild_tucker(t, wa, wb):
    """Tucker core ``t[i, j, ...]`` expanded by ``wa[i, p]`` and ``wb[j, r]`` -> ``[p, r, ...]``."""
    return torch.einsum("i j ..., i p, j r -> p r ...", t, wa, wb)


def tucker_weight_from_conv(up, down, mid):
    up = up.reshape(up.size(0), up.size(1))
    down = down.reshape(down.size(0), down.size(1))
    return torch.einsum("m n ..., i m, n j -> i j ...", mid, up, down)


def tucker_weight(wa, wb, t):
    temp = torch.einsum("i j ..., j r -> i r ...", t, wb)
    return torch.einsum("i j ..., i r -> r j ...", temp, wa)


def apply_dora_scale(org_weight, rebuild, dora_scale, scale):
    """Column-norm DoRA on ``org_weight + rebuild`` blended by ``scale`` (general.py:95-108)."""
    merged = (org_weight + rebuild).to(dora_scale.dtype)
    lead = merged.shape[1]
    norm = merged.transpose(0, 1).reshape(lead, -1).norm(dim=1, keepdim=True)
    norm = norm.reshape(lead, *[1] * (org_weight.dim() - 1)).transpose(0, 1)
    return org_weight + (merged / norm * dora_scale - org_weight) * scale
