"""LoKr adapter:  dW = kron(w1, w2) · scale       (reference lycoris/modules/lokr.py,
lycoris/functional/lokr.py) — the north-star configuration (SDXL, factor 8).

``(a, b) = factorization(out, factor)``, ``(c, d) = factorization(in, factor)``;
``w1`` is ``[a, c]`` (or ``w1_a [a,r] · w1_b [r,c]`` with ``decompose_both``), ``w2`` is
``[b, d(,kh,kw)]`` when ``r >= max(b, d)/2`` or ``full_matrix`` (then scale := 1), otherwise
``w2_a [b,r] · w2_b [r, d·kh·kw]`` (or a Tucker core ``t2``).

Hot path: the Kronecker tile is expanded on the fly while streaming W once (``lyco_merge_weight``);
``g_w1`` / ``g_w2`` come from one pass over dW' (``lyco_factor_grads``) — ``torch.kron``'s
broadcast-multiply + permute copies over N×K and their autograd reductions disappear.
"""

import math
from functools import lru_cache

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..functional import factorization, rebuild_tucker
from ..functional.lokr import make_kron
from ..logging import logger
from .base import LycorisBaseModule


@lru_cache(maxsize=None)
def logging_force_full_matrix(lora_dim, dim, factor):
    logger.warning(f"lora_dim {lora_dim} is too large for dim={dim} and {factor=}, using full matrix mode.")


class LokrModule(LycorisBaseModule):
    name = "kron"
    support_module = {"linear", "conv1d", "conv2d", "conv3d"}
    weight_list = [
        "lokr_w1", "lokr_w1_a", "lokr_w1_b", "lokr_w2", "lokr_w2_a", "lokr_w2_b", "lokr_t1", "lokr_t2",
        "alpha", "dora_scale",
    ]
    weight_list_det = ["lokr_w1", "lokr_w1_a"]

    def __init__(
        self,
        lora_name,
        org_module: nn.Module,
        multiplier=1.0,
        lora_dim=4,
        alpha=1,
        dropout=0.0,
        rank_dropout=0.0,
        module_dropout=0.0,
        use_tucker=False,
        use_scalar=False,
        decompose_both=False,
        factor: int = -1,  # factorization factor
        rank_dropout_scale=False,
        weight_decompose=False,
        wd_on_out=True,
        full_matrix=False,
        bypass_mode=None,
        rs_lora=False,
        unbalanced_factorization=False,
        **kwargs,
    ):
        super().__init__(
            lora_name, org_module, multiplier, dropout, rank_dropout, module_dropout, rank_dropout_scale, bypass_mode
        )
        if self.module_type not in self.support_module:
            raise ValueError(f"{self.module_type} is not supported in LoKr algo.")

        factor = int(factor)
        self.lora_dim = lora_dim
        self.tucker = False
        self.use_w1 = False
        self.use_w2 = False
        self.full_matrix = full_matrix
        self.rs_lora = rs_lora

        is_conv = self.module_type.startswith("conv")
        if is_conv:
            in_dim, out_dim, k_size = org_module.in_channels, org_module.out_channels, tuple(org_module.kernel_size)
        else:
            in_dim, out_dim, k_size = org_module.in_features, org_module.out_features, ()
        self.shape = (out_dim, in_dim, *k_size)

        in_m, in_n = factorization(in_dim, factor)
        out_l, out_k = factorization(out_dim, factor)
        if unbalanced_factorization:
            out_l, out_k = out_k, out_l
        # kron(w1 [out_l, in_m], w2 [out_k, in_n, *k])
        self.tucker = is_conv and bool(use_tucker) and any(i != 1 for i in k_size)
        k_elems = 1
        for i in k_size:
            k_elems *= i

        # small block w1: full unless decompose_both asks for (and the rank allows) a product
        if decompose_both and lora_dim < max(out_l, in_m) / 2 and not self.full_matrix:
            self.lokr_w1_a = nn.Parameter(torch.empty(out_l, lora_dim))
            self.lokr_w1_b = nn.Parameter(torch.empty(lora_dim, in_m))
        else:
            self.use_w1 = True
            self.lokr_w1 = nn.Parameter(torch.empty(out_l, in_m))

        # large block w2: full when the rank would not compress it (or full_matrix)
        if lora_dim >= max(out_k, in_n) / 2 or self.full_matrix:
            if not self.full_matrix:
                logging_force_full_matrix(lora_dim, max(in_dim, out_dim), factor)
            self.use_w2 = True
            self.lokr_w2 = nn.Parameter(torch.empty(out_k, in_n, *k_size))
        elif self.tucker:
            self.lokr_t2 = nn.Parameter(torch.empty(lora_dim, lora_dim, *k_size))
            self.lokr_w2_a = nn.Parameter(torch.empty(lora_dim, out_k))  # 1-mode
            self.lokr_w2_b = nn.Parameter(torch.empty(lora_dim, in_n))  # 2-mode
        else:
            self.lokr_w2_a = nn.Parameter(torch.empty(out_k, lora_dim))
            self.lokr_w2_b = nn.Parameter(torch.empty(lora_dim, in_n * k_elems))

        self._init_dora(org_module, weight_decompose, wd_on_out)

        self.dropout = dropout
        if dropout:
            print("[WARN]LoHa/LoKr haven't implemented normal dropout yet.")
        self.rank_dropout = rank_dropout
        self.rank_dropout_scale = rank_dropout_scale
        self.module_dropout = module_dropout

        alpha, _ = self._init_alpha(alpha, lora_dim, rs_lora)
        if self.use_w2 and self.use_w1:
            alpha = lora_dim  # both blocks full: no low-rank scaling (scale = 1)
        r_factor = math.sqrt(lora_dim) if self.rs_lora else lora_dim
        self.scale = alpha / r_factor
        self.register_buffer("alpha", torch.tensor(alpha * (lora_dim / r_factor)))
        self._init_scalar(use_scalar)

        if self.use_w2:
            if use_scalar:
                torch.nn.init.kaiming_uniform_(self.lokr_w2, a=math.sqrt(5))
            else:
                torch.nn.init.constant_(self.lokr_w2, 0)
        else:
            if self.tucker:
                torch.nn.init.kaiming_uniform_(self.lokr_t2, a=math.sqrt(5))
            torch.nn.init.kaiming_uniform_(self.lokr_w2_a, a=math.sqrt(5))
            if use_scalar:
                torch.nn.init.kaiming_uniform_(self.lokr_w2_b, a=math.sqrt(5))
            else:
                torch.nn.init.constant_(self.lokr_w2_b, 0)
        if self.use_w1:
            torch.nn.init.kaiming_uniform_(self.lokr_w1, a=math.sqrt(5))
        else:
            torch.nn.init.kaiming_uniform_(self.lokr_w1_a, a=math.sqrt(5))
            torch.nn.init.kaiming_uniform_(self.lokr_w1_b, a=math.sqrt(5))

    @classmethod
    def make_module_from_state_dict(cls, lora_name, orig_module, w1, w1a, w1b, w2, w2a, w2b, _, t2, alpha, dora_scale):
        """Re-derive rank / factor / mode from the stored tensor shapes (lokr.py:246-342)."""
        full_matrix = False
        if w1a is not None:
            lora_dim = w1a.size(1)
        elif w2a is not None:
            lora_dim = w2a.size(1)
        else:
            full_matrix, lora_dim = True, 1

        w1_shape = tuple(w1.shape) if w1 is not None else (w1a.size(0), w1b.size(1))
        w2_shape = tuple(w2.shape[:2]) if w2 is not None else (w2a.size(0), w2b.size(1))
        out_dim, in_dim = w1_shape[0] * w2_shape[0], w1_shape[1] * w2_shape[1]

        if w1_shape[0] == factorization(out_dim, -1)[0] and w1_shape[1] == factorization(in_dim, -1)[0]:
            factor = -1
        else:
            rows, cols = (w1_shape[0], w2_shape[0]), (w1_shape[1], w2_shape[1])

            def fits(f):
                return out_dim % f == 0 and in_dim % f == 0 and f in rows and f in cols

            f1 = max(w1.shape) if w1 is not None else max(w1a.size(0), w1b.size(1))
            f2 = max(w2.shape) if w2 is not None else max(w2a.size(0), w2b.size(1))
            factor = f1 if fits(f1) else (f2 if fits(f2) else min(f1, f2))

        module = cls(
            lora_name, orig_module, 1, lora_dim, float(alpha),
            use_tucker=t2 is not None, decompose_both=w1 is None and w2 is None, factor=factor,
            weight_decompose=dora_scale is not None, full_matrix=full_matrix,
        )
        if w1 is not None:
            module.lokr_w1.copy_(w1)
        else:
            module.lokr_w1_a.copy_(w1a)
            module.lokr_w1_b.copy_(w1b)
        if w2 is not None:
            module.lokr_w2.copy_(w2)
        else:
            module.lokr_w2_a.copy_(w2a)
            module.lokr_w2_b.copy_(w2b)
        if t2 is not None:
            module.lokr_t2.copy_(t2)
        if dora_scale is not None:
            module.dora_scale.copy_(dora_scale)
        return module

    def load_weight_hook(self, module: nn.Module, incompatible_keys):
        self._reset_scalar_after_load(incompatible_keys)

    def custom_state_dict(self):
        destination = {"alpha": self.alpha}
        if self.wd:
            destination["dora_scale"] = self.dora_scale
        if self.use_w1:
            destination["lokr_w1"] = self.lokr_w1 * self.scalar
        else:
            destination["lokr_w1_a"] = self.lokr_w1_a * self.scalar
            destination["lokr_w1_b"] = self.lokr_w1_b
        if self.use_w2:
            destination["lokr_w2"] = self.lokr_w2
        else:
            destination["lokr_w2_a"] = self.lokr_w2_a
            destination["lokr_w2_b"] = self.lokr_w2_b
            if self.tucker:
                destination["lokr_t2"] = self.lokr_t2
        return destination

    # ------------------------------------------------------------------ dW (PyTorch ops)
    def _w1(self):
        return self.lokr_w1 if self.use_w1 else self.lokr_w1_a @ self.lokr_w1_b

    def _w2(self):
        if self.use_w2:
            return self.lokr_w2
        if self.tucker:
            return rebuild_tucker(self.lokr_t2, self.lokr_w2_a, self.lokr_w2_b)
        return self.lokr_w2_a @ self.lokr_w2_b

    def get_weight(self, shape):
        weight = make_kron(self._w1(), self._w2(), self.scale)
        if shape is not None:
            weight = weight.view(shape)
        if self.training and self.rank_dropout:
            weight = self._rank_drop_rows(weight)
        return weight

    def get_diff_weight(self, multiplier=1, shape=None, device=None):
        # NB: like the reference (lokr.py:384-385) `scale` is applied again on top of get_weight.
        eng = self._delta_via_engine(self.scale, self.scale * multiplier, shape=shape)
        diff = eng[0] if eng is not None else self.get_weight(shape) * (self.scale * multiplier)
        if device is not None:
            diff = diff.to(device)
        return diff, None

    def get_merged_weight(self, multiplier=1, shape=None, device=None):
        diff = self.get_diff_weight(multiplier=1, shape=shape, device=device)[0]
        weight = self.org_weight
        if self.wd:
            return self.apply_weight_decompose(weight + diff, multiplier), None
        return weight + diff * multiplier, None

    @torch.no_grad()
    def apply_max_norm(self, max_norm, device=None):
        eng = self._delta_via_engine(self.scale, 1.0, want_out=False, want_norm=True)
        if eng is not None:
            orig_norm = eng[1].sqrt().to(self._w1().dtype)  # ||kron(w1, w2) * scale||_F without forming the product
        else:
            orig_norm = self.get_weight(self.shape).norm()
        norm = torch.clamp(orig_norm, max_norm / 2)
        desired = torch.clamp(norm, max=max_norm)
        ratio = desired.cpu() / norm.cpu()
        scaled = norm != desired
        if scaled:
            # spread the correction evenly over the factors whose product forms dW
            parts = [self.lokr_w1] if self.use_w1 else [self.lokr_w1_a, self.lokr_w1_b]
            if self.use_w2:
                parts.append(self.lokr_w2)
            else:
                parts += ([self.lokr_t2] if self.tucker else []) + [self.lokr_w2_a, self.lokr_w2_b]
            modules = 4 - self.use_w1 - self.use_w2 + (not self.use_w2 and self.tucker)
            for p in parts:
                p *= ratio ** (1 / modules)
        return scaled, orig_norm * ratio

    # ------------------------------------------------------------------------- bypass
    def bypass_forward_diff(self, h, scale=1):
        """Structured (w1 ⊗ w2)·x: group channels as (uq, vq), contract vq with w2, then uq with w1."""
        is_conv = self.module_type.startswith("conv")
        ones = [1] * (len(self.shape) - 2)
        c = self._w1()
        uq = c.size(1)
        if is_conv:
            B, _, *rest = h.shape
            grouped = h.reshape(B * uq, -1, *rest)
        else:
            grouped = h.reshape(*h.shape[:-1], uq, -1)

        if self.use_w2:
            hb = self.op(grouped, self.lokr_w2, **self.kw_dict)
        elif self.tucker:
            a = self.lokr_w2_b.view(*self.lokr_w2_b.shape, *ones)
            b = self.lokr_w2_a.view(*self.lokr_w2_a.shape, *ones)
            hb = self.op(self.op(self.op(grouped, a), self.lokr_t2, **self.kw_dict), b)
        elif is_conv:
            a = self.lokr_w2_b.view(*self.lokr_w2_b.shape[:1], -1, *self.shape[2:])
            b = self.lokr_w2_a.view(*self.lokr_w2_a.shape, *ones)
            hb = self.op(self.op(grouped, a, **self.kw_dict), b)
        else:
            hb = self.op(self.op(grouped, self.lokr_w2_b, **self.kw_dict), self.lokr_w2_a)

        if is_conv:
            hb = hb.view(B, -1, *hb.shape[1:])
            hc = F.linear(hb.transpose(1, -1), c).transpose(1, -1)
            out = hc.reshape(B, -1, *hc.shape[3:])
        else:
            hc = F.linear(hb.transpose(-1, -2), c).transpose(-1, -2)
            out = hc.reshape(*hc.shape[:-2], -1)
        return self.drop(out * scale * self.scalar)

    def bypass_forward(self, x, scale=1):
        return self.org_forward(x) + self.bypass_forward_diff(x, scale=scale)

    # ------------------------------------------------------------------------ forward
    def _native_spec(self):
        from ..engine.kernels import ALGO_LOKR
        from ..engine.ops import NativeSpec

        if self.tucker or isinstance(self.scalar, nn.Parameter) or (self.training and self.rank_dropout):
            return None
        # inner low-rank products are tiny (e.g. 160x4 @ 4x160): formed by PyTorch, with autograd
        # carrying g_w2 back to w2_a / w2_b; the kernels see the two Kronecker blocks only.
        w1, w2 = self._w1(), self._w2()
        w2 = w2.reshape(w2.shape[0], -1)
        return NativeSpec(
            algo=ALGO_LOKR,
            factors=(w1, w2),
            up=w1.shape[0], uq=w1.shape[1], vp=w2.shape[0], vq=w2.shape[1],
            matmul_product=False,  # torch.kron is not an autocast op: the product keeps the factor dtype
            m_pre=float(self.scale),
            m_post1=self._scalar_host(),
            m_post2=1.0 if self.wd else float(self.multiplier),
            dora=(self.dora_scale, self.wd_on_out, float(self.multiplier)) if self.wd else None,
        )

    def _assemble(self, base_weight):
        diff = self.get_weight(self.shape).to(base_weight.dtype) * self.scalar
        if self.wd:
            return self.apply_weight_decompose(base_weight + diff, self.multiplier).to(base_weight.dtype)
        if self.multiplier == 1:
            return base_weight + diff
        return base_weight + diff * self.multiplier

    def forward(self, x: torch.Tensor, *args, **kwargs):
        if self._module_dropped():
            return self.org_forward(x, *args, **kwargs)
        if self.bypass_mode:
            return self.bypass_forward(x, self.multiplier)
        return self._fused(x, args, kwargs, self._native_spec, self._assemble)
