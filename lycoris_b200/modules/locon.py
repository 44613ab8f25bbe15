"""LoCon / LoRA adapter:  dW = up · down · (alpha / r)      (reference lycoris/modules/locon.py).

Parameter names, shapes and initialisation follow the reference so checkpoints interchange:
``lora_down.weight [r, K(,kh,kw)]`` kaiming-uniform(a=√5), ``lora_up.weight [N, r(,1,1)]`` zeros,
buffer ``alpha``; optional Tucker core ``lora_mid``, DoRA magnitude ``dora_scale``, trainable ``scalar``.
"""

import math
from functools import lru_cache

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..functional.general import rebuild_tucker
from ..logging import logger
from .base import LycorisBaseModule


@lru_cache(maxsize=None)
def log_wd():
    return logger.warning(
        "Using weight_decompose=True with LoRA (DoRA) will ignore network_dropout."
        "Only rank dropout and module dropout will be applied"
    )


class LoConModule(LycorisBaseModule):
    name = "locon"
    support_module = {"linear", "conv1d", "conv2d", "conv3d"}
    weight_list = ["lora_up.weight", "lora_down.weight", "lora_mid.weight", "alpha", "dora_scale"]
    weight_list_det = ["lora_up.weight"]

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
        rank_dropout_scale=False,
        weight_decompose=False,
        wd_on_out=True,
        bypass_mode=None,
        rs_lora=False,
        **kwargs,
    ):
        """if alpha == 0 or None, alpha is rank (no scaling)."""
        super().__init__(
            lora_name, org_module, multiplier, dropout, rank_dropout, module_dropout, rank_dropout_scale, bypass_mode
        )
        if self.module_type not in self.support_module:
            raise ValueError(f"{self.module_type} is not supported in LoRA/LoCon algo.")
        self.lora_dim = lora_dim
        self.tucker = False
        self.rs_lora = rs_lora

        if self.module_type.startswith("conv"):
            self.isconv = True
            in_dim, out_dim = org_module.in_channels, org_module.out_channels
            k_size, stride, padding = org_module.kernel_size, org_module.stride, org_module.padding
            self.down_op = self.up_op = self.op
            self.tucker = bool(use_tucker) and any(i != 1 for i in k_size)
            if self.tucker:
                # Tucker: 1x1 down, k x k core at rank r, 1x1 up
                self.lora_down = self.module(in_dim, lora_dim, 1, bias=False)
                self.lora_mid = self.module(lora_dim, lora_dim, k_size, stride, padding, bias=False)
            else:
                self.lora_down = self.module(in_dim, lora_dim, k_size, stride, padding, bias=False)
            self.lora_up = self.module(lora_dim, out_dim, 1, bias=False)
        elif isinstance(org_module, nn.Linear):
            self.isconv = False
            self.down_op = self.up_op = F.linear
            self.lora_down = nn.Linear(org_module.in_features, lora_dim, bias=False)
            self.lora_up = nn.Linear(lora_dim, org_module.out_features, bias=False)
        else:
            raise NotImplementedError

        self._init_dora(org_module, weight_decompose, wd_on_out)

        if dropout:
            self.dropout = nn.Dropout(dropout)
            if self.wd:
                log_wd()
        else:
            self.dropout = nn.Identity()

        alpha, r_factor = self._init_alpha(alpha, lora_dim, rs_lora)
        self.scale = alpha / r_factor
        self.register_buffer("alpha", torch.tensor(alpha * (lora_dim / r_factor)))
        self._init_scalar(use_scalar)

        torch.nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        if use_scalar:
            torch.nn.init.kaiming_uniform_(self.lora_up.weight, a=math.sqrt(5))
        else:
            torch.nn.init.constant_(self.lora_up.weight, 0)
        if self.tucker:
            torch.nn.init.kaiming_uniform_(self.lora_mid.weight, a=math.sqrt(5))

    @classmethod
    def make_module_from_state_dict(cls, lora_name, orig_module, up, down, mid, alpha, dora_scale):
        module = cls(
            lora_name, orig_module, 1, down.size(0), float(alpha),
            use_tucker=mid is not None, weight_decompose=dora_scale is not None,
        )
        module.lora_up.weight.data.copy_(up)
        module.lora_down.weight.data.copy_(down)
        if mid is not None:
            module.lora_mid.weight.data.copy_(mid)
        if dora_scale is not None:
            module.dora_scale.copy_(dora_scale)
        return module

    def load_weight_hook(self, module: nn.Module, incompatible_keys):
        self._reset_scalar_after_load(incompatible_keys)

    def custom_state_dict(self):
        destination = {}
        if self.wd:
            destination["dora_scale"] = self.dora_scale
        destination["alpha"] = self.alpha
        destination["lora_up.weight"] = self.lora_up.weight * self.scalar
        destination["lora_down.weight"] = self.lora_down.weight
        if self.tucker:
            destination["lora_mid.weight"] = self.lora_mid.weight
        return destination

    # ------------------------------------------------------------------ dW (PyTorch ops)
    def make_weight(self, device=None):
        """up·down (or the Tucker rebuild) shaped like the base weight, times ``scalar``; the
        cold-path builder used by merge / max-norm and by the option variants of forward."""
        wa = self.lora_up.weight.to(device)
        wb = self.lora_down.weight.to(device)
        if self.tucker:
            weight = rebuild_tucker(self.lora_mid.weight, wa.view(wa.size(0), -1).transpose(0, 1), wb.view(wb.size(0), -1))
        else:
            weight = wa.view(wa.size(0), -1) @ wb.view(wb.size(0), -1)
        weight = weight.view(self.shape)
        if self.training and self.rank_dropout:
            weight = self._rank_drop_rows(weight, device)
        return weight * self.scalar.to(device)

    def get_diff_weight(self, multiplier=1, shape=None, device=None):
        cold = self.tucker or isinstance(self.scalar, nn.Parameter)  # outside the delta kernel's scope: host ops
        eng = None if cold else self._delta_via_engine(self._scalar_host(), self.scale * multiplier)
        diff = eng[0] if eng is not None else self.make_weight(device=device) * (self.scale * multiplier)
        if shape is not None:
            diff = diff.view(shape)
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
        cold = self.tucker or isinstance(self.scalar, nn.Parameter)
        eng = None if cold else self._delta_via_engine(self._scalar_host(), self.scale, want_out=False, want_norm=True)
        if eng is not None:
            orig_norm = eng[1].sqrt().to(self.lora_up.weight.dtype)  # ||dW||_F reduced inside the delta kernel
        else:
            orig_norm = self.make_weight(device).norm() * self.scale
        norm = torch.clamp(orig_norm, max_norm / 2)
        desired = torch.clamp(norm, max=max_norm)
        ratio = desired.cpu() / norm.cpu()
        scaled = norm != desired
        if scaled:
            self.scalar *= ratio
            self._scalar_cache = None
        return scaled, orig_norm * ratio

    # ------------------------------------------------------------------------- bypass
    def bypass_forward_diff(self, x, scale=1):
        mid = self.lora_down(x)
        if self.tucker:
            mid = self.lora_mid(mid)
        if self.rank_dropout and self.training:
            drop = (torch.rand(self.lora_dim, device=mid.device) > self.rank_dropout).to(mid.dtype)
            if self.rank_dropout_scale:
                drop /= drop.mean()
            drop = drop.view(1, -1, 1, 1) if x.dim() == 4 else drop.view(*[1] * (x.dim() - 1), -1)
            mid = mid * drop
        return self.dropout(self.lora_up(mid) * self.scalar * self.scale * scale)

    def bypass_forward(self, x, scale=1):
        return self.org_forward(x) + self.bypass_forward_diff(x, scale=scale)

    # ------------------------------------------------------------------------ forward
    def _native_spec(self):
        from ..engine.ops import NativeSpec
        from ..engine.kernels import ALGO_LOCON

        if self.tucker or isinstance(self.scalar, nn.Parameter) or (self.training and self.rank_dropout):
            return None
        up, down = self.lora_up.weight, self.lora_down.weight
        return NativeSpec(
            algo=ALGO_LOCON,
            factors=(up.view(up.size(0), -1), down.view(down.size(0), -1)),
            rank=self.lora_dim,
            m_pre=self._scalar_host(),
            m_post1=float(self.scale),
            m_post2=1.0 if self.wd else float(self.multiplier),
            dora=(self.dora_scale, self.wd_on_out, float(self.multiplier)) if self.wd else None,
        )

    def _assemble(self, base_weight):
        diff = self.make_weight(base_weight.device).to(base_weight.dtype) * self.scale
        if self.wd:
            return self.apply_weight_decompose(base_weight + diff, self.multiplier).to(base_weight.dtype)
        return base_weight + diff * self.multiplier

    def forward(self, x, *args, **kwargs):
        if self._module_dropped():
            return self.org_forward(x, *args, **kwargs)
        if self.bypass_mode:
            return self.bypass_forward(x, scale=self.multiplier)
        return self._fused(x, args, kwargs, self._native_spec, self._assemble)
