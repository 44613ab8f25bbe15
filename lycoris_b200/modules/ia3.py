"""(IA)^3 adapter: a learned per-channel scale on the output (or input) channels of the base
weight, ``W' = W ⊙ (1 + w·mult)``      (reference lycoris/modules/ia3.py).

Reference quirks kept on purpose: rebuild mode leaves the bias
unscaled while bypass mode scales it; ``apply_to`` does not participate in wrapper stacking.
Unlike the reference, ``"ia3"`` IS registered in ``network_module_dict`` here, and
``make_module_from_state_dict`` accepts the ``on_input`` entry its ``weight_list`` extracts
(the reference raises KeyError / TypeError at those two places).
"""

import torch
import torch.nn as nn

from .base import LycorisBaseModule


class IA3Module(LycorisBaseModule):
    name = "ia3"
    support_module = {"linear", "conv1d", "conv2d", "conv3d"}
    weight_list = ["weight", "on_input"]
    weight_list_det = ["on_input"]

    # positional contract of every adapter constructor (wrapper.py / kohya.py call it positionally)
    def __init__(self, lora_name, org_module: nn.Module, multiplier=1.0, lora_dim=4, alpha=1, dropout=0.0,
                 rank_dropout=0.0, module_dropout=0.0, use_tucker=False, use_scalar=False, rank_dropout_scale=False,
                 weight_decompose=False, bypass_mode=None, rs_lora=False, train_on_input=False, **kwargs):
        """if alpha == 0 or None, alpha is rank (no scaling)."""
        super().__init__(
            lora_name, org_module, multiplier, dropout, rank_dropout, module_dropout, rank_dropout_scale, bypass_mode
        )
        if self.module_type not in self.support_module:
            raise ValueError(f"{self.module_type} is not supported in IA^3 algo.")

        if self.module_type.startswith("conv"):
            self.isconv = True
            train_dim = org_module.in_channels if train_on_input else org_module.out_channels
            self.weight = nn.Parameter(torch.empty(1, train_dim, *(1 for _ in self.shape[2:])))
        else:
            train_dim = org_module.in_features if train_on_input else org_module.out_features
            self.weight = nn.Parameter(torch.empty(train_dim))

        torch.nn.init.constant_(self.weight, 0)
        self.train_input = train_on_input
        self.register_buffer("on_input", torch.tensor(int(train_on_input)))

    @classmethod
    def make_module_from_state_dict(cls, lora_name, orig_module, weight, on_input=None):
        train_on_input = bool(int(on_input)) if on_input is not None else False
        module = cls(lora_name, orig_module, 1, train_on_input=train_on_input)
        module.weight.data.copy_(weight)
        return module

    def apply_to(self):
        self.org_forward = self.org_module[0].forward
        self.org_module[0].forward = self.forward

    def make_weight(self, multiplier=1, shape=None, device=None, diff=False):
        """``W ⊙ (w·mult + 1)`` (or ``W ⊙ w·mult`` with ``diff``) broadcast on the trained axis."""
        weight = self.weight * multiplier + int(not diff)
        if self.train_input:
            out = self.org_weight * weight
        else:
            out = (self.org_weight.transpose(0, 1) * weight).transpose(0, 1)
        if shape is not None:
            out = out.view(shape)
        if device is not None:
            out = out.to(device)
        return out

    def get_diff_weight(self, multiplier=1, shape=None, device=None):
        return self.make_weight(multiplier=multiplier, shape=shape, device=device, diff=True), None

    def get_merged_weight(self, multiplier=1, shape=None, device=None):
        return self.make_weight(multiplier=multiplier, shape=shape, device=device), None

    def _bypass_forward(self, x, scale=1, diff=False):
        weight = self.weight * scale + int(not diff)
        if self.train_input:
            x = x * weight
        out = self.org_forward(x)
        if not self.train_input:
            out = out * weight
        return out

    def bypass_forward_diff(self, x, scale=1):
        return self._bypass_forward(x, scale, diff=True)

    def bypass_forward(self, x, scale=1):
        return self._bypass_forward(x, scale, diff=False)

    def _native_spec(self):
        from ..engine.kernels import ALGO_IA3
        from ..engine.ops import NativeSpec

        group = 1
        for s in self.shape[2:]:
            group *= s
        return NativeSpec(
            algo=ALGO_IA3,
            factors=(self.weight.reshape(-1),),
            on_input=int(self.train_input),
            ia3_group=group if self.train_input else 1,
            matmul_product=False,
            m_post2=float(self.multiplier),
        )

    def _assemble(self, base_weight):
        return self.get_merged_weight(multiplier=self.multiplier)[0].to(base_weight.device, dtype=base_weight.dtype)

    def forward(self, x, *args, **kwargs):
        if self._module_dropped():
            return self.org_forward(x, *args, **kwargs)
        if self.bypass_mode:
            return self.bypass_forward(x, self.multiplier)
        return self._fused(x, args, kwargs, self._native_spec, self._assemble)
