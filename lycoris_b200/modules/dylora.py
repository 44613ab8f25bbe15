"""DyLoRA adapter: LoRA whose rank is re-drawn every forward in blocks of ``block_size``
(reference lycoris/modules/dylora.py).

Per call ``b ~ U{0..blocks-1}`` from Python's ``random`` (host RNG, same stream as the
reference); blocks ``< b`` enter frozen (``.data``), block ``b`` is trainable,
``dW = up[:, :(b+1)·bs] · (down[:(b+1)·bs] · alpha/(b+1) · mult)``.  The engine runs it on the LoCon
kernels with a runtime rank.  Reference quirks kept: the divisor uses the
``alpha`` *buffer*, ``load_state_dict`` is a no-op, bypass mode is broken upstream and refused here.
"""

import math
import random

import torch
import torch.nn as nn

from ..utils import product
from .base import LycorisBaseModule


class DyLoraModule(LycorisBaseModule):
    support_module = {"linear", "conv1d", "conv2d", "conv3d"}

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
        block_size=4,
        use_scalar=False,
        rank_dropout_scale=False,
        weight_decompose=False,
        bypass_mode=None,
        rs_lora=False,
        train_on_input=False,
        **kwargs,
    ):
        """if alpha == 0 or None, alpha is rank (no scaling)."""
        super().__init__(
            lora_name, org_module, multiplier, dropout, rank_dropout, module_dropout, rank_dropout_scale, bypass_mode
        )
        if self.module_type not in self.support_module:
            raise ValueError(f"{self.module_type} is not supported in IA^3 algo.")
        assert lora_dim % block_size == 0, "lora_dim must be a multiple of block_size"
        self.block_count = lora_dim // block_size
        self.block_size = block_size
        self.lora_dim = lora_dim

        out_dim, flat_in = self.shape[0], product(self.shape[1:])
        self.up_list = nn.ParameterList([torch.empty(out_dim, block_size) for _ in range(self.block_count)])
        self.down_list = nn.ParameterList([torch.empty(block_size, flat_in) for _ in range(self.block_count)])

        if isinstance(alpha, torch.Tensor):
            alpha = alpha.detach().float().numpy()
        alpha = lora_dim if alpha is None or alpha == 0 else alpha
        self.scale = alpha / self.lora_dim
        self.register_buffer("alpha", torch.tensor(alpha))
        self._alpha_host = float(alpha)

        for v in self.down_list:
            torch.nn.init.kaiming_uniform_(v, a=math.sqrt(5))
        for v in self.up_list:
            torch.nn.init.zeros_(v)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        return

    def custom_state_dict(self):
        return {
            "alpha": self.alpha,
            "lora_up.weight": nn.Parameter(torch.concat(list(self.up_list), dim=1)),
            "lora_down.weight": nn.Parameter(
                torch.concat(list(self.down_list)).reshape(self.lora_dim, -1, *self.shape[2:])
            ),
        }

    def _blocks(self, rank):
        """Concatenate frozen blocks ``< b`` and the live block ``b``; returns (down, up, b)."""
        b = math.ceil(rank / self.block_size)
        down = torch.concat([p.data for p in self.down_list[:b]] + list(self.down_list[b : b + 1]))
        up = torch.concat([p.data for p in self.up_list[:b]] + list(self.up_list[b : b + 1]), dim=1)
        return down, up, b

    def get_weight(self, rank):
        down, up, b = self._blocks(rank)
        return down, up, self.alpha / (b + 1)

    def get_random_rank_weight(self):
        self._forbid_capture("DyLoRA's per-step rank sample (random.randint)")
        b = random.randint(0, self.block_count - 1)
        return self.get_weight(b * self.block_size)

    def get_diff_weight(self, multiplier=1, shape=None, device=None, rank=None):
        down, up, scale = self.get_random_rank_weight() if rank is None else self.get_weight(rank)
        w = up @ (down * (scale * multiplier))
        if device is not None:
            w = w.to(device)
        return w.view(shape if shape is not None else self.shape), None

    def get_merged_weight(self, multiplier=1, shape=None, device=None, rank=None):
        diff, _ = self.get_diff_weight(multiplier, shape, device, rank)
        return diff + self.org_weight, None

    def bypass_forward_diff(self, x, scale=1, rank=None):
        raise NotImplementedError(
            "DyLoRA bypass mode is broken in the reference (dylora.py:130-138: undefined `gamma`, "
            "views that assume full rank); it is not reproduced. Use the default rebuild mode."
        )

    def bypass_forward(self, x, scale=1, rank=None):
        return self.org_forward(x) + self.bypass_forward_diff(x, scale, rank)

    def _native_spec(self):
        from ..engine.kernels import ALGO_DYLORA
        from ..engine.ops import NativeSpec

        self._forbid_capture("DyLoRA's per-step rank sample (random.randint)")
        b = random.randint(0, self.block_count - 1)  # same host RNG draw as get_random_rank_weight
        down, up, b = self._blocks(b * self.block_size)
        return NativeSpec(
            algo=ALGO_DYLORA,
            factors=(up, down),
            rank=(b + 1) * self.block_size,
            m_in=self._alpha_host / (b + 1) * float(self.multiplier),
        )

    def _assemble(self, base_weight):
        return self.get_merged_weight(multiplier=self.multiplier)[0].to(base_weight.dtype)

    def forward(self, x, *args, **kwargs):
        if self._module_dropped():
            return self.org_forward(x, *args, **kwargs)
        if self.bypass_mode:
            return self.bypass_forward(x, self.multiplier)
        return self._fused(x, args, kwargs, self._native_spec, self._assemble)
