"""LoHa adapter:  dW = (w1a · w1b) ⊙ (w2a · w2b) · (alpha / r)     (reference lycoris/modules/loha.py,
lycoris/functional/loha.py).

The reference's ``HadaWeight`` autograd function recomputes both rank-r products in backward to
save memory; here the Hadamard tile is assembled on the fly inside the merge kernel and the four
factor gradients come from one pass over dW' (``lyco_factor_grads``), never caching N×K products.
"""

import torch
import torch.nn as nn

from ..functional.loha import diff_weight as loha_diff_weight
from .base import LycorisBaseModule


class LohaModule(LycorisBaseModule):
    name = "loha"
    support_module = {"linear", "conv1d", "conv2d", "conv3d"}
    weight_list = ["hada_w1_a", "hada_w1_b", "hada_w2_a", "hada_w2_b", "hada_t1", "hada_t2", "alpha", "dora_scale"]
    weight_list_det = ["hada_w1_a"]

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
        super().__init__(
            lora_name, org_module, multiplier, dropout, rank_dropout, module_dropout, rank_dropout_scale, bypass_mode
        )
        if self.module_type not in self.support_module:
            raise ValueError(f"{self.module_type} is not supported in LoHa algo.")
        self.lora_name = lora_name
        self.lora_dim = lora_dim
        self.tucker = False
        self.rs_lora = rs_lora

        w_shape = self.shape
        if self.module_type.startswith("conv"):
            in_dim, out_dim, k_size = org_module.in_channels, org_module.out_channels, org_module.kernel_size
            self.shape = (out_dim, in_dim, *k_size)
            self.tucker = bool(use_tucker) and any(i != 1 for i in k_size)
            if self.tucker:
                w_shape = (out_dim, in_dim, *k_size)
            else:
                w_shape = (out_dim, in_dim * torch.tensor(k_size).prod().item())

        if self.tucker:
            # factor matrices are [r, dim] (mode products against the k x k cores t1 / t2)
            for tag in ("1", "2"):
                setattr(self, f"hada_t{tag}", nn.Parameter(torch.empty(lora_dim, lora_dim, *w_shape[2:])))
                setattr(self, f"hada_w{tag}_a", nn.Parameter(torch.empty(lora_dim, w_shape[0])))
                setattr(self, f"hada_w{tag}_b", nn.Parameter(torch.empty(lora_dim, w_shape[1])))
        else:
            for tag in ("1", "2"):
                setattr(self, f"hada_w{tag}_a", nn.Parameter(torch.empty(w_shape[0], lora_dim)))
                setattr(self, f"hada_w{tag}_b", nn.Parameter(torch.empty(lora_dim, w_shape[1])))

        self._init_dora(org_module, weight_decompose, wd_on_out)
        if self.dropout:
            print("[WARN]LoHa/LoKr haven't implemented normal dropout yet.")

        alpha, r_factor = self._init_alpha(alpha, lora_dim, rs_lora)
        self.scale = alpha / r_factor
        self.register_buffer("alpha", torch.tensor(alpha * (lora_dim / r_factor)))
        self._init_scalar(use_scalar)

        if self.tucker:
            torch.nn.init.normal_(self.hada_t1, std=0.1)
            torch.nn.init.normal_(self.hada_t2, std=0.1)
        torch.nn.init.normal_(self.hada_w1_b, std=1)
        torch.nn.init.normal_(self.hada_w1_a, std=0.1)
        torch.nn.init.normal_(self.hada_w2_b, std=1)
        if use_scalar:
            torch.nn.init.normal_(self.hada_w2_a, std=0.1)
        else:
            torch.nn.init.constant_(self.hada_w2_a, 0)

    @classmethod
    def make_module_from_state_dict(cls, lora_name, orig_module, w1a, w1b, w2a, w2b, t1, t2, alpha, dora_scale):
        module = cls(
            lora_name, orig_module, 1, w1b.size(0), float(alpha),
            use_tucker=t1 is not None, weight_decompose=dora_scale is not None,
        )
        module.hada_w1_a.copy_(w1a)
        module.hada_w1_b.copy_(w1b)
        module.hada_w2_a.copy_(w2a)
        module.hada_w2_b.copy_(w2b)
        if t1 is not None:
            module.hada_t1.copy_(t1)
            module.hada_t2.copy_(t2)
        if dora_scale is not None:
            module.dora_scale.copy_(dora_scale)
        return module

    def load_weight_hook(self, module: nn.Module, incompatible_keys):
        self._reset_scalar_after_load(incompatible_keys)

    def custom_state_dict(self):
        destination = {"alpha": self.alpha}
        if self.wd:
            destination["dora_scale"] = self.dora_scale
        destination["hada_w1_a"] = self.hada_w1_a * self.scalar
        destination["hada_w1_b"] = self.hada_w1_b
        destination["hada_w2_a"] = self.hada_w2_a
        destination["hada_w2_b"] = self.hada_w2_b
        if self.tucker:
            destination["hada_t1"] = self.hada_t1
            destination["hada_t2"] = self.hada_t2
        return destination

    # ------------------------------------------------------------------ dW (PyTorch ops)
    def get_weight(self, shape):
        gamma = torch.tensor(self.scale, dtype=self.hada_w1_b.dtype, device=self.hada_w1_b.device)
        cores = (self.hada_t1, self.hada_t2) if self.tucker else (None, None)
        weight = loha_diff_weight(self.hada_w1_b, self.hada_w1_a, self.hada_w2_b, self.hada_w2_a, *cores, gamma=gamma)
        if shape is not None:
            weight = weight.reshape(shape)
        if self.training and self.rank_dropout:
            weight = self._rank_drop_rows(weight)
        return weight

    def get_diff_weight(self, multiplier=1, shape=None, device=None):
        # NB: like the reference (loha.py:229-230) this applies `scale` on top of get_weight, which
        # already carries it — merge_to therefore differs from the training forward by 1/scale.
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
        cold = self.tucker or isinstance(self.scalar, nn.Parameter)
        eng = None if cold else self._delta_via_engine(self.scale, self._scalar_host(), want_out=False, want_norm=True)
        if eng is not None:
            orig_norm = eng[1].sqrt().to(self.hada_w1_a.dtype)  # ||dW||_F reduced inside the delta kernel
        else:
            orig_norm = (self.get_weight(self.shape) * self.scalar).norm()
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
        # no cheap activation-side form exists for a Hadamard of two low-rank products
        diff_weight = self.get_weight(self.shape) * self.scalar * scale
        return self.drop(self.op(x, diff_weight, **self.kw_dict))

    def bypass_forward(self, x, scale=1):
        return self.org_forward(x) + self.bypass_forward_diff(x, scale=scale)

    # ------------------------------------------------------------------------ forward
    def _native_spec(self):
        from ..engine.kernels import ALGO_LOHA
        from ..engine.ops import NativeSpec

        if self.tucker or isinstance(self.scalar, nn.Parameter) or (self.training and self.rank_dropout):
            return None
        return NativeSpec(
            algo=ALGO_LOHA,
            factors=(self.hada_w1_a, self.hada_w1_b, self.hada_w2_a, self.hada_w2_b),
            rank=self.lora_dim,
            m_pre=float(self.scale),
            m_post1=self._scalar_host(),
            m_post2=1.0 if self.wd else float(self.multiplier),
            dora=(self.dora_scale, self.wd_on_out, float(self.multiplier)) if self.wd else None,
        )

    def _assemble(self, base_weight):
        diff = self.get_weight(self.shape).to(base_weight.dtype) * self.scalar
        if self.wd:
            return self.apply_weight_decompose(base_weight + diff, self.multiplier).to(base_weight.dtype)
        return base_weight + diff * self.multiplier

    def forward(self, x: torch.Tensor, *args, **kwargs):
        if self._module_dropped():
            return self.org_forward(x, *args, **kwargs)
        if self.bypass_mode:
            return self.bypass_forward(x, scale=self.multiplier)
        return self._fused(x, args, kwargs, self._native_spec, self._assemble)
