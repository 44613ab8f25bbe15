"""Adapter base class: the per-layer object a LyCORIS network attaches to one ``nn.Linear`` /
``nn.ConvNd`` (API contract: reference lycoris/modules/base.py:64-398, docs/API.md:5-29).

What is preserved verbatim is the *interface*: constructor argument order, attribute names
(``lora_name``, ``multiplier``, ``org_module`` held in a list so it is not registered, ``org_forward``,
``shape``, ``op``, ``kw_dict``, ``module_type``), the forward monkey-patch with wrapper stacking,
``merge_to`` / ``onfly_merge`` / ``parametrize`` and the custom ``state_dict``.

What is new is where ``forward`` goes: on a CUDA tensor it calls the fused engine
(``lycoris_b200.engine.ops``): one merged-weight pass + ONE tcgen05 contraction in forward, two
(dX, dW') in backward, instead of the reference's 2 + 3 library GEMMs.  There is no CPU path.
"""

from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.utils.parametrize as parametrize

from ..utils.quant import QuantLinears, log_bypass, log_suspect

_CONV_TYPES = ((nn.Conv1d, "conv1d", F.conv1d), (nn.Conv2d, "conv2d", F.conv2d), (nn.Conv3d, "conv3d", F.conv3d))


class ModuleCustomSD(nn.Module):
    """``nn.Module`` whose ``state_dict`` can be replaced by ``custom_state_dict()`` (checkpoint
    wire format: scalar folded into the first factor, DyLoRA blocks concatenated, ...)."""

    def __init__(self):
        super().__init__()
        self._register_load_state_dict_pre_hook(self.load_weight_prehook)
        self.register_load_state_dict_post_hook(self.load_weight_hook)

    def load_weight_prehook(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        pass

    def load_weight_hook(self, module, incompatible_keys):
        pass

    def custom_state_dict(self):
        return None

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        # positional (destination, prefix, keep_vars) is still accepted by nn.Module
        if args:
            destination = args[0] if destination is None else destination
            if len(args) > 1 and prefix == "":
                prefix = args[1]
            if len(args) > 2 and keep_vars is False:
                keep_vars = args[2]
        if destination is None:
            destination = OrderedDict()
            destination._metadata = OrderedDict()
        if hasattr(destination, "_metadata"):
            destination._metadata[prefix[:-1]] = dict(version=self._version)
        custom = self.custom_state_dict()
        if custom is None:
            return super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        for key, value in custom.items():
            destination[f"{prefix}{key}"] = value
        return destination


def _describe(org_module):
    """(module_type, shape, op, out_dim, kw_dict) of a supported base layer, else None."""
    if isinstance(org_module, nn.Linear):
        return "linear", (org_module.out_features, org_module.in_features), F.linear, org_module.out_features, {}
    for cls, tag, op in _CONV_TYPES:
        if isinstance(org_module, cls):
            kw = {
                "stride": org_module.stride,
                "padding": org_module.padding,
                "dilation": org_module.dilation,
                "groups": org_module.groups,
            }
            shape = (org_module.out_channels, org_module.in_channels, *org_module.kernel_size)
            return tag, shape, op, org_module.out_channels, kw
    if isinstance(org_module, nn.LayerNorm):
        kw = {"normalized_shape": org_module.normalized_shape, "eps": org_module.eps}
        return "layernorm", tuple(org_module.normalized_shape), F.layer_norm, org_module.normalized_shape[0], kw
    if isinstance(org_module, nn.GroupNorm):
        kw = {"num_groups": org_module.num_groups, "eps": org_module.eps}
        return "groupnorm", (org_module.num_channels,), F.group_norm, org_module.num_channels, kw
    return None


class LycorisBaseModule(ModuleCustomSD):
    name: str
    dtype_tensor: torch.Tensor
    support_module = {}
    weight_list = []
    weight_list_det = []

    def __init__(
        self,
        lora_name,
        org_module: nn.Module,
        multiplier=1.0,
        dropout=0.0,
        rank_dropout=0.0,
        module_dropout=0.0,
        rank_dropout_scale=False,
        bypass_mode=None,
        **kwargs,
    ):
        """if alpha == 0 or None, alpha is rank (no scaling)."""
        super().__init__()
        self.lora_name = lora_name
        self.module = type(org_module)
        info = _describe(org_module)
        self.not_supported = info is None
        if info is None:
            self.module_type = "unknown"
        else:
            self.module_type, self.shape, self.op, self.dim, self.kw_dict = info
            if self.module_type == "groupnorm":
                self.group_num = org_module.num_groups

        self.register_buffer("dtype_tensor", torch.tensor(0.0), persistent=False)

        # quantised / foreign Linear subclasses cannot expose a dense weight: force bypass
        self.is_quant = False
        if isinstance(org_module, QuantLinears):
            if not bypass_mode:
                log_bypass()
            self.is_quant = True
            bypass_mode = True
        if isinstance(org_module, nn.Linear) and org_module.__class__.__name__ != "Linear":
            if bypass_mode is None:
                log_suspect()
                bypass_mode = True
            if bypass_mode == True:  # noqa: E712 - mirrors the reference's truthiness rule
                self.is_quant = True
        self.bypass_mode = bypass_mode
        self.dropout = dropout
        self.rank_dropout = rank_dropout
        self.rank_dropout_scale = rank_dropout_scale
        self.module_dropout = module_dropout

        # dropout placement (reference base.py:183-193):
        #   bypass : WX + drop(B rank_drop(A X))   (LoCon)   |  WX + drop(dW X)        (others)
        #   rebuild: (W + B rank_drop(A)) X        (LoCon)   |  (W + rank_drop(dW)) X  (others)
        self.drop = nn.Identity() if dropout == 0 else nn.Dropout(dropout)
        self.rank_drop = nn.Identity() if rank_dropout == 0 else nn.Dropout(rank_dropout)

        self.multiplier = multiplier
        self.org_forward = org_module.forward
        self.org_module = [org_module]  # a list: keeps the base layer out of parameters()/state_dict()

    # ------------------------------------------------------------------ class API
    @classmethod
    def parametrize(cls, org_module, attr, *args, **kwargs):
        target = getattr(org_module, attr)
        kwargs["bypass_mode"] = False
        if target.dim() == 2:
            proxy = nn.Linear(target.shape[0], target.shape[1], bias=False)
        elif 3 <= target.dim() <= 5:
            conv_cls = {3: nn.Conv1d, 4: nn.Conv2d, 5: nn.Conv3d}[target.dim()]
            proxy = conv_cls(target.shape[0], target.shape[1], *target.shape[2:], bias=False)
        else:
            raise ValueError(f"cannot parametrize a {target.dim()}-d tensor")
        proxy.weight = target
        module_obj = cls("", proxy, *args, **kwargs)
        module_obj.forward = module_obj.parametrize_forward
        module_obj.to(target)
        parametrize.register_parametrization(org_module, attr, module_obj)
        return module_obj

    @classmethod
    def algo_check(cls, state_dict, lora_name):
        return any(f"{lora_name}.{k}" in state_dict for k in cls.weight_list_det)

    @classmethod
    def extract_state_dict(cls, state_dict, lora_name):
        return [state_dict.get(f"{lora_name}.{k}", None) for k in cls.weight_list]

    @classmethod
    def make_module_from_state_dict(cls, lora_name, orig_module, *weights):
        raise NotImplementedError

    # ------------------------------------------------------------------ properties
    @property
    def dtype(self):
        return self.dtype_tensor.dtype

    @property
    def device(self):
        return self.dtype_tensor.device

    @property
    def org_weight(self):
        return self.org_module[0].weight

    @org_weight.setter
    def org_weight(self, value):
        self.org_module[0].weight.data.copy_(value)

    def _current_weight(self):
        return self.org_module[0].weight.detach()

    def _current_bias(self):
        bias = self.org_module[0].bias
        return None if bias is None else bias.detach()

    # ------------------------------------------------------------- forward patching
    def apply_to(self, **kwargs):
        """Route ``org_module.forward`` through this adapter; adapters stack (LIFO)."""
        if self.not_supported:
            return
        module = self.org_module[0]
        if not hasattr(module, "_lycoris_original_forward"):
            module._lycoris_original_forward = module.forward
        stack = [w for w in getattr(module, "_lycoris_wrappers", []) if w is not self]
        self.org_forward = module.forward
        stack.append(self)
        module._lycoris_wrappers = stack
        module.forward = self.forward

    def restore(self):
        """Take this adapter out of the forward chain, re-linking the wrapper above it."""
        if self.not_supported:
            return
        module = self.org_module[0]
        pristine = getattr(module, "_lycoris_original_forward", self.org_forward)
        stack = list(getattr(module, "_lycoris_wrappers", []))
        if not stack:
            module.forward = pristine
            return
        if self not in stack:
            module.forward = stack[-1].forward
            return
        idx = stack.index(self)
        stack.pop(idx)
        if idx < len(stack):
            stack[idx].org_forward = self.org_forward
        if stack:
            module._lycoris_wrappers = stack
            module.forward = stack[-1].forward
        else:
            module.forward = pristine
            module.__dict__.pop("_lycoris_wrappers", None)
            module.__dict__.pop("_lycoris_original_forward", None)

    def _is_outermost_on_plain_forward(self):
        """True only when ``org_forward`` is the base layer's OWN class forward bound to that layer — no
        LyCORIS wrapper below us and no foreign instance patch (kohya ``networks.lora``, an accelerate
        offload hook, a custom forward).  Only then may the engine contract ``x . (W + dW)^T + b`` directly
        from ``org.weight``; anything else keeps ``org_forward(x)`` and adds the delta contraction, like
        the reference (base.py:264-287)."""
        module = self.org_module[0]
        below = self.org_forward
        return getattr(below, "__self__", None) is module and getattr(below, "__func__", None) is type(module).forward

    # ----------------------------------------------------------------------- merge
    def _retarget(self):
        first = next(self.parameters())
        return first.device, first.dtype

    def _write_bias(self, bias):
        if bias is None:
            return
        bias = bias.to(self.org_weight)
        if self.org_module[0].bias is not None:
            self.org_module[0].bias.data.copy_(bias)
        else:
            self.org_module[0].bias = nn.Parameter(bias)

    def merge_to(self, multiplier=1.0):
        if self.not_supported:
            return
        device, dtype = self._retarget()
        self.to(self.org_weight)
        weight, bias = self.get_merged_weight(multiplier, self.org_weight.shape, self.org_weight.device)
        self.org_weight = weight.to(self.org_weight)
        self._write_bias(bias)
        self.to(device, dtype)

    def onfly_merge(self, multiplier=1.0):
        if self.not_supported:
            return
        device, dtype = self._retarget()
        self.to(self.org_weight)
        self.cached_org_weight = self.org_weight.data.cpu()
        self.cached_org_bias = None
        weight, bias = self.get_merged_weight(multiplier, self.org_weight.shape, self.org_weight.device)
        self.org_weight = weight
        if bias is not None:
            had_bias = self.org_module[0].bias is not None
            self._write_bias(bias)
            if had_bias:
                self.cached_org_bias = self.org_module[0].bias.data.cpu()
        if self.org_module[0].bias is not None:
            self.org_module[0].bias = self.org_module[0].bias.to(self.org_weight)
        self.to(device, dtype)

    def onfly_restore(self):
        if self.not_supported:
            return
        self.org_weight = self.cached_org_weight.to(self.org_weight)
        if self.cached_org_bias is not None:
            self.org_module[0].bias.data.copy_(self.cached_org_bias.to(self.org_weight))
        del self.cached_org_weight
        del self.cached_org_bias

    # --------------------------------------------------------- algorithm interface
    def get_diff_weight(self, multiplier=1.0, shape=None, device=None):
        raise NotImplementedError

    def get_merged_weight(self, multiplier=1.0, shape=None, device=None):
        raise NotImplementedError

    @torch.no_grad()
    def apply_max_norm(self, max_norm, device=None):
        return None, None

    def bypass_forward_diff(self, x, scale=1):
        raise NotImplementedError

    def bypass_forward(self, x, scale=1):
        raise NotImplementedError

    def parametrize_forward(self, x: torch.Tensor, *args, **kwargs):
        return self.get_merged_weight(multiplier=self.multiplier, shape=x.shape, device=x.device)[0].to(x.dtype)

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    # ------------------------------------------------------------ shared machinery
    def _init_dora(self, org_module, weight_decompose, wd_on_out):
        """DoRA magnitude vector: row (or column) L2 norms of the base weight, fp32
        (reference locon.py:107-129 and its copies in loha.py / lokr.py)."""
        self.wd = weight_decompose
        self.wd_on_out = wd_on_out
        if not self.wd:
            return
        w = org_module.weight.cpu().clone().float()
        self.dora_norm_dims = w.dim() - 1
        ones = [1] * self.dora_norm_dims
        if self.wd_on_out:
            mag = torch.norm(w.reshape(w.shape[0], -1), dim=1, keepdim=True).reshape(w.shape[0], *ones)
        else:
            mag = torch.norm(w.transpose(1, 0).reshape(w.shape[1], -1), dim=1, keepdim=True)
            mag = mag.reshape(w.shape[1], *ones).transpose(1, 0)
        self.dora_scale = nn.Parameter(mag).float()

    def apply_weight_decompose(self, weight, multiplier=1):
        """Rescale ``weight`` so its row/column norms equal ``dora_scale`` (locon.py:239-260)."""
        weight = weight.to(self.dora_scale.dtype)
        ones = [1] * self.dora_norm_dims
        eps = torch.finfo(weight.dtype).eps
        if self.wd_on_out:
            norm = weight.reshape(weight.shape[0], -1).norm(dim=1).reshape(weight.shape[0], *ones) + eps
        else:
            norm = weight.transpose(0, 1).reshape(weight.shape[1], -1).norm(dim=1, keepdim=True)
            norm = norm.reshape(weight.shape[1], *ones).transpose(0, 1) + eps
        scale = self.dora_scale.to(weight.device) / norm
        if multiplier != 1:
            scale = multiplier * (scale - 1) + 1
        return weight * scale

    def _init_alpha(self, alpha, lora_dim, rs_lora=False):
        """scale = alpha / r (or alpha / sqrt(r)); the ``alpha`` buffer stores alpha * r / r_factor so
        that rs_lora round-trips through a checkpoint (locon.py:138-148)."""
        import math

        if isinstance(alpha, torch.Tensor):
            alpha = alpha.detach().float().numpy()  # bf16 tensors cannot go to numpy directly
        alpha = lora_dim if alpha is None or alpha == 0 else alpha
        return alpha, (math.sqrt(lora_dim) if rs_lora else lora_dim)

    def _init_scalar(self, use_scalar):
        if use_scalar:
            self.scalar = nn.Parameter(torch.tensor(0.0))
        else:
            self.register_buffer("scalar", torch.tensor(1.0), persistent=False)

    def _reset_scalar_after_load(self, incompatible_keys):
        """Checkpoints carry the scalar folded into the first factor: drop the missing-key report
        and reset the live scalar to 1 (locon.py:184-196)."""
        self._scalar_cache = None
        missing = incompatible_keys.missing_keys
        for key in [k for k in missing if "scalar" in k]:
            missing.remove(key)
        if isinstance(self.scalar, nn.Parameter):
            self.scalar.data.copy_(torch.ones_like(self.scalar))
        elif getattr(self, "scalar", None) is not None:
            self.scalar.copy_(torch.ones_like(self.scalar))
        else:
            self.register_buffer("scalar", torch.ones_like(self.scalar), persistent=False)

    def _scalar_host(self):
        """Host copy of the (non-trainable) scalar buffer; cached so the hot path never syncs.
        ``apply_max_norm`` and checkpoint loads invalidate it."""
        v = getattr(self, "_scalar_cache", None)
        if v is None:
            v = float(self.scalar)
            self._scalar_cache = v
        return v

    def _forbid_capture(self, what):
        """Host-side random draws (the reference's ``torch.rand(1)`` module-dropout coin, its CPU rank-dropout
        masks, DyLoRA's ``random.randint``) are evaluated ONCE when a step is captured into a CUDA graph and
        every replay would train the same frozen choice: refuse instead of silently freezing."""
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            raise RuntimeError(
                f"lycoris_b200: {what} of {self.lora_name!r} draws from the HOST random stream on every step "
                "(as the reference does) and would be frozen into the captured CUDA graph; run this step "
                "eagerly, or set the dropout to 0 / use a fixed-rank algorithm for graph-captured training")

    def _module_dropped(self):
        if not (self.module_dropout and self.training):
            return False
        self._forbid_capture("module_dropout")
        return bool(torch.rand(1) < self.module_dropout)

    def _rank_drop_rows(self, weight, device=None):
        """Bernoulli mask over output rows of dW (rebuild-mode rank dropout)."""
        if device is None or torch.device(device).type != "cuda":
            self._forbid_capture("rank_dropout (mask drawn on the CPU)")
        drop = (torch.rand(weight.size(0), device=device) > self.rank_dropout).to(weight.dtype)
        drop = drop.view(-1, *[1] * (weight.dim() - 1)).to(weight.device)
        if self.rank_dropout_scale:
            drop /= drop.mean()
        return weight * drop

    def _delta_via_engine(self, m_pre, m_post1, m_post2=1.0, shape=None, want_out=True, want_norm=False):
        """``(dW, sum dW^2)`` through the weight-side CUDA kernel (lyco_delta_weight) when this adapter lives on a
        CUDA device and its factors are in the kernel's scope; ``None`` otherwise (CPU tensors, Tucker cores, a
        trainable scalar, rank dropout in training mode: the host PyTorch path, as upstream)."""
        first = next(iter(self.parameters()), None)
        if first is None or not first.is_cuda or not hasattr(self, "_native_spec"):
            return None
        spec = self._native_spec()
        if spec is None:
            return None
        from ..engine import ops

        return ops.delta_weight(spec, float(m_pre), float(m_post1), float(m_post2), tuple(shape or self.shape),
                                want_out, want_norm)

    def _fused(self, x, args, kwargs, native_spec, assemble_fallback):
        """Dispatch the rebuild-mode forward to the engine.

        ``native_spec()`` returns (spec, factors) for the CUDA merge kernel or None when an option
        outside the kernel's scope is active (DoRA, Tucker, trainable scalar, rank dropout); then
        ``assemble_fallback(base_weight)`` builds W' with PyTorch ops on the GPU and only the dense
        contractions run in the engine.
        """
        from ..engine import ops

        return ops.adapter_forward(self, x, args, kwargs, native_spec, assemble_fallback)
