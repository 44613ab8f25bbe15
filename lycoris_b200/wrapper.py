"""Generic network wrapper: discovers target layers of any ``nn.Module``, attaches one adapter per
layer and manages their life-cycle (API contract: reference lycoris/wrapper.py:64-648).

Host-side only — no kernels here.  The discovery rules (class-name / regex / fnmatch targets,
per-class and per-name algo overrides, LoRA naming ``"{prefix}_{path with . -> _}"``) are restated
in :class:`_AdapterFactory`, shared with the kohya adapter in ``kohya.py``; tests pin names, order,
classes and shapes against fixtures generated from the reference itself.
"""

from __future__ import annotations

import fnmatch
import logging
import os
import re
from typing import Any

import torch
import torch.nn as nn

from .config import PRESET
from .logging import logger
from .modules import get_module, make_module
from .modules.dylora import DyLoraModule
from .modules.ia3 import IA3Module
from .modules.locon import LoConModule
from .modules.loha import LohaModule
from .modules.lokr import LokrModule
from .utils import str_bool
from .utils.preset import read_preset

VALID_PRESET_KEYS = [
    "enable_conv",
    "target_module",
    "target_name",
    "module_algo_map",
    "name_algo_map",
    "lora_prefix",
    "use_fnmatch",
    "unet_target_module",
    "unet_target_name",
    "text_encoder_target_module",
    "text_encoder_target_name",
    "exclude_name",
]

# algo name -> adapter class.  "ia3" is an addition over the reference dict (wrapper.py:45-55),
# which forgets it and raises KeyError for algo=ia3.
network_module_dict = {
    "lora": LoConModule,
    "locon": LoConModule,
    "loha": LohaModule,
    "lokr": LokrModule,
    "dylora": DyLoraModule,
    "ia3": IA3Module,
}
# algorithms the reference offers that are outside this engine's scope
OUT_OF_SCOPE_ALGOS = ("glora", "full", "diag-oft", "boft")

deprecated_arg_dict = {
    "disable_conv_cp": "use_tucker",
    "use_cp": "use_tucker",
    "use_conv_cp": "use_tucker",
    "constrain": "constraint",
}


def _resolve_algo(algo_name):
    try:
        return network_module_dict[algo_name]
    except KeyError:
        if algo_name in OUT_OF_SCOPE_ALGOS:
            raise KeyError(
                f"algo {algo_name!r} is not part of the B200 adapter engine (in scope: "
                f"{sorted(set(network_module_dict))}); use the reference LyCORIS package for it"
            ) from None
        raise


def _translate_deprecated(kwargs):
    for key, value in list(kwargs.items()):
        if key in deprecated_arg_dict:
            logger.warning(f"{key} is deprecated. Please use {deprecated_arg_dict[key]} instead.", stacklevel=3)
            kwargs[deprecated_arg_dict[key]] = value


def _wants_tucker(kwargs):
    return str_bool(
        not kwargs.get("disable_conv_cp", True)
        or kwargs.get("use_conv_cp", False)
        or kwargs.get("use_cp", False)
        or kwargs.get("use_tucker", False)
    )


def _load_weights_file(file):
    if os.path.splitext(file)[1] == ".safetensors":
        from safetensors.torch import load_file

        return load_file(file)
    return torch.load(file, map_location="cpu")


class _AdapterFactory:
    """Creates adapters for the layers of one root model.

    ``dedupe_across_roots`` / ``fix_root_name`` / ``exclude`` reproduce the generic wrapper's extra
    rules (wrapper.py:356-468); the kohya network (kohya.py:417-496) runs with all three off.
    """

    def __init__(self, net, network_module, root_kwargs, train_norm=False, norm_modules=None, *,
                 dedupe_across_roots, fix_root_name, lora_prefix=None):
        self.net = net
        self.network_module = network_module
        self.root_kwargs = root_kwargs
        self.train_norm = train_norm
        self.norm_modules = norm_modules
        self.dedupe_across_roots = dedupe_across_roots
        self.fix_root_name = fix_root_name
        self.lora_prefix = lora_prefix

    # one layer -> one adapter (or None when the layer type / dims rule it out)
    def single(self, lora_name, module, algo_name, dim=None, alpha=None, use_tucker=None, **kwargs):
        net = self.net
        if use_tucker is None:
            use_tucker = net.use_tucker
        for k, v in self.root_kwargs.items():
            kwargs.setdefault(k, v)
        if self.train_norm and "Norm" in module.__class__.__name__:
            if self.norm_modules is None:
                raise NotImplementedError("train_norm: norm-layer adapters are out of scope for the B200 engine")
            return self.norm_modules(lora_name, module, net.multiplier, net.rank_dropout, net.module_dropout, **kwargs)
        if isinstance(module, nn.Linear) and net.lora_dim > 0:
            dim, alpha = dim or net.lora_dim, alpha or net.alpha
        elif isinstance(module, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
            k_size, *_ = module.kernel_size
            if k_size == 1 and net.lora_dim > 0:
                dim, alpha = dim or net.lora_dim, alpha or net.alpha
            elif net.conv_lora_dim > 0 or dim:
                dim, alpha = dim or net.conv_lora_dim, alpha or net.conv_alpha
            else:
                return None
        else:
            return None
        return _resolve_algo(algo_name)(
            lora_name, module, net.multiplier, dim, alpha, net.dropout, net.rank_dropout, net.module_dropout,
            use_tucker, **kwargs,
        )

    # every eligible layer below a class-matched module
    def subtree(self, prefix, root_module, algo, known, configs=None):
        configs = configs or {}
        loras = known if self.dedupe_across_roots else {}
        names = []
        algo_map = self.net.MODULE_ALGO_MAP
        for name, module in root_module.named_modules():
            cls_name = module.__class__.__name__
            if cls_name in algo_map and module is not root_module:
                # a nested class with its own algo/config: handle its whole subtree with that config
                nxt = algo_map[cls_name]
                sub_prefix = f"{prefix}_{name}" if (name or not self.fix_root_name) else prefix
                sub_loras, sub_names, sub_map = self.subtree(sub_prefix, module, nxt.get("algo", algo), loras, nxt)
                if self.dedupe_across_roots:
                    loras = {**loras, **sub_map}
                for sub_name, sub_lora in zip(sub_names, sub_loras):
                    if sub_name not in loras:
                        loras[sub_name] = sub_lora
                    if sub_name not in names:
                        names.append(sub_name)
                continue
            lora_name = prefix + "." + name if name else prefix
            if self.fix_root_name and f"{self.lora_prefix}_." in lora_name:
                lora_name = lora_name.replace(f"{self.lora_prefix}_.", f"{self.lora_prefix}.")
            lora_name = lora_name.replace(".", "_")
            if lora_name in loras:
                continue
            lora = self.single(lora_name, module, algo, **configs)
            if lora is not None:
                loras[lora_name] = lora
                names.append(lora_name)
        return [loras[n] for n in names], names, loras

    # whole model: class-name targets open a subtree, name targets create a single adapter
    def build(self, prefix, root_module, target_modules, target_names=(), exclude_names=()):
        logger.info("Create LyCORIS Module")
        net = self.net
        loras, known = [], {}
        for name, module in root_module.named_modules():
            if name in exclude_names or any(net.match_fn(t, name) for t in exclude_names):
                continue
            cls_name = module.__class__.__name__
            name_hit = any(net.match_fn(t, name) for t in target_names)
            if cls_name in target_modules and not name_hit:
                cfg = net.MODULE_ALGO_MAP.get(cls_name, {})
                algo = cfg.get("algo", self.network_module) if cls_name in net.MODULE_ALGO_MAP else self.network_module
                made, _, sub_map = self.subtree(f"{prefix}_{name}", module, algo, known, cfg)
                if self.dedupe_across_roots:
                    known = {**known, **sub_map}
                loras.extend(made)
            elif name in target_names or name_hit:
                cfg = net.find_conf_for_name(name)
                if cfg is None:
                    cfg = net.MODULE_ALGO_MAP.get(cls_name, {})
                algo = cfg.get("algo", self.network_module)
                lora_name = (prefix + "." + name).replace(".", "_")
                if self.dedupe_across_roots and lora_name in known:
                    continue
                lora = self.single(lora_name, module, algo, **cfg)
                if lora is not None:
                    if self.dedupe_across_roots:
                        known[lora.lora_name] = lora
                    loras.append(lora)
        return loras


def _assert_unique(loras):
    seen = set()
    for lora in loras:
        assert lora.lora_name not in seen, f"duplicated lora name: {lora.lora_name}"
        seen.add(lora.lora_name)


def _algo_table(loras):
    table = {}
    for lora in loras:
        table[lora.__class__.__name__] = table.get(lora.__class__.__name__, 0) + 1
    return table


def _network_kwargs(kwargs, linear_dim, linear_alpha):
    """Parse the string-typed network args shared by create_lycoris and kohya.create_network."""
    conv_dim = int(kwargs.get("conv_dim", linear_dim) or linear_dim)
    conv_alpha = float(kwargs.get("conv_alpha", linear_alpha) or linear_alpha)
    parsed = dict(
        conv_lora_dim=conv_dim,
        conv_alpha=conv_alpha,
        dropout=float(kwargs.get("dropout", 0.0) or 0.0),
        rank_dropout=float(kwargs.get("rank_dropout", 0.0) or 0.0),
        module_dropout=float(kwargs.get("module_dropout", 0.0) or 0.0),
        use_tucker=_wants_tucker(kwargs),
        use_scalar=str_bool(kwargs.get("use_scalar", False)),
        network_module=(kwargs.get("algo", "lora") or "lora").lower(),
        train_norm=str_bool(kwargs.get("train_norm", False)),
        decompose_both=kwargs.get("decompose_both", False),
        factor=kwargs.get("factor", -1),
        block_size=int(kwargs.get("block_size", None) or 4),
        constraint=float(kwargs.get("constraint", None) or 0),
        rescaled=str_bool(kwargs.get("rescaled", False)),
        weight_decompose=str_bool(kwargs.get("dora_wd", False)),
        wd_on_out=str_bool(kwargs.get("wd_on_output", True)),
        full_matrix=str_bool(kwargs.get("full_matrix", False)),
        bypass_mode=str_bool(kwargs.get("bypass_mode", False)),
        unbalanced_factorization=str_bool(kwargs.get("unbalanced_factorization", False)),
    )
    if parsed["unbalanced_factorization"]:
        logger.info("Unbalanced factorization for LoKr is enabled")
    if parsed["bypass_mode"]:
        logger.info("Bypass mode is enabled")
    if parsed["weight_decompose"]:
        logger.info("Weight decomposition is enabled")
    if parsed["full_matrix"]:
        logger.info("Full matrix mode for LoKr is enabled")
    return parsed


def _resolve_preset(name):
    preset = PRESET[name] if name in PRESET else read_preset(name)
    assert preset is not None
    return preset


def create_lycoris(module, multiplier=1.0, linear_dim=4, linear_alpha=1, **kwargs):
    _translate_deprecated(kwargs)
    if linear_dim is None:
        linear_dim = 4  # default
    parsed = _network_kwargs(kwargs, linear_dim, linear_alpha)
    LycorisNetwork.apply_preset(_resolve_preset(kwargs.get("preset", "full")))
    logger.info(f"Using rank adaptation algo: {parsed['network_module']}")
    return LycorisNetwork(module, multiplier=multiplier, lora_dim=linear_dim, alpha=linear_alpha, **parsed)


def create_lycoris_from_weights(multiplier, file, module, weights_sd=None, **kwargs):
    if weights_sd is None:
        weights_sd = _load_weights_file(file)

    wanted = {key.split(".")[0]: None for key in weights_sd if "." in key}
    for name, sub in module.named_modules():
        lora_name = f"{LycorisNetwork.LORA_PREFIX}_{name}".replace(".", "_")
        if lora_name in wanted:
            wanted[lora_name] = sub

    level = logger.level
    logger.setLevel(logging.ERROR)
    network = LycorisNetwork(module, init_only=True)
    network.multiplier = multiplier
    network.loras = []
    logger.setLevel(level)

    logger.info("Loading Modules from state dict...")
    for lora_name, target in wanted.items():
        if target is None:
            continue
        lyco_type, params = get_module(weights_sd, lora_name)
        if lyco_type is None:
            continue
        lora = make_module(lyco_type, params, lora_name, target)
        if lora is not None:
            network.loras.append(lora)
            network.algo_table[lora.__class__.__name__] = network.algo_table.get(lora.__class__.__name__, 0) + 1
    logger.info(f"{len(network.loras)} Modules Loaded")

    for lora in network.loras:
        lora.multiplier = multiplier
    return network, weights_sd


class LycorisNetwork(torch.nn.Module):
    # class-level configuration, mutated by apply_preset (shared by all instances — like the reference)
    ENABLE_CONV = True
    TARGET_REPLACE_MODULE = ["Linear", "Conv1d", "Conv2d", "Conv3d", "GroupNorm", "LayerNorm"]
    TARGET_REPLACE_NAME = []
    LORA_PREFIX = "lycoris"
    MODULE_ALGO_MAP = {}
    NAME_ALGO_MAP = {}
    USE_FNMATCH = False
    TARGET_EXCLUDE_NAME = []

    _PRESET_ATTRS = {
        "enable_conv": "ENABLE_CONV",
        "target_module": "TARGET_REPLACE_MODULE",
        "target_name": "TARGET_REPLACE_NAME",
        "module_algo_map": "MODULE_ALGO_MAP",
        "name_algo_map": "NAME_ALGO_MAP",
        "lora_prefix": "LORA_PREFIX",
        "use_fnmatch": "USE_FNMATCH",
        "exclude_name": "TARGET_EXCLUDE_NAME",
    }

    @classmethod
    def apply_preset(cls, preset):
        for key in preset.keys():
            if key not in VALID_PRESET_KEYS:
                raise KeyError(f'Unknown preset key "{key}". Valid keys: {VALID_PRESET_KEYS}')
        for key, attr in cls._PRESET_ATTRS.items():
            if key in preset:
                setattr(cls, attr, preset[key])
        return cls

    def _common_init(self, multiplier, lora_dim, conv_lora_dim, alpha, conv_alpha, use_tucker, dropout,
                     rank_dropout, module_dropout):
        self.multiplier = multiplier
        self.lora_dim = lora_dim
        if not self.ENABLE_CONV:
            conv_lora_dim = 0
        self.conv_lora_dim = int(conv_lora_dim)
        if self.conv_lora_dim and self.conv_lora_dim != self.lora_dim:
            logger.info("Apply different lora dim for conv layer")
            logger.info(f"Conv Dim: {conv_lora_dim}, Linear Dim: {lora_dim}")
        elif self.conv_lora_dim == 0:
            logger.info("Disable conv layer")
        self.alpha = alpha
        self.conv_alpha = float(conv_alpha)
        if self.conv_lora_dim and self.alpha != self.conv_alpha:
            logger.info("Apply different alpha value for conv layer")
            logger.info(f"Conv alpha: {conv_alpha}, Linear alpha: {alpha}")
        if 1 >= dropout >= 0:
            logger.info(f"Use Dropout value: {dropout}")
        self.dropout = dropout
        self.rank_dropout = rank_dropout
        self.module_dropout = module_dropout
        self.use_tucker = use_tucker

    def __init__(
        self,
        module: nn.Module,
        multiplier=1.0,
        lora_dim=4,
        conv_lora_dim=4,
        alpha=1,
        conv_alpha=1,
        use_tucker=False,
        dropout=0,
        rank_dropout=0,
        module_dropout=0,
        network_module: str = "locon",
        norm_modules=None,
        train_norm=False,
        init_only=False,
        **kwargs,
    ) -> None:
        super().__init__()
        self.weights_sd = None
        if init_only:
            self.multiplier, self.lora_dim, self.alpha = 1, 0, 1
            self.conv_lora_dim, self.conv_alpha = 0, 1
            self.dropout = self.rank_dropout = self.module_dropout = 0
            self.use_tucker = False
            self.loras = []
            self.algo_table = {}
            return
        self._common_init(multiplier, lora_dim, conv_lora_dim, alpha, conv_alpha, use_tucker, dropout,
                          rank_dropout, module_dropout)

        factory = _AdapterFactory(self, network_module, kwargs, train_norm, norm_modules,
                                  dedupe_across_roots=True, fix_root_name=True,
                                  lora_prefix=LycorisNetwork.LORA_PREFIX)
        self.loras = factory.build(
            LycorisNetwork.LORA_PREFIX,
            module,
            list(set([*LycorisNetwork.TARGET_REPLACE_MODULE, *LycorisNetwork.MODULE_ALGO_MAP.keys()])),
            list(set([*LycorisNetwork.TARGET_REPLACE_NAME, *LycorisNetwork.NAME_ALGO_MAP.keys()])),
            exclude_names=LycorisNetwork.TARGET_EXCLUDE_NAME,
        )
        logger.info(f"create LyCORIS: {len(self.loras)} modules.")
        logger.info(f"module type table: {_algo_table(self.loras)}")
        _assert_unique(self.loras)

    # ---------------------------------------------------------------- name matching
    def match_fn(self, pattern: str, name: str) -> bool:
        if self.USE_FNMATCH:
            return fnmatch.fnmatch(name, pattern)
        return bool(re.match(pattern, name))

    def find_conf_for_name(self, name: str) -> dict[str, Any]:
        if name in self.NAME_ALGO_MAP:
            return self.NAME_ALGO_MAP[name]
        for key, value in self.NAME_ALGO_MAP.items():
            if self.match_fn(key, name):
                return value
        return None

    # -------------------------------------------------------------------- life-cycle
    def set_multiplier(self, multiplier):
        self.multiplier = multiplier
        for lora in self.loras:
            lora.multiplier = self.multiplier

    def load_weights(self, file):
        self.weights_sd = _load_weights_file(file)
        missing, unexpected = self.load_state_dict(self.weights_sd, strict=False)
        state = {}
        if missing:
            state["missing keys"] = missing
        if unexpected:
            state["unexpected keys"] = unexpected
        return state

    def apply_to(self):
        """Patch every target layer and register the adapters as sub-modules of the network."""
        for lora in self.loras:
            lora.apply_to()
            self.add_module(lora.lora_name, lora)
        if self.weights_sd:
            # missing keys are fine: a fresh adapter is a no-op (its second factor is zero)
            info = self.load_state_dict(self.weights_sd, False)
            logger.info(f"weights are loaded: {info}")

    def is_mergeable(self):
        return True

    def restore(self):
        for lora in self.loras:
            lora.restore()

    def merge_to(self, weight=1.0):
        for lora in self.loras:
            lora.merge_to(weight)

    def onfly_merge(self, weight=1.0):
        for lora in self.loras:
            lora.onfly_merge(weight)

    def onfly_restore(self):
        for lora in self.loras:
            lora.onfly_restore()

    def _max_norm(self, loras, max_norm_value, device):
        key_scaled, norms = 0, []
        for lora in loras:
            scaled, norm = lora.apply_max_norm(max_norm_value, device)
            if scaled is None:
                continue
            norms.append(norm)
            key_scaled += scaled
        if key_scaled == 0:
            return key_scaled, 0, 0
        return key_scaled, sum(norms) / len(norms), max(norms)

    def apply_max_norm_regularization(self, max_norm_value, device):
        return self._max_norm(self.loras, max_norm_value, device)

    def enable_gradient_checkpointing(self):
        # not supported; mark modules like the reference does
        def mark(m):
            if isinstance(m, torch.nn.Module):
                m.grad_ckpt = True

        self.apply(mark)

    def prepare_optimizer_params(self, lr):
        self.requires_grad_(True)
        params = []
        for lora in self.loras:
            params.extend(lora.parameters())
        group = {"params": params}
        if lr is not None:
            group["lr"] = lr
        return [group]

    def prepare_grad_etc(self, *args):
        self.requires_grad_(True)

    def on_epoch_start(self, *args):
        self.train()

    def get_trainable_params(self, *args):
        return self.parameters()

    def _export_state_dict(self, dtype):
        state_dict = self.state_dict()
        if dtype is not None:
            for key in list(state_dict.keys()):
                state_dict[key] = state_dict[key].detach().clone().to("cpu").to(dtype)
        return state_dict

    def save_weights(self, file, dtype, metadata):
        if metadata is not None and len(metadata) == 0:
            metadata = None
        state_dict = self._export_state_dict(dtype)
        if os.path.splitext(file)[1] == ".safetensors":
            from safetensors.torch import save_file

            save_file(state_dict, file, metadata if metadata is not None else {})
        else:
            torch.save(state_dict, file)

    # ------------------------------------------------------------- data parallel (new)
    def attach_data_parallel(self, process_group=None, bucket_dtype=None, overlap=True):
        """B200 addition: flat adapter-gradient arena + one NCCL all-reduce per step over NVLink
        (the reference has no collective; kohya gets DDP from accelerate).  See engine/ddp.py."""
        from .engine.ddp import FlatGradAllReduce

        self._dp = FlatGradAllReduce(list(self.parameters()), process_group, bucket_dtype, overlap)
        return self._dp
