"""kohya-ss sd-scripts ``network_module`` protocol (API contract: reference lycoris/kohya.py:30-772).

``--network_module lycoris_b200.kohya`` makes sd-scripts call ``create_network`` /
``create_network_from_weights`` here and drive the returned network through ``apply_to``,
``prepare_optimizer_params``, ``on_epoch_start`` … ``save_weights``.  Everything in this file is
host-side plumbing; the adapters it instantiates run the sm_100a engine.

New relative to the reference: ``attach_data_parallel`` (inherited) for NCCL all-reduce of the
adapter gradients, since the reference leaves data parallelism to accelerate/DDP outside its tree.
"""

from __future__ import annotations

import logging
import os

import torch

from .logging import logger
from .modules import get_module, make_module
from .utils import precalculate_safetensors_hashes, str_bool
from .wrapper import (
    LycorisNetwork,
    _AdapterFactory,
    _algo_table,
    _assert_unique,
    _load_weights_file,
    _network_kwargs,
    _resolve_preset,
    _translate_deprecated,
    deprecated_arg_dict,  # noqa: F401  (re-exported like the reference)
    network_module_dict,  # noqa: F401
)


def _opt_float(kwargs, key):
    v = kwargs.get(key, None)
    return float(v) if v is not None else None


def create_network(multiplier, network_dim, network_alpha, vae, text_encoder, unet, **kwargs):
    _translate_deprecated(kwargs)
    if network_dim is None:
        network_dim = 4  # default
    parsed = _network_kwargs(kwargs, network_dim, network_alpha)
    parsed["rs_lora"] = str_bool(kwargs.get("rs_lora", False))
    parsed["train_t5xxl"] = str_bool(kwargs.get("train_t5xxl", False))
    lp = _opt_float(kwargs, "loraplus_lr_ratio")
    lp_unet = _opt_float(kwargs, "loraplus_unet_lr_ratio")
    lp_te = _opt_float(kwargs, "loraplus_text_encoder_lr_ratio")

    preset_str = kwargs.get("preset", "full")
    LycorisNetworkKohya.apply_preset(_resolve_preset(preset_str))

    algo = parsed["network_module"]
    logger.info(f"Using rank adaptation algo: {algo}")
    if algo == "ia3" and preset_str != "ia3":
        logger.warning("It is recommended to use preset ia3 for IA^3 algorithm")

    network = LycorisNetworkKohya(
        text_encoder, unet, multiplier=multiplier, lora_dim=network_dim, alpha=network_alpha, **parsed
    )
    if lp is not None or lp_unet is not None or lp_te is not None:
        network.set_loraplus_lr_ratio(lp, lp_unet, lp_te)
    return network


def _as_list(text_encoder):
    if isinstance(text_encoder, list):
        return text_encoder, True
    return [text_encoder], False


def create_network_from_weights(
    multiplier, file, vae, text_encoder, unet, weights_sd=None, for_inference=False, **kwargs
):
    if weights_sd is None:
        weights_sd = _load_weights_file(file)

    cls = LycorisNetworkKohya
    unet_wanted, te_wanted = {}, {}
    for key in weights_sd:
        if "." not in key:
            continue
        lora_name = key.split(".")[0]
        if lora_name.startswith(cls.LORA_PREFIX_UNET):
            unet_wanted[lora_name] = None
        elif lora_name.startswith(cls.LORA_PREFIX_TEXT_ENCODER):
            te_wanted[lora_name] = None

    for name, sub in unet.named_modules():
        lora_name = f"{cls.LORA_PREFIX_UNET}_{name}".replace(".", "_")
        if lora_name in unet_wanted:
            unet_wanted[lora_name] = sub
    if text_encoder:
        encoders, indexed = _as_list(text_encoder)
        for idx, te in enumerate(encoders):
            prefix = f"{cls.LORA_PREFIX_TEXT_ENCODER}{idx + 1}" if indexed else cls.LORA_PREFIX_TEXT_ENCODER
            for name, sub in te.named_modules():
                lora_name = f"{prefix}_{name}".replace(".", "_")
                if lora_name in te_wanted:
                    te_wanted[lora_name] = sub

    level = logger.level
    logger.setLevel(logging.ERROR)
    network = LycorisNetworkKohya(text_encoder, unet)
    network.unet_loras = []
    network.text_encoder_loras = []
    logger.setLevel(level)

    def rebuild(wanted):
        out = []
        for lora_name, target in wanted.items():
            if target is None:
                continue
            lyco_type, params = get_module(weights_sd, lora_name)
            if lyco_type is None:
                continue
            lora = make_module(lyco_type, params, lora_name, target)
            if lora is not None:
                out.append(lora)
        return out

    logger.info("Loading UNet Modules from state dict...")
    network.unet_loras = rebuild(unet_wanted)
    logger.info(f"{len(network.unet_loras)} Modules Loaded")
    logger.info("Loading TE Modules from state dict...")
    if text_encoder:
        network.text_encoder_loras = rebuild(te_wanted)
        logger.info(f"{len(network.text_encoder_loras)} Modules Loaded")

    for lora in network.unet_loras + network.text_encoder_loras:
        lora.multiplier = multiplier
    return network, weights_sd


class LycorisNetworkKohya(LycorisNetwork):
    """LoRA + LoCon (+ LoHa / LoKr / (IA)^3 / DyLoRA) network in kohya's shape."""

    ENABLE_CONV = True
    UNET_TARGET_REPLACE_MODULE = [
        "Transformer2DModel",
        "ResnetBlock2D",
        "Downsample2D",
        "Upsample2D",
        "HunYuanDiTBlock",
        "DoubleStreamBlock",
        "SingleStreamBlock",
        "SingleDiTBlock",
        "MMDoubleStreamBlock",  # HunYuanVideo
        "MMSingleStreamBlock",  # HunYuanVideo
        "WanAttentionBlock",  # Wan
        "HunyuanVideoTransformerBlock",  # FramePack
        "HunyuanVideoSingleTransformerBlock",  # FramePack
        "JointTransformerBlock",  # lumina-image-2
        "FinalLayer",  # lumina-image-2
        "QwenImageTransformerBlock",  # Qwen
    ]
    UNET_TARGET_REPLACE_NAME = ["conv_in", "conv_out", "time_embedding.linear_1", "time_embedding.linear_2"]
    TEXT_ENCODER_TARGET_REPLACE_MODULE = [
        "CLIPAttention",
        "CLIPSdpaAttention",
        "CLIPMLP",
        "MT5Block",
        "BertLayer",
        "Gemma2Attention",
        "Gemma2FlashAttention2",
        "Gemma2SdpaAttention",
        "Gemma2MLP",
    ]
    TEXT_ENCODER_TARGET_REPLACE_NAME = []
    LORA_PREFIX_UNET = "lora_unet"
    LORA_PREFIX_TEXT_ENCODER = "lora_te"
    MODULE_ALGO_MAP = {}
    NAME_ALGO_MAP = {}
    USE_FNMATCH = False

    _PRESET_ATTRS = {
        "enable_conv": "ENABLE_CONV",
        "unet_target_module": "UNET_TARGET_REPLACE_MODULE",
        "unet_target_name": "UNET_TARGET_REPLACE_NAME",
        "text_encoder_target_module": "TEXT_ENCODER_TARGET_REPLACE_MODULE",
        "text_encoder_target_name": "TEXT_ENCODER_TARGET_REPLACE_NAME",
        "module_algo_map": "MODULE_ALGO_MAP",
        "name_algo_map": "NAME_ALGO_MAP",
        "use_fnmatch": "USE_FNMATCH",
    }

    @classmethod
    def apply_preset(cls, preset):
        # unlike the generic wrapper the kohya flavour does not validate preset keys (kohya.py:286-306)
        for key, attr in cls._PRESET_ATTRS.items():
            if key in preset:
                setattr(cls, attr, preset[key])
        return cls

    def __init__(
        self,
        text_encoder,
        unet,
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
        train_t5xxl=False,
        **kwargs,
    ) -> None:
        torch.nn.Module.__init__(self)
        self.train_t5xxl = train_t5xxl
        self.loraplus_lr_ratio = None
        self.loraplus_unet_lr_ratio = None
        self.loraplus_text_encoder_lr_ratio = None
        self._common_init(multiplier, lora_dim, conv_lora_dim, alpha, conv_alpha, use_tucker, dropout,
                          rank_dropout, module_dropout)

        factory = _AdapterFactory(self, network_module, kwargs, train_norm, norm_modules,
                                  dedupe_across_roots=False, fix_root_name=False)
        cls = LycorisNetworkKohya
        self.text_encoder_loras = []
        if text_encoder:
            encoders, indexed = _as_list(text_encoder)
            for i, te in enumerate(encoders):
                self.text_encoder_loras.extend(
                    factory.build(
                        cls.LORA_PREFIX_TEXT_ENCODER + (f"{i + 1}" if indexed else ""),
                        te,
                        cls.TEXT_ENCODER_TARGET_REPLACE_MODULE,
                        cls.TEXT_ENCODER_TARGET_REPLACE_NAME,
                    )
                )
            logger.info(f"create LyCORIS for Text Encoder: {len(self.text_encoder_loras)} modules.")

        self.unet_loras = factory.build(
            cls.LORA_PREFIX_UNET, unet, cls.UNET_TARGET_REPLACE_MODULE, cls.UNET_TARGET_REPLACE_NAME
        )
        logger.info(f"create LyCORIS for U-Net: {len(self.unet_loras)} modules.")
        logger.info(f"module type table: {_algo_table(self.text_encoder_loras + self.unet_loras)}")

        self.weights_sd = None
        self.loras = self.text_encoder_loras + self.unet_loras
        _assert_unique(self.loras)

    # --------------------------------------------------------------- kohya protocol
    def apply_to(self, text_encoder, unet, apply_text_encoder=None, apply_unet=None):
        assert apply_text_encoder is not None and apply_unet is not None, "internal error: flag not set"
        if apply_text_encoder:
            logger.info("enable LyCORIS for text encoder")
        else:
            self.text_encoder_loras = []
        if apply_unet:
            logger.info("enable LyCORIS for U-Net")
        else:
            self.unet_loras = []
        self.loras = self.text_encoder_loras + self.unet_loras
        for lora in self.loras:
            lora.apply_to()
            self.add_module(lora.lora_name, lora)
        if self.weights_sd:
            info = self.load_state_dict(self.weights_sd, False)
            logger.info(f"weights are loaded: {info}")

    def merge_to(self, text_encoder, unet, weights_sd, dtype, device):
        has_te = any(k.startswith(LycorisNetworkKohya.LORA_PREFIX_TEXT_ENCODER) for k in weights_sd.keys())
        has_unet = any(k.startswith(LycorisNetworkKohya.LORA_PREFIX_UNET) for k in weights_sd.keys())
        if has_te:
            logger.info("enable LoRA for text encoder")
        else:
            self.text_encoder_loras = []
        if has_unet:
            logger.info("enable LoRA for U-Net")
        else:
            self.unet_loras = []
        self.loras = self.text_encoder_loras + self.unet_loras
        super().merge_to(1)

    def apply_max_norm_regularization(self, max_norm_value, device):
        return self._max_norm(self.unet_loras + self.text_encoder_loras, max_norm_value, device)

    def set_loraplus_lr_ratio(self, loraplus_lr_ratio, loraplus_unet_lr_ratio, loraplus_text_encoder_lr_ratio):
        self.loraplus_lr_ratio = loraplus_lr_ratio
        self.loraplus_unet_lr_ratio = loraplus_unet_lr_ratio
        self.loraplus_text_encoder_lr_ratio = loraplus_text_encoder_lr_ratio
        logger.info(f"LoRA+ UNet LR Ratio: {self.loraplus_unet_lr_ratio or self.loraplus_lr_ratio}")
        logger.info(f"LoRA+ Text Encoder LR Ratio: {self.loraplus_text_encoder_lr_ratio or self.loraplus_lr_ratio}")

    @staticmethod
    def _param_groups(loras, lr, ratio):
        """Two groups per model part: ordinary parameters, and ``lora_up`` ones at lr·ratio (LoRA+)."""
        buckets = {"lora": {}, "plus": {}}
        for lora in loras:
            for name, param in lora.named_parameters():
                which = "plus" if (ratio is not None and "lora_up" in name) else "lora"
                buckets[which][f"{lora.lora_name}.{name}"] = param
        groups, notes = [], []
        for key, named in buckets.items():
            if not named:
                continue
            group = {"params": named.values()}
            if lr is not None:
                group["lr"] = lr * ratio if key == "plus" else lr
            if not group.get("lr"):
                logger.info("NO LR skipping!")
                continue
            groups.append(group)
            notes.append("plus" if key == "plus" else "")
        return groups, notes

    def prepare_optimizer_params(self, text_encoder_lr=None, unet_lr: float = 1e-4, learning_rate=None):
        self.requires_grad_(True)
        all_params, lr_descriptions = [], []
        parts = (
            ("textencoder", self.text_encoder_loras, text_encoder_lr,
             self.loraplus_text_encoder_lr_ratio or self.loraplus_lr_ratio),
            ("unet", self.unet_loras, unet_lr, self.loraplus_unet_lr_ratio or self.loraplus_lr_ratio),
        )
        for tag, loras, lr, ratio in parts:
            if not loras:
                continue
            groups, notes = self._param_groups(loras, lr if lr is not None else learning_rate, ratio)
            all_params.extend(groups)
            lr_descriptions.extend([tag + (" " + n if n else "") for n in notes])
        return all_params, lr_descriptions

    def enable_gradient_checkpointing(self):
        pass  # not supported

    def prepare_grad_etc(self, *args):
        self.requires_grad_(True)

    def on_epoch_start(self, *args):
        self.train()

    def on_step_start(self, *args):
        pass

    def get_trainable_params(self):
        return self.parameters()

    def save_weights(self, file, dtype, metadata):
        if metadata is not None and len(metadata) == 0:
            metadata = None
        state_dict = self._export_state_dict(dtype)
        if os.path.splitext(file)[1] == ".safetensors":
            from safetensors.torch import save_file

            if metadata is None:
                metadata = {}
            metadata["sshs_model_hash"] = precalculate_safetensors_hashes(state_dict)
            save_file(state_dict, file, metadata)
        else:
            torch.save(state_dict, file)
