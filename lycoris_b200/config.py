"""Built-in presets: which module classes / names of a model receive an adapter.

Data only.  The preset names, keys and target lists are the reference's public configuration
surface (lycoris/config.py:1-196, docs/Preset.md) and have to match it for drop-in use; they are
assembled here from shared building blocks instead of nine literal dictionaries.
"""

_DIT_BLOCKS = [
    "HunYuanDiTBlock",
    "DoubleStreamBlock",  # Flux
    "SingleStreamBlock",  # Flux
    "SingleDiTBlock",  # SD3.5
    "MMDoubleStreamBlock",  # HunYuanVideo
    "MMSingleStreamBlock",  # HunYuanVideo
    "WanAttentionBlock",  # Wan
    "HunyuanVideoTransformerBlock",  # FramePack
    "HunyuanVideoSingleTransformerBlock",  # FramePack
    "JointTransformerBlock",  # lumina-image-2
    "FinalLayer",  # lumina-image-2
    "QwenImageTransformerBlock",  # Qwen
]
_UNET_TRANSFORMER = ["Transformer2DModel"] + _DIT_BLOCKS
_UNET_CONV_BLOCKS = ["ResnetBlock2D", "Downsample2D", "Upsample2D"]
_UNET_FULL = ["Transformer2DModel"] + _UNET_CONV_BLOCKS + _DIT_BLOCKS
_UNET_FULL_LIN = ["Transformer2DModel", "ResnetBlock2D"] + _DIT_BLOCKS
_UNET_IO_CONVS = ["conv_in", "conv_out"]
_UNET_TIME_EMB = ["time_embedding.linear_1", "time_embedding.linear_2"]
_GEMMA_ATTN = ["Gemma2Attention", "Gemma2FlashAttention2", "Gemma2SdpaAttention"]
_TE_FULL = ["CLIPAttention", "CLIPSdpaAttention", "CLIPMLP", "MT5Block", "BertLayer"] + _GEMMA_ATTN + ["Gemma2MLP"]
_TE_ATTN = ["CLIPAttention", "CLIPSdpaAttention", "BertAttention", "MT5LayerSelfAttention"] + _GEMMA_ATTN


def _preset(conv, unet_modules, unet_names, te_modules, te_names, **extra):
    out = {
        "enable_conv": conv,
        "unet_target_module": list(unet_modules),
        "unet_target_name": list(unet_names),
        "text_encoder_target_module": list(te_modules),
        "text_encoder_target_name": list(te_names),
    }
    out.update(extra)
    return out


PRESET = {
    "full": _preset(True, _UNET_FULL, _UNET_IO_CONVS + _UNET_TIME_EMB, _TE_FULL, []),
    "full-lin": _preset(False, _UNET_FULL_LIN, _UNET_TIME_EMB, _TE_FULL, []),
    "attn-mlp": _preset(False, _UNET_TRANSFORMER, [], _TE_FULL, []),
    "attn-only": _preset(False, ["CrossAttention", "SelfAttention"], [], _TE_ATTN, []),
    "unet-only": _preset(True, _UNET_FULL, _UNET_IO_CONVS + _UNET_TIME_EMB, [], []),
    "unet-transformer-only": _preset(False, _UNET_TRANSFORMER, [], [], []),
    "unet-convblock-only": _preset(True, _UNET_CONV_BLOCKS, _UNET_IO_CONVS, [], []),
    "ia3": _preset(
        False,
        [],
        ["to_k", "to_v", "ff.net.2"],
        [],
        ["k_proj", "v_proj", "mlp.fc2"],
        name_algo_map={"mlp.fc2": {"train_on_input": True}, "ff.net.2": {"train_on_input": True}},
    ),
}
