"""SDXL- / SD1.5-shaped UNet skeleton (random init, synthetic inputs).

Class names (``Transformer2DModel``, ``ResnetBlock2D``, ``Downsample2D``, ``Upsample2D``,
``Attention``, ``FeedForward``) and attribute paths (``down_blocks.1.attentions.0.transformer_blocks
.0.attn1.to_q`` …) follow diffusers' ``UNet2DConditionModel`` so the reference presets
(lycoris/config.py) and kohya's ``lora_unet_*`` names resolve as on the real checkpoints.
Shapes come from the public SDXL-base / SD1.5 configs.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers: tuple = (0, 2, 10)  # per down block; 0 = no attention in that block
    mid_transformer_layers: int = 10
    cross_attention_dim: int = 2048
    head_dim: int = 64
    num_heads: int = 0  # >0: fixed head count (SD1.5 uses 8), else channels // head_dim
    linear_projection: bool = True  # SDXL: proj_in/out are Linear; SD1.5: 1x1 Conv2d
    addition_embed_dim: int = 2816  # SDXL added-cond (text_embeds 1280 + 6 time ids x 256); 0 = none
    norm_groups: int = 32
    sample_size: int = 128
    context_len: int = 77
    name: str = "sdxl"


SDXL = UNetConfig()
SD15 = UNetConfig(
    block_out_channels=(320, 640, 1280, 1280), transformer_layers=(1, 1, 1, 0), mid_transformer_layers=1,
    cross_attention_dim=768, num_heads=8, linear_projection=False, addition_embed_dim=0, sample_size=64,
    name="sd15",
)
TOY = UNetConfig(
    block_out_channels=(32, 64), transformer_layers=(0, 1), mid_transformer_layers=1, cross_attention_dim=48,
    head_dim=16, addition_embed_dim=0, norm_groups=8, sample_size=16, context_len=7, layers_per_block=1,
    name="toy",
)


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class Attention(nn.Module):
    def __init__(self, dim, heads, context_dim=None):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(context_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(context_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, t, d = x.shape
        q = self.to_q(x).view(b, t, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.to_out[0](o.transpose(1, 2).reshape(b, t, d))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, context_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, context_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, depth, context_dim, linear_projection, groups):
        super().__init__()
        self.linear_projection = linear_projection
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        if linear_projection:
            self.proj_in = nn.Linear(channels, channels)
            self.proj_out = nn.Linear(channels, channels)
        else:
            self.proj_in = nn.Conv2d(channels, channels, 1)
            self.proj_out = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, heads, context_dim) for _ in range(depth)]
        )

    def forward(self, x, context):
        b, c, h, w = x.shape
        res = x
        x = self.norm(x)
        if self.linear_projection:
            x = self.proj_in(x.permute(0, 2, 3, 1).reshape(b, h * w, c))
        else:
            x = self.proj_in(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        if self.linear_projection:
            x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2)
        else:
            x = self.proj_out(x.reshape(b, h, w, c).permute(0, 3, 1, 2))
        return x + res


class ResnetBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, temb_ch, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_ch, eps=1e-5)
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, out_ch)
        self.norm2 = nn.GroupNorm(groups, out_ch, eps=1e-5)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList()
        self.downsamplers = None
        self.upsamplers = None


class UNetSkeleton(nn.Module):
    def __init__(self, cfg: UNetConfig = SDXL):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb = ch[0] * 4
        g = cfg.norm_groups

        def heads(c):
            return cfg.num_heads or c // cfg.head_dim

        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.add_embedding = TimestepEmbedding(cfg.addition_embed_dim, temb) if cfg.addition_embed_dim else None

        self.down_blocks = nn.ModuleList()
        skips = [ch[0]]
        cur = ch[0]
        for i, out in enumerate(ch):
            blk = _Block()
            for _ in range(cfg.layers_per_block):
                blk.resnets.append(ResnetBlock2D(cur, out, temb, g))
                cur = out
                if cfg.transformer_layers[i]:
                    blk.attentions.append(Transformer2DModel(out, heads(out), cfg.transformer_layers[i],
                                                             cfg.cross_attention_dim, cfg.linear_projection, g))
                skips.append(cur)
            if i != len(ch) - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(cur)])
                skips.append(cur)
            self.down_blocks.append(blk)

        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(cur, cur, temb, g))
        self.mid_block.attentions.append(Transformer2DModel(cur, heads(cur), cfg.mid_transformer_layers,
                                                            cfg.cross_attention_dim, cfg.linear_projection, g))
        self.mid_block.resnets.append(ResnetBlock2D(cur, cur, temb, g))

        self.up_blocks = nn.ModuleList()
        for i, out in reversed(list(enumerate(ch))):
            blk = _Block()
            for _ in range(cfg.layers_per_block + 1):
                skip = skips.pop()
                blk.resnets.append(ResnetBlock2D(cur + skip, out, temb, g))
                cur = out
                if cfg.transformer_layers[i]:
                    blk.attentions.append(Transformer2DModel(out, heads(out), cfg.transformer_layers[i],
                                                             cfg.cross_attention_dim, cfg.linear_projection, g))
            if i != 0:
                blk.upsamplers = nn.ModuleList([Upsample2D(cur)])
            self.up_blocks.append(blk)

        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.in_channels, 3, padding=1)

    def forward(self, sample, timesteps, context, added_cond=None):
        temb = self.time_embedding(timestep_embedding(timesteps, self.cfg.block_out_channels[0]).to(sample.dtype))
        if self.add_embedding is not None:
            temb = temb + self.add_embedding(added_cond.to(sample.dtype))
        x = self.conv_in(sample)
        stack = [x]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                x = res(x, temb)
                if len(blk.attentions):
                    x = blk.attentions[j](x, context)
                stack.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0](x)
                stack.append(x)
        x = self.mid_block.resnets[0](x, temb)
        x = self.mid_block.attentions[0](x, context)
        x = self.mid_block.resnets[1](x, temb)
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                x = res(torch.cat([x, stack.pop()], dim=1), temb)
                if len(blk.attentions):
                    x = blk.attentions[j](x, context)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0](x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))

    def synthetic_batch(self, batch, device, dtype, seed=2, sample_size=None):
        """Seeded N(0,1) latents / context / target, uniform timesteps."""
        cfg = self.cfg
        s = sample_size or cfg.sample_size
        g = torch.Generator().manual_seed(seed)
        out = {
            "sample": torch.randn(batch, cfg.in_channels, s, s, generator=g),
            "timesteps": torch.randint(0, 1000, (batch,), generator=g),
            "context": torch.randn(batch, cfg.context_len, cfg.cross_attention_dim, generator=g),
            "target": torch.randn(batch, cfg.in_channels, s, s, generator=g),
        }
        if cfg.addition_embed_dim:
            out["added_cond"] = torch.randn(batch, cfg.addition_embed_dim, generator=g)
        return out


def wrapped_layer_flops(model, batch, sample_size=None):
    """One dense pass (2*M*N*K*kh*kw) over every Linear/Conv2d inside the preset-"full" target
    classes, measured by running a meta-device forward with hooks.  Returns (flops, n_layers)."""
    targets = ("Transformer2DModel", "ResnetBlock2D", "Downsample2D", "Upsample2D")
    layers = {}
    for _, mod in model.named_modules():
        if mod.__class__.__name__ in targets:
            for _, sub in mod.named_modules():
                if isinstance(sub, (nn.Linear, nn.Conv2d)):
                    layers[id(sub)] = sub
    total = [0]

    def hook(m, inp, out):
        if isinstance(m, nn.Linear):
            total[0] += 2 * out.numel() * m.in_features
        else:
            total[0] += 2 * out.numel() * m.in_channels * m.kernel_size[0] * m.kernel_size[1]

    hs = [m.register_forward_hook(hook) for m in layers.values()]
    try:
        cfg = model.cfg
        s = sample_size or cfg.sample_size
        dev = next(model.parameters()).device
        dt = next(model.parameters()).dtype
        with torch.no_grad():
            b = {
                "sample": torch.zeros(batch, cfg.in_channels, s, s, device=dev, dtype=dt),
                "timesteps": torch.zeros(batch, device=dev, dtype=torch.long),
                "context": torch.zeros(batch, cfg.context_len, cfg.cross_attention_dim, device=dev, dtype=dt),
            }
            if cfg.addition_embed_dim:
                b["added_cond"] = torch.zeros(batch, cfg.addition_embed_dim, device=dev, dtype=dt)
            model(b["sample"], b["timesteps"], b["context"], b.get("added_cond"))
    finally:
        for h in hs:
            h.remove()
    return total[0], len(layers)
