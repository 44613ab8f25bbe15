"""Small models for the CPU test-suite and the golden-fixture generators."""

import torch
import torch.nn as nn
import torch.nn.functional as F

from workloads.unet_skeleton import TOY, UNetSkeleton


def ToyUNet():
    return UNetSkeleton(TOY)


# A CLIP-shaped text encoder: the class NAMES are what the adapter discovery matches on
# (TEXT_ENCODER_TARGET_REPLACE_MODULE, lycoris/kohya.py:40-46 / config.py), the paths are transformers' own
# (text_model.encoder.layers.N.self_attn.q_proj ...), so the adapter names come out as kohya checkpoints have them.
class CLIPAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.q_proj = nn.Linear(dim, dim)
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, L, C = x.shape
        q, k, v = (p(x).view(B, L, self.heads, C // self.heads).transpose(1, 2)
                   for p in (self.q_proj, self.k_proj, self.v_proj))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.out_proj(o.transpose(1, 2).reshape(B, L, C))


class CLIPMLP(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class CLIPEncoderLayer(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.self_attn = CLIPAttention(dim, heads)
        self.layer_norm1 = nn.LayerNorm(dim)
        self.mlp = CLIPMLP(dim, 4 * dim)
        self.layer_norm2 = nn.LayerNorm(dim)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Encoder(nn.Module):
    def __init__(self, dim, heads, layers):
        super().__init__()
        self.layers = nn.ModuleList(CLIPEncoderLayer(dim, heads) for _ in range(layers))


class _TextModel(nn.Module):
    def __init__(self, vocab, dim, heads, layers, max_len):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, dim)
        self.position_embedding = nn.Embedding(max_len, dim)
        self.encoder = _Encoder(dim, heads, layers)
        self.final_layer_norm = nn.LayerNorm(dim)


class ToyTextEncoder(nn.Module):
    def __init__(self, dim=32, heads=2, layers=2, vocab=64, max_len=16):
        super().__init__()
        self.text_model = _TextModel(vocab, dim, heads, layers, max_len)

    def forward(self, ids):
        tm = self.text_model
        x = tm.token_embedding(ids) + tm.position_embedding(torch.arange(ids.shape[1], device=ids.device))[None]
        for layer in tm.encoder.layers:
            x = layer(x)
        return tm.final_layer_norm(x)
