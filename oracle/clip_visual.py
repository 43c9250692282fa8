"""ORACLE (test infrastructure, not product code).

Restatement of the image branch the reference's style reward uses: `CLIPEncoder.get_gram_matrix_residual`
(/root/reference/text-guided-n-style/clip_guidance/base_clip.py:55-66) on top of `CLIP.encode_image_with_features`
(clip_guidance/clip/model.py:339-366) for the ViT variants (`VisionTransformer` :202-221, `ResidualAttentionBlock` :167-188,
`QuickGELU` :162-164), with the reference's parameter names so a reference `CLIP.visual.state_dict()` loads unchanged.
The GPU box has no /root/reference, hence this copy of the arithmetic; tests/test_oracle_pin.py pins it live against the reference
classes when they are importable.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x):
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution=224, patch_size=16, width=768, layers=12, heads=12, output_dim=512):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def features(self, x, upto: int = 3):
        """encode_image_with_features (model.py:339-359): token features after each of the first `upto` blocks, shape (L, N, D)."""
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        x = torch.cat([self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device), x], dim=1)
        x = self.ln_pre(x + self.positional_embedding.to(x.dtype)).permute(1, 0, 2)
        feats = []
        for blk in list(self.transformer.resblocks)[:upto]:
            x = blk(x)
            feats.append(x)
        return feats


class _ClipModelView:
    """`.clip_model.visual` access path of the reference's CLIPEncoder (base_clip.py:33)."""

    def __init__(self, visual):
        self.visual = visual


class GramStyleEncoder(nn.Module):
    """`image_encoder` of the style sampler: get_gram_matrix_residual(img in [-1,1], NCHW) -> (D,D) (base_clip.py:55-66)."""

    MEAN = (0.48145466 * 2 - 1, 0.4578275 * 2 - 1, 0.40821073 * 2 - 1)      # base_clip.py:38-41
    STD = (0.26862954 * 2, 0.26130258 * 2, 0.27577711 * 2)

    def __init__(self, visual: VisionTransformer, ref: torch.Tensor):
        super().__init__()
        self.visual = visual
        self.clip_model = _ClipModelView(visual)
        self.register_buffer("ref", ref)                                   # (1,3,224,224), already CLIP-normalised (:43-52)
        self.register_buffer("mean", torch.tensor(self.MEAN).reshape(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor(self.STD).reshape(1, 3, 1, 1))
        for p in self.parameters():
            p.requires_grad_(False)

    def get_gram_matrix_residual(self, im1):
        im1 = F.interpolate(im1, size=(224, 224), mode="bicubic")
        im1 = (im1 - self.mean.to(im1.dtype)) / self.std.to(im1.dtype)     # torchvision Normalize
        f1 = self.visual.features(im1.type(self.visual.conv1.weight.dtype))[2][1:, 0, :]
        f2 = self.visual.features(self.ref.type(self.visual.conv1.weight.dtype))[2][1:, 0, :]
        return torch.mm(f1.t(), f1) - torch.mm(f2.t(), f2)


def tiny_style_encoder(seed: int = 11, width: int = 64, layers: int = 3) -> GramStyleEncoder:
    """Seeded small ViT (patch 16, 224 px, 197 tokens like ViT-B/16) + seeded reference style image: CPU / GPU fixtures."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    vis = VisionTransformer(224, 16, width, layers, max(1, width // 64), 32)
    with torch.no_grad():
        for name, p in vis.state_dict().items():
            if p.dim() >= 2:
                fan_in = p[0].numel() if "positional" not in name and name != "proj" else p.shape[-1]
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5)
            elif "ln_" in name and name.endswith("weight"):
                p.copy_(0.8 + 0.4 * torch.rand(p.shape, generator=g))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.1)
    ref = torch.randn(1, 3, 224, 224, generator=g)
    return GramStyleEncoder(vis.eval(), ref)
