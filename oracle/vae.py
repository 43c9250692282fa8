"""ORACLE (test infrastructure, not product code).

CPU fp32 restatement of the VAE DECODER the reference's style path differentiates through
(`model.vae.decode(1 / 0.18215 * x0).sample`, /root/reference/text-guided-n-style/inversion/h_edit.py:158-159; also the final
decode at text-guided/main_p2p.py:262-275): `diffusers==0.18.0` `AutoencoderKL` in its SD-1.x configuration
(block_out_channels (128,256,512,512), layers_per_block 2, 32 groups, eps 1e-6, one single-head attention in the mid block,
nearest-2x + conv upsamplers, post_quant_conv 1x1).  `diffusers` is not vendored under /root/reference and not installable
offline, so the published architecture is restated with diffusers' parameter names.  PARITY NOTE: the reference has no tests at
this boundary -> "parity unpinned" against real diffusers; both sides of every parity test use this restatement.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sd_unet import seeded_init_


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32

    @staticmethod
    def tiny() -> "VAEConfig":
        """Same topology (3 upsamplers = 8x), narrow (channel counts stay multiples of 64, the native conv's TMA box): fixtures."""
        return VAEConfig(block_out_channels=(64, 64, 128, 128))


class _Sample:
    def __init__(self, sample):
        self.sample = sample


class VaeResnet(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class VaeAttention(nn.Module):
    """diffusers-0.18 `Attention` as built by UNetMidBlock2D for the VAE: one head of dim C, GroupNorm, biased projections,
    residual connection."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x.reshape(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        p = torch.softmax(torch.baddbmm(torch.empty(b, h * w, h * w, dtype=q.dtype, device=q.device), q, k.transpose(1, 2), beta=0,
                                        alpha=c ** -0.5), dim=-1)
        o = self.to_out[0](torch.bmm(p, v)).transpose(1, 2).reshape(b, c, h, w)
        return o + x


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(c, c, groups), VaeResnet(c, c, groups)])
        self.attentions = nn.ModuleList([VaeAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Upsampler(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(cin if i == 0 else cout, cout, groups) for i in range(n)])
        self.upsamplers = nn.ModuleList([_Upsampler(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.upsamplers is None else self.upsamplers[0](x)


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = _Mid(boc[-1], g)
        rev = list(reversed(boc))
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, c in enumerate(rev):
            self.up_blocks.append(_UpBlock(prev, c, cfg.layers_per_block + 1, g, add_up=i != len(rev) - 1))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class _Downsampler(nn.Module):
    """diffusers Downsample2D(padding=0): zero-pad (0,1,0,1) then a stride-2 3x3 conv."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(cin if i == 0 else cout, cout, groups) for i in range(n)])
        self.downsamplers = nn.ModuleList([_Downsampler(cout)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.downsamplers is None else self.downsamplers[0](x)


class Encoder(nn.Module):
    """diffusers-0.18 `Encoder` (AutoencoderKL encode direction, SD-1.x): conv_in, 4 DownEncoderBlock2D, the same mid block as the
    decoder, GroupNorm + SiLU + conv_out to 2 x latent channels (mean | logvar)."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.out_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        prev = boc[0]
        for i, c in enumerate(boc):
            self.down_blocks.append(_DownBlock(prev, c, cfg.layers_per_block, g, add_down=i != len(boc) - 1))
            prev = c
        self.mid_block = _Mid(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class _Gaussian:
    def __init__(self, moments):
        self.mean, self.logvar = moments.chunk(2, dim=1)

    def mode(self):
        return self.mean


class _EncOut:
    def __init__(self, moments):
        self.latent_dist = _Gaussian(moments)


class AutoencoderKLEncoder(nn.Module):
    """`model.vae` for the encode direction: `.encode(x).latent_dist.mode()` / `.mean` (text-guided/main_p2p.py:154-159)."""

    def __init__(self, cfg: VAEConfig = VAEConfig(), seed: int = 8):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        seeded_init_(self, seed)
        for p in self.parameters():
            p.requires_grad_(False)

    def encode(self, x):
        return _EncOut(self.quant_conv(self.encoder(x)))


class AutoencoderKLDecoder(nn.Module):
    """`model.vae` for the decode direction: `.decode(z).sample`."""

    def __init__(self, cfg: VAEConfig = VAEConfig(), seed: int = 7):
        super().__init__()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)
        seeded_init_(self, seed)
        for p in self.parameters():
            p.requires_grad_(False)

    def decode(self, z):
        return _Sample(self.decoder(self.post_quant_conv(z)))
