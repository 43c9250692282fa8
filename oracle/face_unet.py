"""ORACLE (test infrastructure, not product code).

CPU fp32 restatement of the pixel-space DDPM denoiser of the face-swapping path
(/root/reference/face-swapping/diffusion/diffusion.py:193-341 `Model`, with ResnetBlock :74-137, AttnBlock :140-189, Downsample :56-71
(zero padding (0,1,0,1) then a stride-2 3x3 conv), Upsample :37-53, get_timestep_embedding :6-24) in the CelebA-HQ configuration the
driver builds (face-swapping/main_edit.py:84-100: ch 128, ch_mult (1,1,2,2,4,4), 2 res blocks, attention at 16x16, 256x256 images),
keeping the reference's parameter names so a reference state dict loads unchanged.  The GPU box has no /root/reference, hence this
copy of the arithmetic; tests/test_oracle_pin.py pins it live against the reference class when that is importable.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sd_unet import seeded_init_


@dataclass
class FaceUNetConfig:
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 4, 4)
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (16,)
    in_channels: int = 3
    out_ch: int = 3
    image_size: int = 256

    @staticmethod
    def tiny() -> "FaceUNetConfig":
        return FaceUNetConfig(ch=64, ch_mult=(1, 2, 2), attn_resolutions=(16,), image_size=64)

    def as_reference_dict(self) -> dict:
        return dict(type="simple", in_channels=self.in_channels, out_ch=self.out_ch, ch=self.ch, ch_mult=list(self.ch_mult),
                    num_res_blocks=self.num_res_blocks, attn_resolutions=list(self.attn_resolutions), dropout=0.0, var_type="fixedsmall",
                    ema_rate=0.999, ema=True, resamp_with_conv=True, image_size=self.image_size, num_diffusion_timesteps=1000)


def sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusion.py:6-24: [sin | cos], frequencies exp(-log(1e4) k / (dim/2 - 1))."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(10000) / (half - 1)))
    ang = t.float()[:, None] * freq[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


def _gn(c):
    return nn.GroupNorm(32, c, eps=1e-6)


class _Res(nn.Module):
    def __init__(self, cin, cout, temb):
        super().__init__()
        self.norm1, self.conv1 = _gn(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.temb_proj = nn.Linear(temb, cout)
        self.norm2, self.conv2 = _gn(cout), nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)
        self.has_sc = cin != cout

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x))) + self.temb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.nin_shortcut(x) if self.has_sc else x) + h


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _gn(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.norm(x)
        q = self.q(y).reshape(b, c, h * w).transpose(1, 2)
        k = self.k(y).reshape(b, c, h * w)
        v = self.v(y).reshape(b, c, h * w)
        p = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
        o = torch.bmm(v, p.transpose(1, 2)).reshape(b, c, h, w)
        return x + self.proj_out(o)


class _Conv(nn.Module):
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=0 if stride == 2 else 1)
        self.stride = stride


class _Level(nn.Module):
    pass


class FaceUNet(nn.Module):
    def __init__(self, cfg: FaceUNetConfig = FaceUNetConfig(), seed: int = 3):
        super().__init__()
        self.cfg = cfg
        self.in_channels, self.resolution = cfg.in_channels, cfg.image_size          # read by sde_inversion.py:32-33
        ch, mult, nres = cfg.ch, tuple(cfg.ch_mult), cfg.num_res_blocks
        tch = 4 * ch
        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, tch), nn.Linear(tch, tch)])
        self.conv_in = nn.Conv2d(cfg.in_channels, ch, 3, padding=1)
        in_mult = (1,) + mult
        res = cfg.image_size
        self.down = nn.ModuleList()
        cur = ch
        for i in range(len(mult)):
            lv = _Level()
            lv.block, lv.attn = nn.ModuleList(), nn.ModuleList()
            cur = ch * in_mult[i]
            for _ in range(nres):
                lv.block.append(_Res(cur, ch * mult[i], tch))
                cur = ch * mult[i]
                if res in cfg.attn_resolutions:
                    lv.attn.append(_Attn(cur))
            if i != len(mult) - 1:
                lv.downsample = _Conv(cur, 2)
                res //= 2
            self.down.append(lv)
        self.mid = nn.Module()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = _Res(cur, cur, tch), _Attn(cur), _Res(cur, cur, tch)
        ups = []
        for i in reversed(range(len(mult))):
            lv = _Level()
            lv.block, lv.attn = nn.ModuleList(), nn.ModuleList()
            skip = ch * mult[i]
            for j in range(nres + 1):
                if j == nres:
                    skip = ch * in_mult[i]
                lv.block.append(_Res(cur + skip, ch * mult[i], tch))
                cur = ch * mult[i]
                if res in cfg.attn_resolutions:
                    lv.attn.append(_Attn(cur))
            if i != 0:
                lv.upsample = _Conv(cur, 1)
                res *= 2
            ups.insert(0, lv)
        self.up = nn.ModuleList(ups)
        self.norm_out = _gn(cur)
        self.conv_out = nn.Conv2d(cur, cfg.out_ch, 3, padding=1)
        seeded_init_(self, seed)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x, t):
        cfg = self.cfg
        temb = self.temb.dense[1](F.silu(self.temb.dense[0](sinusoid(t, cfg.ch))))
        hs = [self.conv_in(x)]
        n = len(cfg.ch_mult)
        for i, lv in enumerate(self.down):
            for j in range(cfg.num_res_blocks):
                h = lv.block[j](hs[-1], temb)
                if len(lv.attn) > 0:
                    h = lv.attn[j](h)
                hs.append(h)
            if i != n - 1:
                hs.append(lv.downsample.conv(F.pad(hs[-1], (0, 1, 0, 1))))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(hs[-1], temb)), temb)
        for i in reversed(range(n)):
            lv = self.up[i]
            for j in range(cfg.num_res_blocks + 1):
                h = lv.block[j](torch.cat([h, hs.pop()], dim=1), temb)
                if len(lv.attn) > 0:
                    h = lv.attn[j](h)
            if i != 0:
                h = lv.upsample.conv(F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return self.conv_out(F.silu(self.norm_out(h)))


# ---- seeded stand-ins for the two reward models of the face path (protocol: SURVEY 8b) -------------------------------------------
class TinyIDLoss(nn.Module):
    """Stand-in for arcface_model.IDLoss (IR-SE50 weights are not available offline): `get_cosine_loss(x0) -> scalar`
    = mean over the batch of 1 - cos(feat(x0), feat(ref)) (arcface_model.py:48-67) on a small seeded conv encoder."""

    def __init__(self, ref: torch.Tensor, seed: int = 21):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.PReLU(16), nn.Conv2d(16, 32, 3, stride=2, padding=1), nn.PReLU(32),
                                 nn.AdaptiveAvgPool2d(4), nn.Flatten(), nn.Linear(512, 64))
        seeded_init_(self.net, seed)
        self.register_buffer("ref", ref)
        for p in self.parameters():
            p.requires_grad_(False)

    def get_cosine_sim(self, image):
        a, b = F.normalize(self.net(image), p=2, dim=-1), F.normalize(self.net(self.ref), p=2, dim=-1)
        return F.cosine_similarity(b, a, dim=-1)

    def get_cosine_loss(self, image):
        return (1 - self.get_cosine_sim(image)).mean()


class TinyLPIPSLoss(nn.Module):
    """Stand-in for arcface_model.LPIPS_Loss (VGG-LPIPS weights are not available offline): `get_lpips_loss(x) -> scalar`
    = batch mean of a channel-normalised feature distance to the source image (arcface_model.py:91-95)."""

    def __init__(self, src: torch.Tensor, seed: int = 22):
        super().__init__()
        self.f1 = nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU())
        self.f2 = nn.Sequential(nn.MaxPool2d(2), nn.Conv2d(16, 32, 3, padding=1), nn.ReLU())
        seeded_init_(self, seed)
        self.register_buffer("src", src)
        for p in self.parameters():
            p.requires_grad_(False)

    def _feats(self, x):
        a = self.f1(x)
        return [a, self.f2(a)]

    def get_lpips_loss(self, x):
        d = 0
        for fx, fs in zip(self._feats(x), self._feats(self.src)):
            fx = fx / (fx.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            fs = fs / (fs.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            d = d + (fx - fs).pow(2).sum(1).mean(dim=(1, 2))
        return d.mean()
