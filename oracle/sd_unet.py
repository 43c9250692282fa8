"""ORACLE (test infrastructure, not product code).

CPU fp32 restatement of the third-party denoiser the reference's hot loop calls:
`diffusers==0.18.0` `UNet2DConditionModel` in its Stable-Diffusion-1.x configuration
(pinned by /root/reference/text-guided/environment_p2p.yaml:88).  `diffusers` is not
vendored under /root/reference and is not installable here (no network), so its published
architecture is restated from the release, anchored on the call sites and attributes the
reference itself touches:

  * UNet call sites:            text-guided/inversion/p2p_h_edit.py:613,644,652
  * attention module API:       text-guided/p2p/ptp_utils.py:65-120  (spatial_norm, group_norm,
                                to_q/to_k/to_v, norm_cross, head_to_batch_dim, get_attention_scores,
                                batch_to_head_dim, to_out[0..1], residual_connection,
                                rescale_output_factor)
  * processor registry:         text-guided/p2p/ptp_utils.py:277-295 (attn_processors keys start with
                                down_blocks|mid_block|up_blocks; set_attn_processor(dict))
  * module tree / class name:   text-guided/masactrl/masactrl_utils.py:50-89 ('Attention'),
                                text-guided/plug_n_play/pnp_utils.py:13-26,88-93

Parameter names follow the diffusers state-dict so real SD-1.x weights would load unchanged.
PARITY NOTE: the reference ships no numeric tests for this boundary -> "parity unpinned" against real
diffusers; both sides of every parity test in this repo use this same restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 64
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: int = 8          # diffusers-0.18 SD-1.x quirk: this is the number of HEADS
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    # informational, read by the reference (inversion/inversion_utils.py:76)
    num_train_timesteps: int = 1000

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def tiny(sample_size: int = 16, cross_attention_dim: int = 64) -> "UNetConfig":
        """Same topology, 5x narrower: for CPU tests that must finish in seconds."""
        return UNetConfig(sample_size=sample_size, block_out_channels=(64, 128, 256, 256),
                          cross_attention_dim=cross_attention_dim)


class UNetOutput(dict):
    """Supports both `.sample` (p2p_h_edit.py:613) and `["sample"]` (ddim_inversion.py:48)."""

    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


def timestep_sinusoid(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers `get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)`."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - 0.0)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)   # flip_sin_to_cos
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, out_dim)
        self.linear_2 = nn.Linear(out_dim, out_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None
        # parameter-free attributes of diffusers' ResnetBlock2D that the Plug-and-Play patched forward reads
        # (text-guided/plug_n_play/pnp_utils.py:99-160); SD-1.x values
        self.nonlinearity = nn.SiLU()
        self.dropout = nn.Dropout(0.0)
        self.upsample = self.downsample = None
        self.time_embedding_norm = "default"
        self.output_scale_factor = 1.0

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class AttnProcessor:
    """diffusers default processor (what runs before register_attention_control replaces it)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **_):
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(ctx), attn.to_v(ctx)
        q, k, v = attn.head_to_batch_dim(q), attn.head_to_batch_dim(k), attn.head_to_batch_dim(v)
        probs = attn.get_attention_scores(q, k, attention_mask)
        out = attn.batch_to_head_dim(torch.bmm(probs, v))
        return attn.to_out[1](attn.to_out[0](out))


class Attention(nn.Module):
    """Class name must be literally 'Attention' (masactrl_utils.py:85)."""

    def __init__(self, query_dim, cross_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_dim if cross_dim is not None else query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_dim if cross_dim is not None else query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        # attributes read by P2PCrossAttnProcessor (ptp_utils.py:65-120)
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = AttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        return attention_mask

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)

    def get_attention_scores(self, query, key, attention_mask=None):
        # diffusers: baddbmm(empty, q, k^T, beta=0, alpha=scale) -> softmax(dim=-1)
        scores = torch.baddbmm(
            torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device),
            query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        return scores.softmax(dim=-1)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.proj = nn.Linear(din, dout * 2)

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
    def __init__(self, dim, heads, dim_head, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx, cross_attention_kwargs):
        x = self.attn1(self.norm1(x), **cross_attention_kwargs) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=ctx, **cross_attention_kwargs) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, channels, cross_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, dim_head, cross_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x, ctx, cross_attention_kwargs):
        b, c, h, w = x.shape
        res = x
        y = self.proj_in(self.norm(x))
        y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            y = blk(y, ctx, cross_attention_kwargs)
        y = y.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(y) + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, temb_dim, has_attn, add_down):
        super().__init__()
        heads = cfg.attention_head_dim
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for i in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(heads, cout // heads, cout, cfg.cross_attention_dim, cfg.norm_num_groups))
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx, kw):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx, kw)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, c, temb_dim):
        super().__init__()
        heads = cfg.attention_head_dim
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, cfg.cross_attention_dim, cfg.norm_num_groups)])

    def forward(self, x, temb, ctx, kw):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx, kw)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg, cin, cout, cprev, temb_dim, has_attn, add_up):
        super().__init__()
        heads = cfg.attention_head_dim
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if has_attn else None
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            self.resnets.append(ResnetBlock2D(rin + skip, cout, temb_dim, cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(heads, cout // heads, cout, cfg.cross_attention_dim, cfg.norm_num_groups))
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx, kw):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx, kw)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class _Cfg:
    def __init__(self, d):
        self.__dict__.update(d)


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        self.config = _Cfg(dict(in_channels=cfg.in_channels, sample_size=cfg.sample_size,
                                block_out_channels=cfg.block_out_channels))
        self.in_channels = cfg.in_channels     # ddpm_inversion.py:29
        self.sample_size = cfg.sample_size     # ddpm_inversion.py:30
        boc = cfg.block_out_channels
        temb_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            last = i == len(boc) - 1
            self.down_blocks.append(DownBlock(cfg, cin, cout, temb_dim, has_attn=not last, add_down=not last))
        self.mid_block = MidBlock(cfg, boc[-1], temb_dim)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        cout = rev[0]
        for i, c in enumerate(rev):
            cprev, cout = cout, c
            cin = rev[min(i + 1, len(boc) - 1)]
            last = i == len(boc) - 1
            self.up_blocks.append(UpBlock(cfg, cin, cout, cprev, temb_dim, has_attn=i != 0, add_up=not last))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    # ---- processor registry (ptp_utils.py:277-295) ----
    def _attn_modules(self):
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                yield name, m

    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": m.processor for name, m in self._attn_modules()}

    def set_attn_processor(self, processors):
        for name, m in self._attn_modules():
            if isinstance(processors, dict):
                m.set_processor(processors[f"{name}.processor"])
            else:
                m.set_processor(processors)

    def forward(self, sample, timestep, encoder_hidden_states=None, cross_attention_kwargs=None, **_):
        kw = dict(cross_attention_kwargs or {})
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        timestep = timestep.reshape(-1).to(sample.device).float().expand(sample.shape[0])
        temb = self.time_embedding(timestep_sinusoid(timestep, self.cfg.block_out_channels[0]))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states, kw)
            skips += outs
        x = self.mid_block(x, temb, encoder_hidden_states, kw)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states, kw)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return UNetOutput(x)


def seeded_init_(model: nn.Module, seed: int = 0, gain: float = 1.0) -> nn.Module:
    """Deterministic, platform-independent weight init (CPU generator; default-PyTorch-like scales).

    weights ~ U(-b, b), b = gain*sqrt(3/fan_in)  (unit-variance-preserving; the PyTorch default's 1/sqrt(3) shrink
    would make a 60-layer random net numerically dead), biases ~ U(-0.1, 0.1), norm gamma ~ U(0.8, 1.2).
    Parameters are visited in state_dict order so the stream is reproducible anywhere.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in model.state_dict().items():
            if not torch.is_floating_point(p):
                continue
            if p.dim() >= 2:
                fan_in = p[0].numel()
                b = gain * math.sqrt(3.0 / fan_in)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * b)
            elif name.endswith("weight"):      # norm gamma
                p.copy_(0.8 + 0.4 * torch.rand(p.shape, generator=g))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.1)
    return model
