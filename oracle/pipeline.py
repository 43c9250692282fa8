"""ORACLE (test infrastructure, not product code).

Duck-typed stand-in for the `StableDiffusionPipeline` object the reference samplers receive as `model`
(attributes read by the hot path are listed in SURVEY.md section 8b):

  model.unet, model.scheduler.{alphas_cumprod, final_alpha_cumprod, timesteps, num_inference_steps, alphas,
  config.num_train_timesteps}, model.tokenizer, model.text_encoder, model.device

Restates `diffusers==0.18.0` `DDIMScheduler` tables as configured at
/root/reference/text-guided/main_p2p.py:139-146 (scaled_linear betas 0.00085..0.012, 1000 train steps,
set_alpha_to_one=False, timestep_spacing="leading", steps_offset=1 for the hub config used when eta>0).

The tokenizer / text encoder are L0 third-party components outside the hot path (CLIP ViT-L/14 text tower,
no weights or vocabulary available offline); seeded stand-ins with the same call signatures are used so that the
controller set-up code (word -> token indices, sequence aligner) and `encode_text` (inversion_utils.py:13) run
unchanged.
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence, Union

import torch
import torch.nn as nn

from .sd_unet import UNet2DConditionModel, UNetConfig, seeded_init_


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class DDIMSchedulerTables:
    """Only the tables and `set_timesteps` the reference reads."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1,
                 set_alpha_to_one=False):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                           timestep_spacing="leading")
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.config.num_train_timesteps // n
        ts = (torch.arange(0, n) * ratio).round().flip(0).to(torch.int64) + self.config.steps_offset
        self.timesteps = ts


class ToyTokenizer:
    """Word-level tokenizer with the HF-CLIP call surface used by the reference:
    `tokenizer(prompts, padding=, max_length=, truncation=, return_tensors=).input_ids`,
    `.encode(text)` (BOS + words + EOS), `.decode([id])`, `.model_max_length` (inversion_utils.py:25-31,
    ptp_utils.py:305, seq_aligner.py:112-113)."""

    model_max_length = 77
    bos, eos = 49406, 49407

    def __init__(self):
        self._words = {}

    def _id(self, w: str) -> int:
        i = int(hashlib.md5(w.encode()).hexdigest(), 16) % 49000 + 1
        self._words.setdefault(i, w)
        return i

    def encode(self, text: str) -> List[int]:
        return [self.bos] + [self._id(w) for w in text.split(" ") if w != ""] + [self.eos]

    def decode(self, ids: Sequence[int]) -> str:
        out = []
        for i in ids:
            i = int(i)
            out.append("<|startoftext|>" if i == self.bos else "<|endoftext|>" if i == self.eos else self._words.get(i, "?"))
        return " ".join(out)

    def __call__(self, prompts: Union[str, List[str]], padding="max_length", max_length=77, truncation=True,
                 return_tensors="pt"):
        if isinstance(prompts, str):
            prompts = [prompts]
        ids = torch.full((len(prompts), max_length), self.eos, dtype=torch.int64)
        for r, p in enumerate(prompts):
            e = self.encode(p)[:max_length]
            e[-1] = self.eos
            ids[r, : len(e)] = torch.tensor(e)
        return _Cfg(input_ids=ids)


class ToyTextEncoder(nn.Module):
    """Seeded stand-in for CLIPTextModel: ids (B,77) -> [(B,77,D)].  Token + position embedding and one
    causal mixing layer so that every position depends on the prompt prefix like CLIP's text tower."""

    def __init__(self, dim=768, vocab=49408, seed=1234):
        super().__init__()
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.dim = dim
        self.register_buffer("pos", torch.randn(77, dim, generator=g) * 0.5)
        self.register_buffer("proj", torch.randn(dim, dim, generator=g) / dim ** 0.5)
        self.seed = seed
        self.vocab = vocab

    def _tok(self, ids: torch.Tensor) -> torch.Tensor:
        # hash-seeded per-token vectors (avoids a 49408 x dim table)
        flat = ids.reshape(-1).tolist()
        out = torch.empty(len(flat), self.dim)
        cache = {}
        for n, i in enumerate(flat):
            if i not in cache:
                g = torch.Generator(device="cpu").manual_seed(self.seed * 1000003 + int(i))
                cache[i] = torch.randn(self.dim, generator=g)
            out[n] = cache[i]
        return out.reshape(*ids.shape, self.dim)

    @torch.no_grad()
    def forward(self, ids):
        x = self._tok(ids.cpu()) + self.pos
        mix = torch.cumsum(x, dim=1) / torch.arange(1, 78, dtype=torch.float32)[None, :, None]
        y = torch.tanh((x + mix) @ self.proj) * 1.5
        return (y.to(ids.device if ids.device.type != "meta" else "cpu"),)


class OraclePipeline:
    def __init__(self, cfg: UNetConfig = UNetConfig(), seed: int = 0, steps_offset: int = 1, build_unet: bool = True):
        self.device = torch.device("cpu")
        self.unet = seeded_init_(UNet2DConditionModel(cfg), seed).eval() if build_unet else None
        if self.unet is not None:
            for p in self.unet.parameters():
                p.requires_grad_(False)
        self.scheduler = DDIMSchedulerTables(steps_offset=steps_offset)
        self.tokenizer = ToyTokenizer()
        self.text_encoder = ToyTextEncoder(dim=cfg.cross_attention_dim)
        self.vae = None
