"""A reference-shaped pipeline object assembled from native engines with synthetic weights.

The reference's samplers take `model` = a diffusers StableDiffusionPipeline (duck-typed: `unet`, `scheduler`, `tokenizer`,
`text_encoder`, `device`; text-guided/main_p2p.py:98-146).  No pretrained weights, diffusers or CLIP vocabulary exist offline, so
benchmarks and demos build this stand-in: SD-1.5 UNet geometry and the CLIP ViT-L/14 text-tower geometry with seeded random weights
generated on the device, the word-level tokenizer and the DDIM tables.  Everything the public samplers do with a real pipeline (tokenise
-> text tower -> controller set-up -> edit plan -> native loop) runs unchanged on it."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import UNetEngine
from .schedule import DDIMTables
from .text_encoder import TextEncoderEngine
from .tokenizer import WordTokenizer

SD15_UNET = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
                 cross_attention_dim=768, norm_groups=32, ctx_len=77)
CLIP_L14_TEXT = dict(vocab=49408, width=768, heads=12, layers=12, ffn=3072, tokens=77)


def random_text_engine(cfg=None, seed: int = 1, device: int = 0) -> TextEncoderEngine:
    """CLIP text tower (transformers CLIPTextModel parameter names) with seeded random weights generated on the device."""
    cfg = dict(CLIP_L14_TEXT if cfg is None else cfg)
    eng = TextEncoderEngine(cfg["vocab"], cfg["width"], cfg["heads"], cfg["layers"], cfg["ffn"], cfg["tokens"], device)
    W, F = cfg["width"], cfg["ffn"]
    g = torch.Generator(device=f"cuda:{device}").manual_seed(seed)
    dev = torch.device("cuda", device)

    def put(name, shape, kind):
        if kind == "w":
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * (3.0 / shape[-1]) ** 0.5
        elif kind == "e":
            t = torch.randn(shape, generator=g, device=dev) * 0.02
        elif kind == "g":
            t = 0.9 + 0.2 * torch.rand(shape, generator=g, device=dev)
        else:
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * 0.05
        dims = (C.c_int64 * len(shape))(*shape)
        _lib.check(eng.lib.hedit_text_load_tensor(eng.handle, name.encode(), t.data_ptr(), dims, len(shape)), f"load {name}")

    put("text_model.embeddings.token_embedding.weight", (cfg["vocab"], W), "e")
    put("text_model.embeddings.position_embedding.weight", (cfg["tokens"], W), "e")
    for i in range(cfg["layers"]):
        p = f"text_model.encoder.layers.{i}"
        for n in ("layer_norm1", "layer_norm2"):
            put(f"{p}.{n}.weight", (W,), "g"); put(f"{p}.{n}.bias", (W,), "b")
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            put(f"{p}.self_attn.{n}.weight", (W, W), "w"); put(f"{p}.self_attn.{n}.bias", (W,), "b")
        put(f"{p}.mlp.fc1.weight", (F, W), "w"); put(f"{p}.mlp.fc1.bias", (F,), "b")
        put(f"{p}.mlp.fc2.weight", (W, F), "w"); put(f"{p}.mlp.fc2.bias", (W,), "b")
    put("text_model.final_layer_norm.weight", (W,), "g"); put("text_model.final_layer_norm.bias", (W,), "b")
    _lib.check(eng.lib.hedit_text_finalize(eng.handle), "finalize text encoder weights")
    return eng


class SyntheticPipeline:
    """`model` for the h_Edit_* callables: .unet (geometry only), .scheduler, .tokenizer, .text_encoder, .device, with the native engines
    pre-attached where `get_engine` / `get_text_engine` look for them."""

    class _UNetStub:
        def __init__(self, cfg):
            self.cfg = dict(cfg)

        def state_dict(self):
            raise RuntimeError("SyntheticPipeline carries device-generated random weights only; size it with max_batch at construction")

    def __init__(self, max_batch: int = 8, num_inference_steps: int = 50, steps_offset: int = 1, unet_cfg=None, text_cfg=None, seed: int = 0,
                 device: int = 0):
        cfg = dict(SD15_UNET if unet_cfg is None else unet_cfg)
        self.device = torch.device("cuda", device)
        self.unet = SyntheticPipeline._UNetStub(cfg)
        self.scheduler = DDIMTables(num_inference_steps, steps_offset=steps_offset)
        self.tokenizer = WordTokenizer()
        eng = UNetEngine(cfg, max_samples=5 * max_batch, max_contexts=1 + 2 * max_batch, device=device)
        eng.load_random_weights(seed=seed)
        self._hedit_b200_engine = eng
        self.text_encoder = random_text_engine(text_cfg, seed=seed + 1, device=device)
        self._hedit_b200_text = self.text_encoder
