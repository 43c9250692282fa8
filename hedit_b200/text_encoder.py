"""Python handle on the native CLIP text tower (C ABI: hedit_text_*): `model.text_encoder(ids)[0]` as used by the reference's
`encode_text` (text-guided/inversion/inversion_utils.py:13-36).  Built from a transformers `CLIPTextModel`'s state dict."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class TextEncoderEngine:
    def __init__(self, vocab: int, width: int, heads: int, layers: int, ffn: int, tokens: int = 77, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        c = _lib.TextConfigC(vocab, width, heads, layers, ffn, tokens)
        self.width, self.tokens, self.device = width, tokens, device
        self.handle = self.lib.hedit_text_create(C.byref(c), device)
        if not self.handle:
            raise RuntimeError("hedit_b200: text encoder creation failed: " + _lib.last_error())

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hedit_text_destroy(h)

    @classmethod
    def from_text_encoder(cls, text_encoder, device: int = 0) -> "TextEncoderEngine":
        cfg = text_encoder.config
        if getattr(cfg, "hidden_act", "quick_gelu") != "quick_gelu":
            raise ValueError("only the quick_gelu CLIP text towers (SD-1.x: openai/clip-vit-large-patch14) are supported")
        eng = cls(cfg.vocab_size, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, cfg.intermediate_size,
                  cfg.max_position_embeddings, device)
        for name, t in text_encoder.state_dict().items():
            if not torch.is_floating_point(t):
                continue
            t = t.detach().to(torch.float32).contiguous()
            dims = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(eng.lib.hedit_text_load_tensor(eng.handle, name.encode(), t.data_ptr(), dims, t.dim()), f"load {name}")
        _lib.check(eng.lib.hedit_text_finalize(eng.handle), "finalize text encoder weights")
        return eng

    def __call__(self, input_ids: torch.Tensor):
        """ids (B, tokens) -> (last_hidden_state (B, tokens, width),) like `model.text_encoder(ids)`."""
        ids = np.ascontiguousarray(input_ids.detach().cpu().numpy().astype(np.int32))
        B = ids.shape[0]
        assert ids.shape[1] == self.tokens
        dev = torch.device("cuda", self.device)
        out = torch.empty(B, self.tokens, self.width, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.hedit_text_encode(self.handle, ids.ctypes.data, B, out.data_ptr(), stream), "text encode")
        return (out,)
