"""Python handle on the native CLIP-Gram style reward (C ABI: hedit_clip_* in include/hedit_b200.h): the loss
`||Gram(features(img)) - Gram(features(ref))||_F` of the reference's `CLIPEncoder.get_gram_matrix_residual`
(text-guided-n-style/clip_guidance/base_clip.py:55-66) and its gradient with respect to the image."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class ClipGramEngine:
    def __init__(self, resolution: int, patch: int, width: int, heads: int, layers: int = 3, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        c = _lib.ClipConfigC(resolution, patch, width, heads, layers)
        self.device, self.resolution = device, resolution
        self.handle = self.lib.hedit_clip_create(C.byref(c), device)
        if not self.handle:
            raise RuntimeError("hedit_b200: CLIP engine creation failed: " + _lib.last_error())
        self._shape = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hedit_clip_destroy(h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_state_dict(self, sd) -> None:
        """sd = state dict of the CLIP image tower (`clip_model.visual`); tensors this path does not use are skipped."""
        for name, t in sd.items():
            if not torch.is_floating_point(t):
                continue
            t = t.detach().to(torch.float32).contiguous()
            dims = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.hedit_clip_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, t.dim()), f"load {name}")
        _lib.check(self.lib.hedit_clip_finalize(self.handle), "finalize CLIP weights")

    @classmethod
    def from_image_encoder(cls, image_encoder, device: int = 0) -> "ClipGramEngine":
        """image_encoder: the reference's `CLIPEncoder` (attributes `.clip_model.visual`, `.ref`) or an object with `.visual`, `.ref`."""
        vis = image_encoder.clip_model.visual if hasattr(image_encoder, "clip_model") else image_encoder.visual
        width = vis.conv1.weight.shape[0]
        patch = vis.conv1.weight.shape[-1]
        res = int(round((vis.positional_embedding.shape[0] - 1) ** 0.5)) * patch
        heads = vis.transformer.resblocks[0].attn.num_heads
        eng = cls(res, patch, width, heads, 3, device)
        eng.load_state_dict(vis.state_dict())
        eng.set_reference(image_encoder.ref)
        return eng

    def set_reference(self, ref: torch.Tensor) -> None:
        ref = ref.detach().to(torch.device("cuda", self.device), torch.float32).contiguous()
        assert tuple(ref.shape) == (1, 3, self.resolution, self.resolution)
        _lib.check(self.lib.hedit_clip_set_reference(self.handle, ref.data_ptr(), self._stream()), "clip set_reference")

    def loss(self, img: torch.Tensor) -> torch.Tensor:
        """img (B,3,H,W) in [-1,1] -> per-image loss (B,)."""
        dev = torch.device("cuda", self.device)
        img = img.detach().to(dev, torch.float32).contiguous()
        B, _, H, W = img.shape
        out = torch.empty(B, dtype=torch.float32, device=dev)
        _lib.check(self.lib.hedit_clip_gram_loss(self.handle, img.data_ptr(), B, H, W, out.data_ptr(), self._stream()), "clip gram loss")
        self._shape = tuple(img.shape)
        return out

    def backward(self) -> torch.Tensor:
        """d loss[b] / d img[b] (B,3,H,W) for the last loss() call."""
        assert self._shape is not None, "call loss() first"
        dimg = torch.empty(self._shape, dtype=torch.float32, device=torch.device("cuda", self.device))
        _lib.check(self.lib.hedit_clip_gram_backward(self.handle, dimg.data_ptr(), self._stream()), "clip gram backward")
        return dimg
