"""Python handle on the native VAE decoder (C ABI: hedit_vae_* in include/hedit_b200.h): `decode(z).sample` of the reference's
`model.vae` (text-guided/main_p2p.py:262-275) and the input gradient of that decode which the style path back-propagates through
(text-guided-n-style/inversion/h_edit.py:158-164).  PyTorch only owns the device buffers."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib


def vae_config_of(vae) -> dict:
    cfg = getattr(vae, "cfg", None) or getattr(vae, "config", None)
    get = (lambda k, d=None: getattr(cfg, k, d)) if not isinstance(cfg, dict) else (lambda k, d=None: cfg.get(k, d))
    return dict(latent_channels=get("latent_channels", 4), out_channels=get("out_channels", 3),
                block_out_channels=tuple(get("block_out_channels", (128, 256, 512, 512))), layers_per_block=get("layers_per_block", 2),
                norm_groups=get("norm_num_groups", 32))


class _Sample:
    def __init__(self, sample):
        self.sample = sample


class VaeDecoderEngine:
    def __init__(self, config: dict, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        c = _lib.VaeConfigC()
        c.latent_channels, c.out_channels = config["latent_channels"], config["out_channels"]
        for i, v in enumerate(config["block_out_channels"]):
            c.block_out_channels[i] = v
        c.layers_per_block, c.norm_groups = config["layers_per_block"], config["norm_groups"]
        self.config, self.device = dict(config), device
        self.handle = self.lib.hedit_vae_create(C.byref(c), device)
        if not self.handle:
            raise RuntimeError("hedit_b200: VAE engine creation failed: " + _lib.last_error())
        self._shape = None
        self.last_stats: Dict[str, float] = {}

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hedit_vae_destroy(h)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for name, t in sd.items():
            if not torch.is_floating_point(t) or name.startswith(("encoder.", "quant_conv.")):
                continue
            t = t.detach().to(torch.float32).contiguous()
            dims = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.hedit_vae_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, t.dim()), f"load {name}")
        _lib.check(self.lib.hedit_vae_finalize(self.handle), "finalize VAE weights")

    @classmethod
    def from_vae(cls, vae, device: int = 0) -> "VaeDecoderEngine":
        eng = cls(vae_config_of(vae), device)
        eng.load_state_dict(vae.state_dict())
        return eng

    def tensor_specs(self):
        out, buf, dims = [], C.create_string_buffer(256), (C.c_int64 * 4)()
        for i in range(self.lib.hedit_vae_tensor_count(self.handle)):
            nd = _lib.check(self.lib.hedit_vae_tensor_info(self.handle, i, buf, 256, dims), "tensor_info")
            out.append((buf.value.decode(), tuple(int(dims[k]) for k in range(nd))))
        return out

    def load_random_weights(self, seed: int = 0) -> None:
        g = torch.Generator(device=f"cuda:{self.device}").manual_seed(seed)
        dev = torch.device("cuda", self.device)
        for name, shape in self.tensor_specs():
            if len(shape) >= 2:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * (3.0 / fan_in) ** 0.5
            elif name.endswith("weight"):
                t = 0.8 + 0.4 * torch.rand(shape, generator=g, device=dev)
            else:
                t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * 0.1
            dims = (C.c_int64 * len(shape))(*shape)
            _lib.check(self.lib.hedit_vae_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, len(shape)), f"load {name}")
        _lib.check(self.lib.hedit_vae_finalize(self.handle), "finalize VAE weights")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def decode_tensor(self, z: torch.Tensor) -> torch.Tensor:
        """z (B,4,h,w) -> image (B,3,8h,8w), both fp32 on this device."""
        dev = torch.device("cuda", self.device)
        z = z.detach().to(dev, torch.float32).contiguous()
        B, _, h, w = z.shape
        img = torch.empty(B, self.config["out_channels"], 8 * h, 8 * w, dtype=torch.float32, device=dev)
        n = _lib.check(self.lib.hedit_vae_decode(self.handle, z.data_ptr(), img.data_ptr(), B, h, w, self._stream()), "vae decode")
        self._shape = tuple(z.shape)
        self.last_stats = {"kernel_launches": n, "flops": self.lib.hedit_vae_last_flops(self.handle)}
        return img

    def decode(self, z: torch.Tensor) -> _Sample:
        """`model.vae.decode(z).sample` protocol."""
        return _Sample(self.decode_tensor(z))

    def backward(self, dimg: torch.Tensor) -> torch.Tensor:
        """dLoss/dimg (B,3,8h,8w) of the last decode -> dLoss/dz (B,4,h,w).  The gradient is rescaled per image to unit max-abs on the
        way in (16-bit conv operands) and scaled back on the way out, so callers see plain gradients."""
        assert self._shape is not None, "call decode() first"
        dev = torch.device("cuda", self.device)
        dimg = dimg.detach().to(dev, torch.float32)
        scale = dimg.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
        g = (dimg / scale).contiguous()
        dz = torch.empty(self._shape, dtype=torch.float32, device=dev)
        n = _lib.check(self.lib.hedit_vae_decode_backward(self.handle, g.data_ptr(), dz.data_ptr(), self._stream()), "vae decode backward")
        self.last_stats = {"kernel_launches": n, "flops": self.lib.hedit_vae_last_flops(self.handle)}
        return dz * scale


class _Gaussian:
    def __init__(self, moments):
        self.mean, self.logvar = moments.chunk(2, dim=1)

    def mode(self):
        return self.mean


class _EncOut:
    def __init__(self, moments):
        self.latent_dist = _Gaussian(moments)


class VaeEncoderEngine:
    """`model.vae.encode(x).latent_dist.mode()` (text-guided/main_p2p.py:154-159) on the native encoder."""

    def __init__(self, config: dict, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        c = _lib.VaeConfigC()
        c.latent_channels, c.out_channels = config["latent_channels"], config["out_channels"]
        for i, v in enumerate(config["block_out_channels"]):
            c.block_out_channels[i] = v
        c.layers_per_block, c.norm_groups = config["layers_per_block"], config["norm_groups"]
        self.config, self.device = dict(config), device
        self.handle = self.lib.hedit_vae_enc_create(C.byref(c), device)
        if not self.handle:
            raise RuntimeError("hedit_b200: VAE encoder creation failed: " + _lib.last_error())

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hedit_vae_enc_destroy(h)

    @classmethod
    def from_vae(cls, vae, device: int = 0) -> "VaeEncoderEngine":
        eng = cls(vae_config_of(vae), device)
        for name, t in vae.state_dict().items():
            if not torch.is_floating_point(t) or not name.startswith(("encoder.", "quant_conv.")):
                continue
            t = t.detach().to(torch.float32).contiguous()
            dims = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(eng.lib.hedit_vae_enc_load_tensor(eng.handle, name.encode(), t.data_ptr(), dims, t.dim()), f"load {name}")
        _lib.check(eng.lib.hedit_vae_enc_finalize(eng.handle), "finalize VAE encoder weights")
        return eng

    def encode(self, x: torch.Tensor) -> _EncOut:
        dev = torch.device("cuda", self.device)
        x = x.detach().to(dev, torch.float32).contiguous()
        B, _, H, W = x.shape
        mom = torch.empty(B, 2 * self.config["latent_channels"], H // 8, W // 8, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.hedit_vae_encode(self.handle, x.data_ptr(), mom.data_ptr(), B, H, W, stream), "vae encode")
        return _EncOut(mom)
