"""Native reward networks of the face-swapping sampler: drop-ins for what `h_Edit_R` does with its `idloss` / `lpipsloss` arguments
(face-swapping/inversion/h_edit_R.py:104-133): `loss = idloss.get_cosine_loss(x0)`, `torch.autograd.grad(loss, x)` and the same for
`lpipsloss.get_lpips_loss`.  `ArcFaceEngine` / `LpipsEngine` hold the frozen networks on the native kernels (csrc/reward.cu) and return
the loss and its image gradient in one call, without autograd.

`native_id_grad(idloss)` / `native_lpips_grad(lpipsloss, n)` recognise the reference's modules (`IDLoss.facenet` = Backbone(112, 50,
'ir_se'), arcface_model.py:16; `LPIPS_Loss.lpips_loss` = lpips.LPIPS(net='vgg'), arcface_model.py:76) by their state_dict layout and
return gradient hooks for `FaceUNetEngine.edit`; anything else keeps the torch-autograd plug-in route of face.py."""
from __future__ import annotations

import ctypes as C
import re
from typing import Optional

import torch

from . import _lib

VGG16_CONV_INDEX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)      # torchvision vgg16().features conv positions
LPIPS_TAP_CHANNELS = (64, 128, 256, 512, 512)


def _dims(t):
    return (C.c_int64 * max(1, t.dim()))(*(t.shape if t.dim() else (1,)))


class _Engine:
    _prefix = ""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        self.device = device
        self.handle = getattr(self.lib, f"hedit_{self._prefix}_create")(device)
        if not self.handle:
            raise RuntimeError(f"hedit_b200: {self._prefix} engine creation failed: " + _lib.last_error())
        self.last_stats = {}

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            getattr(self.lib, f"hedit_{self._prefix}_destroy")(h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _load(self, name: str, t: torch.Tensor):
        t = t.detach().to(torch.float32).contiguous()
        _lib.check(getattr(self.lib, f"hedit_{self._prefix}_load_tensor")(self.handle, name.encode(), t.data_ptr(), _dims(t), max(1, t.dim())), f"load {name}")

    def _finalize(self):
        _lib.check(getattr(self.lib, f"hedit_{self._prefix}_finalize")(self.handle), f"finalize {self._prefix} weights")

    def _img(self, x: torch.Tensor) -> torch.Tensor:
        return x.detach().to(torch.device("cuda", self.device), torch.float32).contiguous()


class ArcFaceEngine(_Engine):
    """IDLoss (arcface_model.py:12-70) on the native IR-SE50.  Images are (B,3,256,256) in [-1, 1]."""
    _prefix = "arcface"

    @staticmethod
    def is_irse50_state_dict(sd) -> bool:
        return ("input_layer.0.weight" in sd and "body.23.res_layer.5.fc2.weight" in sd and "body.24.res_layer.1.weight" not in sd
                and "output_layer.3.weight" in sd and tuple(sd["output_layer.3.weight"].shape) == (512, 512 * 49))

    def load_state_dict(self, sd) -> None:
        for name, t in sd.items():
            if torch.is_floating_point(t):
                self._load(name, t)
        self._finalize()

    @classmethod
    def from_facenet(cls, facenet: torch.nn.Module, device: int = 0) -> "ArcFaceEngine":
        eng = cls(device)
        eng.load_state_dict(facenet.state_dict())
        return eng

    def features(self, img: torch.Tensor) -> torch.Tensor:
        """l2-normalised embedding of `IDLoss.extract_feats` (arcface_model.py:41-47)."""
        x = self._img(img)
        assert x.dim() == 4 and x.shape[1:] == (3, 256, 256), "ArcFaceEngine takes (B,3,256,256) images"
        out = torch.empty(x.shape[0], 512, device=x.device, dtype=torch.float32)
        _lib.check(self.lib.hedit_arcface_features(self.handle, x.data_ptr(), x.shape[0], out.data_ptr(), self._stream()), "arcface features")
        return out

    def set_reference(self, ref_img: torch.Tensor) -> None:
        x = self._img(ref_img).reshape(-1, 3, 256, 256)[:1].contiguous()
        self._ref = x
        _lib.check(self.lib.hedit_arcface_set_reference(self.handle, x.data_ptr(), self._stream()), "arcface set_reference")

    def loss_grad(self, img: torch.Tensor, grad_out: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None):
        """(loss (B,), grad (B,3,256,256)): loss[b] = 1 - cos(ref, f(img[b])) = `get_cosine_loss(img[b:b+1])`, grad = its image gradient."""
        x = img if (img.is_cuda and img.dtype == torch.float32 and img.is_contiguous()) else self._img(img)
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 256, 256) or x.device.index != self.device:
            raise ValueError(f"ArcFaceEngine takes (B,3,256,256) images on cuda:{self.device}, got {tuple(x.shape)} on {x.device}")
        B = x.shape[0]
        grad = torch.empty_like(x) if grad_out is None else grad_out
        loss = torch.empty(B, device=x.device, dtype=torch.float32) if loss_out is None else loss_out
        if grad.shape != x.shape or not grad.is_contiguous() or grad.dtype != torch.float32 or loss.numel() != B:
            raise ValueError("grad_out / loss_out must be contiguous fp32 buffers of the image's / batch's shape")
        n = _lib.check(self.lib.hedit_arcface_loss_grad(self.handle, x.data_ptr(), B, loss.data_ptr(), grad.data_ptr(), self._stream()), "arcface loss_grad")
        self.last_stats = {"kernel_launches": n, "flops": self.lib.hedit_arcface_last_flops(self.handle)}
        return loss, grad


def lpips_native_tensors(sd) -> dict:
    """Map an LPIPS-VGG state_dict (lpips package: `net.slice{j}.{idx}.weight`, `lin{k}.model.1.weight`, `scaling_layer.shift`; or
    hedit_b200.reward_nets.LPIPSVGG16: `features.{idx}.weight`, `lins.{k}`, `shift`) to the C-ABI's tensor names."""
    out = {}
    for name, t in sd.items():
        m = re.search(r"(?:slice\d+|features)\.(\d+)\.(weight|bias)$", name)
        if m and int(m.group(1)) in VGG16_CONV_INDEX:
            out[f"conv{VGG16_CONV_INDEX.index(int(m.group(1)))}.{m.group(2)}"] = t
            continue
        m = re.search(r"(?:^|\.)lin(\d)\.model\.1\.weight$", name) or re.search(r"(?:^|\.)lins\.(\d)(?:\.model\.1\.weight)?$", name)
        if m:
            out.setdefault(f"lin{m.group(1)}.weight", t.reshape(-1))
            continue
        if name.endswith("shift") or name.endswith("scale"):
            out[name.rsplit(".", 1)[-1]] = t.reshape(-1)
    return out


class LpipsEngine(_Engine):
    """LPIPS_Loss (arcface_model.py:72-95) on the native VGG16 feature stack."""
    _prefix = "lpips"

    def load_state_dict(self, sd) -> None:
        t = lpips_native_tensors(sd)
        want = [f"conv{i}.{k}" for i in range(13) for k in ("weight", "bias")] + [f"lin{k}.weight" for k in range(5)] + ["shift", "scale"]
        missing = [k for k in want if k not in t]
        if missing:
            raise KeyError(f"hedit_b200: not an LPIPS-VGG16 state_dict (missing {missing[:4]} ...)")
        for k in want:
            self._load(k, t[k])
        self._finalize()

    @classmethod
    def from_module(cls, lpips_module: torch.nn.Module, device: int = 0) -> "LpipsEngine":
        eng = cls(device)
        eng.load_state_dict(lpips_module.state_dict())
        return eng

    def set_source(self, src_img: torch.Tensor) -> None:
        x = self._img(src_img)
        assert x.dim() == 4 and x.shape[1] == 3 and x.shape[2] == x.shape[3]
        self._src = x
        _lib.check(self.lib.hedit_lpips_set_source(self.handle, x.data_ptr(), x.shape[0], x.shape[2], self._stream()), "lpips set_source")

    def loss_grad(self, img: torch.Tensor, grad_out: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None):
        x = img if (img.is_cuda and img.dtype == torch.float32 and img.is_contiguous()) else self._img(img)
        src = getattr(self, "_src", None)
        if src is None:
            raise RuntimeError("LpipsEngine.set_source(...) first")
        if x.dim() != 4 or tuple(x.shape[1:]) != tuple(src.shape[1:]) or src.shape[0] not in (1, x.shape[0]) or x.device.index != self.device:
            raise ValueError(f"LpipsEngine: images must be (B,{','.join(str(v) for v in src.shape[1:])}) on cuda:{self.device} with B matching the "
                             f"{src.shape[0]} source image(s), got {tuple(x.shape)} on {x.device}")
        B = x.shape[0]
        grad = torch.empty_like(x) if grad_out is None else grad_out
        loss = torch.empty(B, device=x.device, dtype=torch.float32) if loss_out is None else loss_out
        if grad.shape != x.shape or not grad.is_contiguous() or grad.dtype != torch.float32 or loss.numel() != B:
            raise ValueError("grad_out / loss_out must be contiguous fp32 buffers of the image's / batch's shape")
        n = _lib.check(self.lib.hedit_lpips_loss_grad(self.handle, x.data_ptr(), B, loss.data_ptr(), grad.data_ptr(), self._stream()), "lpips loss_grad")
        self.last_stats = {"kernel_launches": n, "flops": self.lib.hedit_lpips_last_flops(self.handle)}
        return loss, grad


def native_id_grad(idloss, device: int = 0):
    """Gradient hook x0 -> d(sum_b get_cosine_loss(x0[b])) / d x0 on the native IR-SE50, or None when `idloss` is not the reference's
    IDLoss layout (an object with `.facenet` = IR-SE50 Backbone and `.ref` = the reference face)."""
    net, ref = getattr(idloss, "facenet", None), getattr(idloss, "ref", None)
    if net is None or ref is None or not hasattr(net, "state_dict") or tuple(ref.shape[-3:]) != (3, 256, 256):
        return None
    if type(idloss).get_cosine_loss is not getattr(_stock_class(idloss, "IDLoss"), "get_cosine_loss", None):
        return None
    sd = net.state_dict()
    if not ArcFaceEngine.is_irse50_state_dict(sd):
        return None
    eng = getattr(idloss, "_hedit_b200_arcface", None)
    if eng is None or eng.device != device:
        eng = ArcFaceEngine(device)
        eng.load_state_dict(sd)
        eng.set_reference(ref)
        idloss._hedit_b200_arcface = eng

    def fn(x0: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if tuple(x0.shape[-2:]) != (256, 256):
            raise ValueError("native ArcFace reward takes 256x256 images")
        return eng.loss_grad(x0, grad_out=out, loss_out=_loss_buf(fn, x0))[1]

    fn.engine = eng
    return fn


def native_lpips_grad(lpipsloss, device: int = 0):
    """Gradient hook on the native VGG16-LPIPS, or None when `lpipsloss` is not the reference's LPIPS_Loss layout (`.lpips_loss` with the
    lpips-package VGG state_dict, `.src` = the anchor image(s))."""
    mod, src = getattr(lpipsloss, "lpips_loss", None), getattr(lpipsloss, "src", None)
    if mod is None or src is None or not hasattr(mod, "state_dict"):
        return None
    if type(lpipsloss).get_lpips_loss is not getattr(_stock_class(lpipsloss, "LPIPS_Loss"), "get_lpips_loss", None):
        return None
    t = lpips_native_tensors(mod.state_dict())
    if "conv12.weight" not in t or "lin4.weight" not in t or "shift" not in t or src.shape[-1] not in (128, 256, 512):
        return None
    if getattr(mod, "pnet_type", "vgg") not in ("vgg", "vgg16") or getattr(mod, "spatial", False) or not getattr(mod, "lpips", True):
        return None
    eng = getattr(lpipsloss, "_hedit_b200_lpips", None)
    if eng is None or eng.device != device:
        eng = LpipsEngine(device)
        eng.load_state_dict(mod.state_dict())
        eng.set_source(src)
        lpipsloss._hedit_b200_lpips = eng

    def fn(x0: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return eng.loss_grad(x0, grad_out=out, loss_out=_loss_buf(fn, x0))[1]

    fn.engine = eng
    return fn


def _loss_buf(fn, x0: torch.Tensor) -> torch.Tensor:
    """Persistent per-hook loss buffer: with fixed (image, gradient, loss) addresses the engine replays the whole forward + backward from
    one CUDA graph (netexec.h)."""
    buf = getattr(fn, "loss", None)
    if buf is None or buf.shape[0] != x0.shape[0] or buf.device != x0.device:
        buf = torch.empty(x0.shape[0], device=x0.device, dtype=torch.float32)
        fn.loss = buf
    return buf


def _stock_class(obj, name: str):
    """The class named `name` in obj's MRO (the reference's own class), so a user subclass overriding the loss method is detected."""
    for k in type(obj).__mro__:
        if k.__name__ == name:
            return k
    return type(obj)
