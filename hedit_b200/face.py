"""Face swapping: drop-in for the reference's face-swapping/inversion/h_edit_R.py (`h_Edit_R`) and its pixel-space DDPM denoiser
(face-swapping/diffusion/diffusion.py `Model`).  The denoiser forward, the DDPM-style reverse step, the Tweedie prediction and the
reward-guided updates run in the native loop (csrc/face.cu, csrc/face_loop.cu); the two reward gradients (ArcFace identity loss,
LPIPS) are evaluated through the reference's reward-model protocol (`idloss.get_cosine_loss`, `lpipsloss.get_lpips_loss`: the
caller's torch modules, differentiated by torch.autograd on the same CUDA stream)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


def face_config_of(model) -> dict:
    cfg = getattr(model, "config", None) or getattr(model, "cfg", None)
    get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    return dict(ch=get("ch", 128), ch_mult=tuple(get("ch_mult", (1, 1, 2, 2, 4, 4))), num_res_blocks=get("num_res_blocks", 2),
                attn_resolution=int(tuple(get("attn_resolutions", (16,)))[0]), image_size=get("image_size", 256),
                in_channels=get("in_channels", 3), out_ch=get("out_ch", 3))


class FaceUNetEngine:
    def __init__(self, config: dict, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("hedit_b200: no CUDA device visible; the B200 path has no CPU fallback")
        c = _lib.FaceConfigC()
        c.ch, c.n_levels = config["ch"], len(config["ch_mult"])
        for i, v in enumerate(config["ch_mult"]):
            c.ch_mult[i] = v
        c.num_res_blocks, c.attn_resolution, c.image_size = config["num_res_blocks"], config["attn_resolution"], config["image_size"]
        c.in_channels, c.out_ch = config["in_channels"], config["out_ch"]
        self.config, self.device = dict(config), device
        self.handle = self.lib.hedit_face_create(C.byref(c), device)
        if not self.handle:
            raise RuntimeError("hedit_b200: face UNet engine creation failed: " + _lib.last_error())
        self.last_stats: Dict[str, int] = {}

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hedit_face_destroy(h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_state_dict(self, sd) -> None:
        for name, t in sd.items():
            if not torch.is_floating_point(t) or name == "logvar":
                continue
            t = t.detach().to(torch.float32).contiguous()
            dims = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.hedit_face_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, t.dim()), f"load {name}")
        _lib.check(self.lib.hedit_face_finalize(self.handle), "finalize face UNet weights")

    @classmethod
    def from_model(cls, model, device: int = 0) -> "FaceUNetEngine":
        eng = cls(face_config_of(model), device)
        eng.load_state_dict(model.state_dict())
        return eng

    def tensor_specs(self):
        out, buf, dims = [], C.create_string_buffer(256), (C.c_int64 * 4)()
        for i in range(self.lib.hedit_face_tensor_count(self.handle)):
            nd = _lib.check(self.lib.hedit_face_tensor_info(self.handle, i, buf, 256, dims), "tensor_info")
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
            _lib.check(self.lib.hedit_face_load_tensor(self.handle, name.encode(), t.data_ptr(), dims, len(shape)), f"load {name}")
        _lib.check(self.lib.hedit_face_finalize(self.handle), "finalize face UNet weights")

    def forward(self, x: torch.Tensor, t, out: torch.Tensor = None) -> torch.Tensor:
        """eps = model(x, t) (diffusion.py:301): x (S,3,R,R); t scalar or (S,).  Calls that repeat the same (x, out) buffers and batch
        are replayed from a CUDA graph from the third call on (netexec.h); results are bit-identical to the direct launches."""
        dev = torch.device("cuda", self.device)
        x = x.detach().to(dev, torch.float32).contiguous()
        S = x.shape[0]
        tt = np.ascontiguousarray(np.broadcast_to(np.asarray(torch.as_tensor(t).detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float32).reshape(-1), (S,)))
        eps = torch.empty_like(x) if out is None else out
        assert eps.is_contiguous() and eps.shape == x.shape and eps.dtype == torch.float32 and eps.device == x.device
        n = _lib.check(self.lib.hedit_face_unet_forward(self.handle, x.data_ptr(), tt.ctypes.data, S, eps.data_ptr(), self._stream()), "face unet forward")
        self.last_stats = {"kernel_launches": n, "sample_forwards": S, "flops": self.lib.hedit_face_last_flops(self.handle)}
        return eps

    __call__ = forward

    def edit(self, xT, zs, coef: np.ndarray, weight: float, optimization_steps: int, id_grad=None, lpips_grad=None, mask=None):
        """xT (B,3,R,R), zs (B,steps,3,R,R) on this device; coef (steps, 8) rows of hedit_face_step_coef; id_grad / lpips_grad:
        callables x0 (B,3,R,R) -> d loss / d x0 or None."""
        dev = torch.device("cuda", self.device)
        xT = xT.detach().to(dev, torch.float32).contiguous()
        zs = zs.detach().to(dev, torch.float32).contiguous()
        B, steps = xT.shape[0], zs.shape[1]
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        assert coef.shape == (steps, 8)
        a = _lib.FaceArgsC()
        a.B, a.steps, a.opt_steps = B, steps, int(optimization_steps)
        a.xT, a.zs, a.coef, a.weight = xT.data_ptr(), zs.data_ptr(), coef.ctypes.data, float(weight)
        edited = torch.empty_like(xT)
        x0_buf, grad_buf = torch.empty_like(xT), torch.zeros_like(xT)
        keep = [xT, zs, coef, edited, x0_buf, grad_buf]
        if mask is not None:
            m = mask.detach().to(dev, torch.float32).expand_as(xT).contiguous()
            keep.append(m)
            a.mask = m.data_ptr()
        a.use_id, a.use_lpips = int(id_grad is not None), int(lpips_grad is not None)
        errs = []

        def _cb(_user, which, step, opt_step):
            try:
                f = id_grad if which == 0 else lpips_grad
                if hasattr(f, "engine"):           # native reward network: writes the gradient in place
                    f(x0_buf, out=grad_buf)
                else:
                    grad_buf.copy_(f(x0_buf).reshape(xT.shape).to(torch.float32))
                return 0
            except Exception as exc:
                errs.append(exc)
                return -1

        cfn = _lib.REWARD_FN(_cb)
        keep.append(cfn)
        a.reward = C.cast(cfn, C.c_void_p)
        a.reward_x0, a.reward_grad, a.edited = x0_buf.data_ptr(), grad_buf.data_ptr(), edited.data_ptr()
        rc = self.lib.hedit_face_edit(self.handle, C.byref(a), self._stream())
        if errs:
            raise errs[0]
        _lib.check(rc, "face edit")
        self.last_stats = {"sample_forwards": int(a.n_sample_forwards), "kernel_launches": int(a.n_kernel_launches)}
        return edited


def face_step_tables(betas: torch.Tensor, seq, num_inference_steps: int, after_skip_steps: int, eta=1.0) -> np.ndarray:
    """Per-step scalars of h_Edit_R (h_edit_R.py:37-54,68-88,106) in the reference's fp32 operation order."""
    etas = [eta] * num_inference_steps if isinstance(eta, (int, float)) else list(eta)
    ab = (1.0 - betas.detach().float().cpu()).cumprod(dim=0)
    op = [int(v) for v in list(seq)[-after_skip_steps:]]
    out = np.zeros((len(op), 8), dtype=np.float32)
    for i, t in enumerate(op):
        idx = num_inference_steps - i - (num_inference_steps - after_skip_steps + 1)
        tm1 = op[i + 1] if i < len(op) - 1 else 0
        e = 0.5                                                   # the step hard-codes eta = 0.5 for the c1 / c2 split (:82)
        c1 = (1 - ab[tm1]).sqrt() * e
        c2 = (1 - ab[tm1]).sqrt() * ((1 - e ** 2) ** 0.5)
        out[i] = [t, tm1, float((1 - ab[t]) ** 0.5), float(ab[t] ** 0.5), float((1 - ab[tm1]) ** 0.5), float(ab[tm1].sqrt()), float(c2), float(etas[idx] * c1)]
    return out


def _reward_grad(loss_fn):
    """d(sum_b loss(x0[b])) / d x0 through the caller's reward module: the reference differentiates a batch-mean loss for its single
    image (arcface_model.py:67,95); per image that is the same number."""
    if loss_fn is None:
        return None

    def fn(x0: torch.Tensor) -> torch.Tensor:
        with torch.enable_grad():
            x = x0.detach().clone().requires_grad_(True)
            total = None
            for b in range(x.shape[0]):
                l = loss_fn(x[b:b + 1])
                total = l if total is None else total + l
            return torch.autograd.grad(outputs=total, inputs=x)[0]

    return fn


def get_face_engine(model, device: int = 0) -> FaceUNetEngine:
    eng = getattr(model, "_hedit_b200_face", None)
    if eng is None:
        eng = FaceUNetEngine.from_model(model, device)
        model._hedit_b200_face = eng
    return eng


def h_Edit_R(model, lpipsloss, idloss, xT, betas, seq, eta=1.0, zs=None, weight_edit_face=50.0, optimization_steps=3, after_skip_steps=100,
             num_inference_steps=100, soft_face_mask=None):
    """Reference signature (face-swapping/inversion/h_edit_R.py:7).  Returns the edited sample (B,3,R,R)."""
    dev = xT.device
    x = xT if xT.dim() == 4 else xT.unsqueeze(0)
    B = x.shape[0]
    eng = get_face_engine(model, dev.index or 0 if dev.type == "cuda" else 0)
    z = zs[:after_skip_steps]
    z = z.reshape(1, after_skip_steps, *x.shape[-3:]).expand(B, -1, -1, -1, -1) if z.dim() == 4 else z
    coef = face_step_tables(betas, seq, num_inference_steps, after_skip_steps, eta)
    # the reference's own reward modules (IR-SE50 IDLoss, VGG16 LPIPS_Loss) run on the native kernels (reward.py: loss + image gradient,
    # no autograd); any other object keeps the reward-model protocol and is differentiated by torch.autograd
    from .reward import native_id_grad, native_lpips_grad
    native = os.environ.get("HEDIT_NATIVE_REWARD", "1") != "0" and tuple(x.shape[-2:]) == (256, 256)
    id_fn = (native_id_grad(idloss, eng.device) if (idloss and native) else None) or _reward_grad(idloss.get_cosine_loss if idloss else None)
    lp_fn = (native_lpips_grad(lpipsloss, eng.device) if (lpipsloss and native) else None) or _reward_grad(lpipsloss.get_lpips_loss if lpipsloss else None)
    out = eng.edit(x, z, coef, weight_edit_face, optimization_steps, id_grad=id_fn, lpips_grad=lp_fn, mask=soft_face_mask)
    eng.last_stats["native_rewards"] = (hasattr(id_fn, "engine"), hasattr(lp_fn, "engine"))
    return out.to(dev)


@torch.no_grad()
def sample_xts_from_x0_sde(model, x0, betas, seq, num_inference_steps=100):
    """Reference signature (face-swapping/inversion/sde_inversion.py:4): independent draws x_t ~ q(x_t | x_0), seeded 42 like the reference."""
    torch.manual_seed(42)
    torch.cuda.manual_seed(42)
    ab = (1.0 - betas).cumprod(dim=0)
    pos = {int(v): k for k, v in enumerate(seq)}
    shape = (num_inference_steps + 1,) + tuple(x0.shape[1:])
    xts = torch.zeros(shape, device=x0.device)
    noise_added = torch.zeros(shape, device=x0.device)
    xts[0] = x0[0]
    for t in reversed(list(seq)):
        idx = num_inference_steps - pos[int(t)]
        noise = torch.randn_like(x0)
        xts[idx] = (x0 * (ab[int(t)] ** 0.5) + noise * ((1 - ab[int(t)]) ** 0.5))[0]
        noise_added[idx] = noise[0]
    return xts, noise_added


@torch.no_grad()
def inversion_forward_process_sde(model, x0, betas, seq, etas=1.0, num_inference_steps=100, device=None):
    """Reference signature (face-swapping/inversion/sde_inversion.py:54); returns (xt, zs, xts, noise_added).  Given the independently
    sampled x_t's the T noise predictions do not depend on each other, so they are ONE batched denoiser call on the native engine instead
    of T calls of batch 1."""
    assert not (etas is None or (isinstance(etas, (int, float)) and etas == 0)), "eta must be > 0 (reference assert, sde_inversion.py:127)"
    T = num_inference_steps
    etas = [etas] * T if isinstance(etas, (int, float)) else list(etas)
    ts = [int(t) for t in seq]
    xts, noise_added = sample_xts_from_x0_sde(model, x0, betas, seq, num_inference_steps=T)
    eng = get_face_engine(model, x0.device.index or 0 if x0.device.type == "cuda" else 0)
    dev = torch.device("cuda", eng.device)
    ab = (1.0 - betas.to(dev)).cumprod(dim=0)
    x_in = torch.stack([xts[T - k] for k in range(T)]).to(dev)              # step k (timestep ts[k]) starts from xts[T - k]
    eps = torch.cat([eng(x_in[lo:lo + 16], ts[lo:lo + 16]) for lo in range(0, T, 16)])
    zs = torch.zeros((T,) + tuple(x0.shape[1:]), device=x0.device)
    for k, t in enumerate(ts):
        idx = T - k - 1
        tm1 = ts[k + 1] if k < T - 1 else 0
        x0_hat = (x_in[k] - (1 - ab[t]) ** 0.5 * eps[k]) / ab[t] ** 0.5
        c1 = (1 - ab[tm1]).sqrt() * 0.5
        c2 = (1 - ab[tm1]).sqrt() * ((1 - 0.5 ** 2) ** 0.5)
        mu = ab[tm1].sqrt() * x0_hat + c2 * eps[k]
        z = (xts[idx].to(dev) - mu) / (etas[idx] * c1)
        zs[idx] = z.to(x0.device)
        xts[idx] = (mu + (etas[idx] * c1) * z).to(x0.device)
    return xts[1][None], zs, xts, noise_added
