"""Combined text-guided + style editing: drop-in for the reference's text-guided-n-style/inversion/h_edit.py.

The text-guided part of every step (two classifier-free-guided UNet calls with P2P injection, h-term move) runs in the native
batched loop; after each implicit-loop iteration the loop hands the Tweedie prediction x0 to a reward hook and applies the
Langevin move x <- x - rho * dLoss/dx with fused element kernels (csrc/hstep.cuh: hstep_x0pred / guid_norm / guid_update).
The reward gradient runs on native kernels too: VAE decode of x0 (csrc/vae.cu, built from `model.vae`'s weights) -> CLIP-Gram
loss (csrc/clip.cu, built from `image_encoder`'s ViT weights and reference image) -> CLIP backward -> VAE backward.  Image encoders
the CLIP engine does not support fall back to differentiating `image_encoder.get_gram_matrix_residual(img)` (the reference's
reward-model protocol) with torch.autograd on the same CUDA stream."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from .samplers import h_edit_p2p_batch
from .schedule import x0_tables

VAE_SCALE = 0.18215


def clip_gram_guidance(model, image_encoder, autocast: bool = True) -> Callable[[torch.Tensor], torch.Tensor]:
    """dLoss/dx0 of loss = ||Gram(CLIP features of decode(x0 / 0.18215)) - Gram(reference style)||_F
    (text-guided-n-style/inversion/h_edit.py:155-164; clip_guidance/base_clip.py:55-66), evaluated per image: the reference's
    Gram residual only looks at batch element 0 (base_clip.py:62), so a batch is a loop over images."""

    def fn(x0: torch.Tensor) -> torch.Tensor:
        with torch.enable_grad():
            x = x0.detach().clone().requires_grad_(True)
            total = None
            for b in range(x.shape[0]):
                with torch.autocast("cuda", enabled=autocast):
                    img = model.vae.decode(1 / VAE_SCALE * x[b:b + 1]).sample
                loss = torch.linalg.norm(image_encoder.get_gram_matrix_residual(img))
                total = loss if total is None else total + loss
            return torch.autograd.grad(outputs=total, inputs=x)[0]

    return fn


def get_vae_engine(model, device: int = 0):
    """One persistent native decoder per pipeline object, built from `model.vae`'s state dict."""
    from .vae import VaeDecoderEngine
    eng = getattr(model, "_hedit_b200_vae", None)
    if eng is None:
        eng = VaeDecoderEngine.from_vae(model.vae, device=device)
        model._hedit_b200_vae = eng
    return eng


def clip_gram_guidance_native(vae_engine, image_encoder, vae_batch: int = 2) -> Callable[[torch.Tensor], torch.Tensor]:
    """Same reward gradient as `clip_gram_guidance`, with the VAE decode and its backward on the native kernels
    (csrc/vae.cu: ~2.5 TFLOP forward + ~2.5 TFLOP backward per image at 512x512, the dominant cost of a Langevin step); only the
    small CLIP branch (bicubic resize, 3 ViT blocks, Gram residual norm; ~0.03 TFLOP) is differentiated by torch.autograd.
    The decoder computes in fp32 with 16-bit tensor-core operands, which is at least the precision of the reference's
    `autocast("cuda")` decode (h_edit.py:158)."""

    def fn(x0: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(x0)
        for lo in range(0, x0.shape[0], vae_batch):
            hi = min(x0.shape[0], lo + vae_batch)
            img = vae_engine.decode_tensor(x0[lo:hi] * (1 / VAE_SCALE))
            with torch.enable_grad():
                im = img.detach().requires_grad_(True)
                total = None
                for b in range(hi - lo):
                    loss = torch.linalg.norm(image_encoder.get_gram_matrix_residual(im[b:b + 1]))
                    total = loss if total is None else total + loss
                dimg = torch.autograd.grad(outputs=total, inputs=im)[0]
            out[lo:hi] = vae_engine.backward(dimg) * (1 / VAE_SCALE)
        return out

    return fn


def get_clip_engine(image_encoder, device: int = 0):
    """Native CLIP-Gram engine built from the caller's image encoder (its ViT weights and reference style image), cached on it.
    Returns None when the encoder is not a ViT tower this engine supports (head dim 64, <= 256 tokens): the torch branch is used then."""
    from .clip_gram import ClipGramEngine
    if hasattr(image_encoder, "_hedit_b200_clip"):
        return image_encoder._hedit_b200_clip
    eng = None
    vis = getattr(getattr(image_encoder, "clip_model", image_encoder), "visual", None)
    if vis is not None and hasattr(vis, "transformer") and hasattr(vis, "conv1"):
        width, heads = vis.conv1.weight.shape[0], vis.transformer.resblocks[0].attn.num_heads
        tokens = vis.positional_embedding.shape[0]
        if width == 64 * heads and tokens <= 256 and len(vis.transformer.resblocks) >= 3:
            eng = ClipGramEngine.from_image_encoder(image_encoder, device=device)
    image_encoder._hedit_b200_clip = eng
    return eng


def clip_gram_guidance_fused(vae_engine, clip_engine, vae_batch: int = 2) -> Callable[[torch.Tensor], torch.Tensor]:
    """The whole reward branch on native kernels: VAE decode -> CLIP-Gram loss -> CLIP backward -> VAE backward (no torch.autograd)."""

    def fn(x0: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(x0)
        for lo in range(0, x0.shape[0], vae_batch):
            hi = min(x0.shape[0], lo + vae_batch)
            img = vae_engine.decode_tensor(x0[lo:hi] * (1 / VAE_SCALE))
            clip_engine.loss(img)
            out[lo:hi] = vae_engine.backward(clip_engine.backward()) * (1 / VAE_SCALE)
        return out

    return fn


def h_edit_style_batch(model, image_encoder, xT: torch.Tensor, zs: torch.Tensor, prompt_pairs: Sequence[Sequence[str]], cfg_scales,
                       controllers, eta=1.0, weight_edit_clip=0.55, optimization_steps=1, after_skip_steps=None, is_ddim_inversion=False,
                       schedule=1, engine=None, autocast=True, guidance_fn: Optional[Callable] = None, native_vae: bool = True, native_clip: bool = True):
    """B independent text+style edits in one native call (xT (B,C,h,w) and zs (B,steps,C,h,w) on the GPU).
    native_vae = True (default): the VAE decode inside the guidance loop and its backward run on the native decoder engine;
    False: the whole reward branch is differentiated by torch.autograd through `model.vae` (compat path).
    native_clip = True (default): the CLIP-Gram loss and its image gradient also run on native kernels when `image_encoder` is a ViT
    tower the engine supports (ViT-B/16 is); otherwise `image_encoder.get_gram_matrix_residual` is differentiated by torch.autograd."""
    steps = after_skip_steps if after_skip_steps is not None else model.scheduler.num_inference_steps
    guidance = None
    if image_encoder or guidance_fn is not None:
        if guidance_fn is not None:
            fn = guidance_fn
        elif native_vae:
            dev_i = xT.device.index or 0
            clip_eng = get_clip_engine(image_encoder, dev_i) if native_clip else None
            if clip_eng is not None:
                fn = clip_gram_guidance_fused(get_vae_engine(model, dev_i), clip_eng)
            else:
                fn = clip_gram_guidance_native(get_vae_engine(model, dev_i), image_encoder)
        else:
            fn = clip_gram_guidance(model, image_encoder, autocast)
        guidance = (fn, weight_edit_clip, x0_tables(model.scheduler, steps))
    return h_edit_p2p_batch(model, xT, zs, prompt_pairs, cfg_scales, controllers, eta, 0.0, optimization_steps, steps, is_ddim_inversion,
                            False, schedule=schedule, engine=engine, mos_pull=False, guidance=guidance)


def h_Edit_p2p_implicit(model, image_encoder, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
                        weight_edit_clip=0.55, optimization_steps=1, after_skip_steps=100, is_ddim_inversion=False, autocast=True,
                        native_vae=True, native_clip=True):
    """Reference signature (text-guided-n-style/inversion/h_edit.py:14).  Returns (edited, reconstructed), each (1,C,h,w)."""
    assert len(prompts) >= 2, "only support prompt editing"
    dev = xT.device
    cdev = dev if dev.type == "cuda" else torch.device("cuda", 0)
    x = xT.reshape(1, *xT.shape[-3:]).to(cdev)
    z = zs[:after_skip_steps].reshape(1, after_skip_steps, *xT.shape[-3:]).to(cdev)
    from .compat import controller_kind
    kind = controller_kind(controller)
    if kind == "custom":
        raise NotImplementedError("the style sampler compiles the stock P2P controllers only (controller_kind() == 'custom': a user class "
                                  "with its own hooks); wrap the edit with hedit_b200.h_Edit_p2p_implicit's compat path instead")
    ctrl = [controller] if kind == "stock" else None
    edited, recon = h_edit_style_batch(model, image_encoder, x, z, [prompts[:2]], cfg_scales, ctrl, eta, weight_edit_clip, optimization_steps,
                                       after_skip_steps, is_ddim_inversion, autocast=autocast, native_vae=native_vae, native_clip=native_clip)
    return edited.to(dev), recon.to(dev)
