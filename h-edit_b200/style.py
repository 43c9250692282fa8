"""Combined text-guided + style editing: drop-in for the reference's text-guided-n-style/inversion/h_edit.py.

The text-guided part of every step (two classifier-free-guided UNet calls with P2P injection, h-term move) runs in the native
batched loop; after each implicit-loop iteration the loop hands the Tweedie prediction x0 to a reward hook and applies the
Langevin move x <- x - rho * dLoss/dx with fused element kernels (csrc/hstep.cuh: hstep_x0pred / guid_norm / guid_update).
The reward gradient itself is evaluated through the reference's reward-model protocol: `model.vae.decode(z).sample`
and `image_encoder.get_gram_matrix_residual(img)` are caller-supplied torch modules differentiated with torch.autograd on the
same CUDA stream.  (A hand-written VAE-decoder / CLIP forward+backward is the next step for this path; see DESIGN.md.)"""
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


def h_edit_style_batch(model, image_encoder, xT: torch.Tensor, zs: torch.Tensor, prompt_pairs: Sequence[Sequence[str]], cfg_scales,
                       controllers, eta=1.0, weight_edit_clip=0.55, optimization_steps=1, after_skip_steps=None, is_ddim_inversion=False,
                       schedule=1, engine=None, autocast=True, guidance_fn: Optional[Callable] = None):
    """B independent text+style edits in one native call (xT (B,C,h,w) and zs (B,steps,C,h,w) on the GPU)."""
    steps = after_skip_steps if after_skip_steps is not None else model.scheduler.num_inference_steps
    guidance = None
    if image_encoder or guidance_fn is not None:
        fn = guidance_fn if guidance_fn is not None else clip_gram_guidance(model, image_encoder, autocast)
        guidance = (fn, weight_edit_clip, x0_tables(model.scheduler, steps))
    return h_edit_p2p_batch(model, xT, zs, prompt_pairs, cfg_scales, controllers, eta, 0.0, optimization_steps, steps, is_ddim_inversion,
                            False, schedule=schedule, engine=engine, mos_pull=False, guidance=guidance)


def h_Edit_p2p_implicit(model, image_encoder, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
                        weight_edit_clip=0.55, optimization_steps=1, after_skip_steps=100, is_ddim_inversion=False, autocast=True):
    """Reference signature (text-guided-n-style/inversion/h_edit.py:14).  Returns (edited, reconstructed), each (1,C,h,w)."""
    assert len(prompts) >= 2, "only support prompt editing"
    dev = xT.device
    x = xT.reshape(1, *xT.shape[-3:]).cuda()
    z = zs[:after_skip_steps].reshape(1, after_skip_steps, *xT.shape[-3:]).cuda()
    ctrl = [controller] if (controller is not None and hasattr(controller, "cross_replace_alpha")) else None
    edited, recon = h_edit_style_batch(model, image_encoder, x, z, [prompts[:2]], cfg_scales, ctrl, eta, weight_edit_clip, optimization_steps,
                                       after_skip_steps, is_ddim_inversion, autocast=autocast)
    return edited.to(dev), recon.to(dev)
