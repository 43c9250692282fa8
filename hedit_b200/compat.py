"""Compatibility path for ARBITRARY controller objects ("compat path").

The fast path recognises the stock Prompt-to-Prompt controllers and compiles them into device edit tables consumed by the fused attention
kernels.  Anything else -- a user subclass with its own `replace_cross_attention`, a plain callable that rescales some maps, an attention
recorder -- only promises the reference's PROTOCOL (text-guided/p2p/ptp_utils.py:96-107, ptp_classes.py:91-107):

    controller(attention_probs[(B*heads), N, M], is_cross, place_in_unet, save_attn)     # edits in place, counts layers itself
    x_t = controller.step_callback(x_t)                                                  # after every timestep

so this module runs the sampler the reference's way -- three UNet launches per step with the reference's batch composition
(p2p_h_edit.py:606-652: [xo, xe] x [null, src]; xe with src; [xo, xe, xo, xe] x [null, null, src, tar] with the controller on) -- on the
native UNet with the attention layers of the controlled launch materialising their probabilities (`hedit_unet_forward_compat`).
The controller therefore sees tensors of exactly the shape, order and content the reference would hand it.  Slow by construction."""
from __future__ import annotations

from typing import Optional

import torch

from .schedule import step_tables

PLACES = ("down", "mid", "up")
STOCK_CLASSES = ("EditController", "AttentionReplace", "AttentionRefine", "AttentionReweight")
PASSIVE_CLASSES = ("AttentionStore", "EmptyControl", "SpatialReplace")


def controller_kind(controller) -> str:
    """'none' (no P2P edit: None or a bare store), 'stock' (compiles to the fused kernels) or 'custom' (needs the compat path)."""
    if controller is None:
        return "none"
    chain, c = [], controller
    while c is not None and len(chain) < 8:
        # a stock class is recognised by name AND defining module (ours, or the reference's p2p/ptp_classes.py): a user class that merely
        # reuses a stock name, or subclasses one to override a hook, only promises the protocol -> compat path
        mod = type(c).__module__ or ""
        stock_home = mod.endswith("ptp_classes") or mod.startswith("hedit_b200")
        chain.append(type(c).__name__ if stock_home else "user:" + type(c).__name__)
        c = getattr(c, "prev_controller", None)
    if all(n in STOCK_CLASSES for n in chain) and hasattr(controller, "cross_replace_alpha"):
        return "stock"
    if len(chain) == 1 and chain[0] in PASSIVE_CLASSES and not hasattr(controller, "cross_replace_alpha"):
        return "none"
    return "custom"


class _Out:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, k):
        return self.sample if k in (0, "sample") else None


class CompatUNet:
    """`model.unet` stand-in over the native engine honouring the reference's `cross_attention_kwargs` (`use_controller`, `save_attn`:
    ptp_utils.py:38-46).  Launches without a controller use the fused kernels."""

    def __init__(self, engine, controller=None, editor=None):
        self.engine, self.controller, self.editor = engine, controller, editor
        self.in_channels = engine.config["in_channels"]
        self.sample_size = engine.config["sample_size"]

    def __call__(self, sample, timestep, encoder_hidden_states=None, cross_attention_kwargs=None, **_):
        kw = cross_attention_kwargs or {}
        use = kw.get("use_controller", True) and self.controller is not None
        save = kw.get("save_attn", True)
        t = timestep.detach().cpu().numpy() if torch.is_tensor(timestep) else timestep
        if self.editor is not None and kw.get("use_editor", True):       # MasaCtrl protocol (masactrl_utils.py:40-89)
            ed = self.editor

            def hook(_layer, is_cross, place, q, k, v, sim, attn, heads):
                return ed(q, k, v, sim, attn, is_cross, PLACES[place], heads, scale=float(q.shape[-1]) ** -0.5)
            return _Out(self.engine.forward_editor(sample, t, encoder_hidden_states, hook))
        if not use:
            return _Out(self.engine.forward(sample, t, encoder_hidden_states))
        ctrl = self.controller
        return _Out(self.engine.forward_compat(sample, t, encoder_hidden_states,
                                               lambda _layer, is_cross, place, probs: ctrl(probs, is_cross, PLACES[place], save)))

    forward = __call__


def register_attention_control_compat(model, controller):
    """ptp_utils.py:277-295 for the compat path: returns the UNet callable whose attention layers call `controller`, and records the
    layer count on it (2 per transformer block)."""
    from .samplers import get_engine
    eng = get_engine(model, max_samples=5)
    controller.num_att_layers = 2 * eng.n_transformer_blocks()
    return CompatUNet(eng, controller)


@torch.no_grad()
def h_edit_p2p_implicit_compat(model, xT, eta, prompts, cfg_scales, zs, controller, weight_reconstruction=0.075, optimization_steps=1,
                               after_skip_steps: Optional[int] = None, is_ddim_inversion=False, unet=None, editor_mode=False):
    """Implicit h-Edit + P2P for one image with a protocol-only controller (p2p_h_edit.py:529-701).  Returns (edited, reconstructed)."""
    from .samplers import encode_text
    steps = after_skip_steps if after_skip_steps is not None else model.scheduler.num_inference_steps
    unet = unet or register_attention_control_compat(model, controller)
    # `unet` is normally the CompatUNet over the native engine; any callable with the reference's `model.unet` protocol is accepted so
    # that the loop arithmetic itself can be checked against the reference sampler (tests/test_oracle_pin.py does, on a torch UNet)
    dev = torch.device("cuda", unet.engine.device) if hasattr(unet, "engine") else torch.as_tensor(xT).device
    w_src, w_src_edit, w_tar = (float(c) for c in cfg_scales)
    null = encode_text(model, [""]).float().to(dev)
    src_tar = encode_text(model, list(prompts[:2])).float().to(dev)
    src, tar = src_tar[:1], src_tar[1:2]
    ctx_a = torch.cat([null, null, src, src])
    ctx_c = torch.cat([null, null, src, tar])
    ts, coef = step_tables(model.scheduler, steps, eta, is_ddim_inversion)
    # launches A and B run without control; launch C with it: the P2P controller (`save_attn` per inner iteration) or, in editor mode,
    # MasaCtrl's editor, which the reference leaves on by passing no kwargs at all (masactrl_h_edit.py:98,124,131)
    off = {"use_editor": False} if editor_mode else {"use_controller": False}
    x = xT.reshape(1, *xT.shape[-3:]).to(dev, torch.float32)
    xt = torch.cat([x, x])
    zs = zs.to(dev, torch.float32)
    for i in range(steps):
        t, tt = ts[i], ts[i + 1]
        s1m, sa, sap, direction, noise, coeff = (float(v) for v in coef[i])
        z = zs[steps - 1 - i]
        # reverse step of both rows under the source prompt (:606-619)
        e_u, e_c = unet(torch.cat([xt, xt]), t, encoder_hidden_states=ctx_a, cross_attention_kwargs=off).sample.chunk(2)
        eps = e_u + w_src * (e_c - e_u)
        prev = sap * ((xt - s1m * eps) / sa) + direction * eps + noise * z
        x_orig, x_base = prev[:1], prev[1:]
        x_opt = x_base.clone()
        for k in range(optimization_steps):
            save = k == optimization_steps - 1                                            # :637-640
            c_src = unet(x_opt, tt, encoder_hidden_states=src, cross_attention_kwargs=off).sample                      # :644
            out = unet(torch.cat([x_orig, x_opt, x_orig, x_opt]), tt, encoder_hidden_states=ctx_c,                   # :652
                       cross_attention_kwargs=None if editor_mode else {"save_attn": save}).sample
            u_tar, c_tar = out[1:2], out[3:4]
            corr = (u_tar + w_tar * (c_tar - u_tar)) - (u_tar + w_src_edit * (c_src - u_tar))                          # :659-667
            rec = x_opt
            if k > 0 and not editor_mode:                                                                              # :670-686 (the MasaCtrl sampler has no pull)
                g = torch.sign(x_opt - x_base) / x_opt.numel()
                rho = corr.pow(2).mean().sqrt() / (g.pow(2).mean().sqrt() + 1e-8) * weight_reconstruction
                rec = x_opt - rho * g
            x_opt = rec + coeff * corr                                                                                 # :689-692
        xt = torch.cat([x_orig, x_opt])
        if controller is not None and hasattr(controller, "step_callback"):
            xt = controller.step_callback(xt)                                                                          # :698-699
    return xt[1:2].clone(), xt[0:1].clone()


def register_attention_editor_compat(model, editor):
    """masactrl_utils.py:35-107 for the compat path: the UNet callable whose attention layers hand q, k, v, sim, attn to `editor`."""
    from .samplers import get_engine
    eng = get_engine(model, max_samples=5)
    editor.num_att_layers = 2 * eng.n_transformer_blocks()
    return CompatUNet(eng, editor=editor)


def h_edit_masactrl_implicit_compat(model, xT, eta, prompts, cfg_scales, zs, editor, optimization_steps=1, after_skip_steps=None,
                                    is_ddim_inversion=True, unet=None):
    """Implicit h-Edit with an ARBITRARY MasaCtrl-protocol editor object (masactrl_h_edit.py:14-160): the same three launches per step as
    the P2P sampler, the third with the editor on, no reconstruction pull, no step callback."""
    unet = unet or register_attention_editor_compat(model, editor)
    return h_edit_p2p_implicit_compat(model, xT, eta, prompts, cfg_scales, zs, None, 0.0, optimization_steps, after_skip_steps,
                                      is_ddim_inversion, unet=unet, editor_mode=True)
