"""Drop-in sampler callables: same names and argument meaning as the reference's
text-guided/inversion/p2p_h_edit.py (h_Edit_p2p_implicit :529, h_Edit_p2p_explicit :380), backed by the native
batched loop.  `model` is the reference's duck-typed StableDiffusionPipeline (unet, scheduler, tokenizer,
text_encoder, device)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from .engine import UNetEngine
from .p2p import compile_edit_plan
from .schedule import skip_pre_coeff, step_tables


def get_text_engine(model, device: int = 0):
    """Native CLIP text tower built once from `model.text_encoder` when that is a transformers CLIPTextModel with quick_gelu (SD-1.x);
    None otherwise (any other text-encoder object is simply called, as the reference does)."""
    if hasattr(model, "_hedit_b200_text"):
        return model._hedit_b200_text
    eng = None
    te = getattr(model, "text_encoder", None)
    cfg = getattr(te, "config", None)
    if te is not None and hasattr(te, "text_model") and getattr(cfg, "hidden_act", None) == "quick_gelu" and torch.cuda.is_available() and \
            cfg.hidden_size == 64 * cfg.num_attention_heads:
        from .text_encoder import TextEncoderEngine
        eng = TextEncoderEngine.from_text_encoder(te, device=device)
    model._hedit_b200_text = eng
    return eng


def encode_text(model, prompts: Union[str, List[str]]) -> torch.Tensor:
    """text-guided/inversion/inversion_utils.py:13-36 (tokenise to max_length, run the text encoder)."""
    tok = model.tokenizer(prompts, padding="max_length", max_length=model.tokenizer.model_max_length, truncation=True,
                          return_tensors="pt")
    eng = get_text_engine(model)
    if eng is not None:
        return eng(tok.input_ids)[0]
    with torch.no_grad():
        return model.text_encoder(tok.input_ids.to(model.device))[0]


def encode_prompts(model, prompts: List[str]) -> torch.Tensor:
    """encode_text for a whole batch in ONE tokenizer call and ONE text-tower launch: identical strings (the unconditional "" of every
    image, repeated source prompts) are encoded once.  Returns (len(prompts), 77, D) in prompt order."""
    uniq = list(dict.fromkeys(prompts))
    emb = encode_text(model, uniq)
    if len(uniq) == len(prompts):
        return emb
    pos = {p: k for k, p in enumerate(uniq)}
    return emb[torch.as_tensor([pos[p] for p in prompts], device=emb.device)]


def _device_index(t: torch.Tensor, default: int = 0) -> int:
    return t.device.index if (t.is_cuda and t.device.index is not None) else default


def get_engine(model, max_samples: int = 5, device: Optional[int] = None) -> UNetEngine:
    """One persistent engine per pipeline object (replaces the per-image deepcopy at main_p2p.py:119)."""
    eng = getattr(model, "_hedit_b200_engine", None)
    if device is None:                      # keep the engine's device; a first engine goes where the pipeline lives
        mdev = getattr(model, "device", None)
        device = eng.device if eng is not None else (torch.device(mdev).index or 0 if mdev is not None and torch.device(mdev).type == "cuda" else 0)
    key = _weights_fingerprint(model.unet)
    if eng is None or eng.max_samples < max_samples or eng.device != device or getattr(model, "_hedit_b200_engine_key", key) != key:
        eng = UNetEngine.from_unet(model.unet, max_samples=max_samples, max_contexts=max(8, 1 + 2 * (max_samples // 5 + 1)), device=device)
        model._hedit_b200_engine = eng
    model._hedit_b200_engine_key = key
    return eng


def _weights_fingerprint(unet):
    """Cheap identity of the weights an engine was built from: the module object plus the in-place version counters of its parameters
    (load_state_dict, LoRA merges and optimizer steps bump them), so a changed `model.unet` is re-ingested instead of silently ignored
    (the reference always runs the live module).  `invalidate_engines(model)` forces it."""
    params = getattr(unet, "parameters", None)
    if params is None:
        return (id(unet),)
    try:
        return (id(unet),) + tuple(int(p._version) for p in params())
    except Exception:
        return (id(unet),)


def invalidate_engines(model) -> None:
    """Drop the native engines cached on a pipeline object (after swapping or editing `model.unet`, `text_encoder` or `vae`)."""
    for k in ("_hedit_b200_engine", "_hedit_b200_engine_key", "_hedit_b200_text", "_hedit_b200_vae", "_hedit_b200_face"):
        if hasattr(model, k):
            delattr(model, k)


def h_edit_p2p_batch(model, xT: torch.Tensor, zs: torch.Tensor, prompt_pairs: Sequence[Sequence[str]], cfg_scales, controllers,
                     eta=1.0, weight_reconstruction=0.075, optimization_steps=1, after_skip_steps=None, is_ddim_inversion=False,
                     explicit_form=False, schedule=1, engine: Optional[UNetEngine] = None, trace=False, variant=0, masactrl=None,
                     mos_pull=True, pnp=None, pre_coeff=None, guidance=None, coef_edit=None, null_prompts=None):
    """B independent edits in one native call.  xT (B,C,h,w); zs (B,steps,C,h,w); prompt_pairs[b] = [src, tar];
    controllers[b] = P2P controller of image b (ours or the reference's) or None for all (P2P off)."""
    B = xT.shape[0]
    steps = after_skip_steps if after_skip_steps is not None else model.scheduler.num_inference_steps
    eng = engine or get_engine(model, max_samples=5 * B, device=_device_index(xT, None) if xT.is_cuda else None)
    if engine is None:
        # one image per call (the reference's signature): 2-5 samples per launch leave most SMs idle in the deep levels -> split-K.  Batches
        # keep it off so that an image's result does not depend on what it is batched with.
        eng.set_splitk(B == 1)
    # context 0 is the unconditional one: "" -- or, for Negative-Prompt inversion, whatever the caller substitutes for it
    flat = [null_prompts if isinstance(null_prompts, str) else ""]
    for src, tar in prompt_pairs:
        flat += [src, tar]
    ctx = encode_prompts(model, flat).float()        # stays where the text tower left it (the C ABI accepts a context pointer on either side)
    if xT.is_cuda:
        ctx = ctx.to(xT.device)
    ts, coef = step_tables(model.scheduler, steps, eta, is_ddim_inversion)
    plan = None
    if controllers is not None:
        from .compat import controller_kind
        kinds = [controller_kind(c) for c in controllers]
        if any(k == "custom" for k in kinds):
            raise NotImplementedError("h_edit_p2p_batch compiles the stock P2P controllers only; a controller object with its own hooks "
                                      "(controller_kind() == 'custom') is served per image by h_Edit_p2p_implicit (compat path)")
        if any(k == "stock" for k in kinds):
            if not all(k == "stock" for k in kinds):
                raise ValueError("controllers mixes P2P controllers with None / passive stores: split the batch")
            plan = compile_edit_plan(controllers, steps)
    out = eng.edit(xT, zs[:, :steps], ctx, ts, coef, cfg_scales, plan, weight_reconstruction, optimization_steps, explicit_form, schedule, trace,
                   variant=variant, masactrl=masactrl, mos_pull=mos_pull, pnp=pnp, pre_coeff=pre_coeff, guidance=guidance, coef_edit=coef_edit)
    if plan is not None:
        for c in controllers:       # keep the controller's observable counters consistent with the reference
            c.cur_step = getattr(c, "cur_step", 0) + steps
            if getattr(c, "local_blend", None) is not None:
                c.local_blend.counter += steps
    return out


def _single(model, xT, eta, prompts, cfg_scales, zs, controller, weight_reconstruction, optimization_steps, after_skip_steps,
            is_ddim_inversion, explicit_form, variant=0, masactrl=None, mos_pull=True, pnp=None, pre_coeff=None):
    assert len(prompts) >= 2, "only support prompt editing"
    dev = xT.device
    x = xT.reshape(1, *xT.shape[-3:])
    steps = after_skip_steps
    z = zs[:steps].reshape(1, steps, *xT.shape[-3:])
    # hand a stock P2P controller to the fused path; a bare AttentionStore (no-P2P modes) carries no edit tables; any other object only
    # promises the reference's call protocol and is served by the compat path (materialised probabilities, compat.py)
    from .compat import controller_kind, h_edit_p2p_implicit_compat
    kind = controller_kind(controller)
    if kind == "none" and controller is not None and variant == 0 and masactrl is None and pnp is None and not explicit_form:
        # a bare AttentionStore handed to the P2P sampler: the reference would fill its attention_store (ptp_classes.py:135-160); the fused
        # path never materialises maps, so it is served by the compat path like any other object with observable state
        kind = "custom"
    if kind == "custom":
        if explicit_form or variant != 0 or masactrl is not None or pnp is not None:
            raise NotImplementedError("custom controller objects are served on the implicit h-Edit + P2P sampler only (compat path)")
        edited, recon = h_edit_p2p_implicit_compat(model, xT, eta, prompts, cfg_scales, zs, controller, weight_reconstruction,
                                                   optimization_steps, after_skip_steps, is_ddim_inversion)
        return edited.to(dev), recon.to(dev)
    ctrl = [controller] if kind == "stock" else None
    use_cuda = torch.device(dev).type == "cuda"
    x, z = (x, z.to(dev)) if use_cuda else (x.cpu(), z.cpu())
    edited, recon = h_edit_p2p_batch(model, x, z, [prompts[:2]], cfg_scales, ctrl, eta, weight_reconstruction, optimization_steps, steps,
                                     is_ddim_inversion, explicit_form, variant=variant, masactrl=masactrl, mos_pull=mos_pull, pnp=pnp, pre_coeff=pre_coeff)
    if kind == "stock" and not explicit_form:
        # observable controller state (SURVEY 8b ii): the maps the reference's AttentionStore would hold are materialised on first access
        from .p2p import attach_lazy_store
        xT_c, zs_c, prompts_c, cfgs_c = xT.detach().clone(), zs[:steps].detach().clone(), list(prompts[:2]), list(cfg_scales)
        attach_lazy_store(controller, lambda twin: h_edit_p2p_implicit_compat(model, xT_c, eta, prompts_c, cfgs_c, zs_c, twin, weight_reconstruction,
                                                                              optimization_steps, steps, is_ddim_inversion))
    return edited.to(dev), recon.to(dev)


def h_Edit_p2p_implicit(model, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
                        weight_reconstruction=0.075, optimization_steps=1, after_skip_steps=35, is_ddim_inversion=True):
    """Reference signature (p2p_h_edit.py:529).  Returns (edited, reconstructed), each (1,C,h,w)."""
    return _single(model, xT, eta, prompts, cfg_scales, zs, controller, weight_reconstruction, optimization_steps, after_skip_steps,
                   is_ddim_inversion, False)


def h_Edit_p2p_explicit(model, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
                        is_ddim_inversion=True, after_skip_steps=35):
    """Reference signature (p2p_h_edit.py:380)."""
    return _single(model, xT, eta, prompts, cfg_scales, zs, controller, 0.0, 1, after_skip_steps, is_ddim_inversion, True)


def h_Edit_R_implicit(model, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
                      weight_reconstruction=0.1, optimization_steps=1, after_skip_steps=35, is_ddim_inversion=False):
    """Reference signature (p2p_h_edit.py:162): h-Edit-R, implicit form, no P2P; on a skipped schedule the edit row is first moved
    once at the first executed timestep (:239-267)."""
    assert not is_ddim_inversion, "only DDPM sampling (reference assert, p2p_h_edit.py:196)"
    return _single(model, xT, eta, prompts, cfg_scales, zs, None, weight_reconstruction, optimization_steps, after_skip_steps,
                   is_ddim_inversion, False, variant=1, pre_coeff=skip_pre_coeff(model.scheduler, after_skip_steps, eta, is_ddim_inversion))


def h_Edit_R_explicit(model, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
                      is_ddim_inversion=False, after_skip_steps=35):
    """Reference signature (p2p_h_edit.py:21): h-Edit-R, explicit form, no P2P."""
    return _single(model, xT, eta, prompts, cfg_scales, zs, None, 0.0, 1, after_skip_steps, is_ddim_inversion, True, variant=1)


class MutualSelfAttentionControl:
    """MasaCtrl's schedule (masactrl/masactrl.py:11-36): mutual self-attention at the editor steps of `step_idx` (default
    range(start_step, total_steps): the injection ENDS once cur_step reaches total_steps) in the transformer blocks of `layer_idx`
    (default range(start_layer, 16)).  The control itself runs inside the self-attention kernel as a K/V source-sample swap; the
    editor's cur_step advances once per attention-controlled UNet launch (masactrl_utils.py:15-23)."""
    MODEL_TYPE = {"SD": 16, "SDXL": 70}

    def __init__(self, start_step=4, start_layer=10, layer_idx=None, step_idx=None, total_steps=50, model_type="SD"):
        self.total_steps = total_steps
        self.total_layers = self.MODEL_TYPE.get(model_type, 16)
        self.start_step, self.start_layer = start_step, start_layer
        self.layer_idx = list(layer_idx) if layer_idx is not None else list(range(start_layer, self.total_layers))
        self.step_idx = list(step_idx) if step_idx is not None else list(range(start_step, total_steps))
        self.cur_step = 0
        self.cur_att_layer = 0
        self.num_att_layers = -1

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    def launch_plan(self, n_launches: int, n_blocks: int = 16):
        """(layer bit mask, per-launch on/off flags) for the next `n_launches` attention-controlled UNet launches."""
        mask = 0
        for l in self.layer_idx:
            if 0 <= int(l) < min(n_blocks, 32):
                mask |= 1 << int(l)
        steps = set(int(v) for v in self.step_idx)
        return mask, [int(self.cur_step + c in steps) for c in range(n_launches)]


def regiter_attention_editor_diffusers(model, editor) -> None:
    """Same (misspelt) name as the reference's registration hook (masactrl/masactrl_utils.py:35): records the editor on the
    pipeline object; no per-layer monkey-patching is needed."""
    model._hedit_masactrl_editor = editor
    editor.num_att_layers = 32


def h_Edit_masactrl_implicit(model, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, optimization_steps=1,
                             after_skip_steps=35, is_ddim_inversion=True):
    """Reference signature (masactrl_h_edit.py:14)."""
    ed = getattr(model, "_hedit_masactrl_editor", None)
    assert ed is not None, "call regiter_attention_editor_diffusers(model, MutualSelfAttentionControl(...)) first"
    if type(ed).__name__ != "MutualSelfAttentionControl":
        # an editor object of the user's own class only promises the call protocol: compat path (materialised q, k, v, sim, attn)
        from .compat import h_edit_masactrl_implicit_compat
        edited, recon = h_edit_masactrl_implicit_compat(model, xT, eta, prompts, cfg_scales, zs, ed, optimization_steps, after_skip_steps,
                                                        is_ddim_inversion)
        return edited.to(xT.device), recon.to(xT.device)
    n_launches = after_skip_steps * max(1, optimization_steps)
    out = _single(model, xT, eta, prompts, cfg_scales, zs, None, 0.0, optimization_steps, after_skip_steps, is_ddim_inversion, False,
                  masactrl=ed.launch_plan(n_launches, get_engine(model).n_transformer_blocks()), mos_pull=False)
    ed.cur_step += n_launches
    return out


def h_Edit_masactrl_explicit(model, xT, eta=1.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, after_skip_steps=35,
                             is_ddim_inversion=True):
    """Explicit-form h-Edit with MasaCtrl (BASELINE.json configs[2]).  The reference ships NO function of this form (only
    `h_Edit_masactrl_implicit`); this composes the update of `h_Edit_p2p_explicit` (p2p_h_edit.py:480-514) with the MasaCtrl editor
    active in its attention-controlled call, as SURVEY 8d describes.  One 5-sample UNet launch per step."""
    ed = getattr(model, "_hedit_masactrl_editor", None)
    assert ed is not None, "call regiter_attention_editor_diffusers(model, MutualSelfAttentionControl(...)) first"
    out = _single(model, xT, eta, prompts, cfg_scales, zs, None, 0.0, 1, after_skip_steps, is_ddim_inversion, True,
                  masactrl=ed.launch_plan(after_skip_steps, get_engine(model).n_transformer_blocks()), mos_pull=False)
    ed.cur_step += after_skip_steps
    return out


# ---- Plug-and-Play (text-guided/plug_n_play/pnp_utils.py, inversion/pnp_h_edit.py) ---------------------------------------
PNP_ATTN_BLOCKS = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}     # pnp_utils.py:88: decoder self-attention layers 4-11


def register_time(model, t) -> None:
    """pnp_utils.py:12 / pnp_h_edit.py:10.  The fused loop derives the injection flags of every step from the schedules recorded by
    the two register_* functions below, so this only keeps the attribute observable."""
    model._hedit_pnp_t = t


def register_attention_control_efficient(model, injection_schedule) -> None:
    """Same name as pnp_utils.py:29: records the timesteps at which the target's self-attention q and k are replaced by the source's
    in `PNP_ATTN_BLOCKS` (the swap itself is a source-sample index inside self_attn_kernel)."""
    model._hedit_pnp_qk = None if injection_schedule is None else {int(t) for t in injection_schedule}


def register_conv_control_efficient(model, injection_schedule) -> None:
    """Same name as pnp_utils.py:97: records the timesteps at which the target takes the source's conv2 output in
    up_blocks[1].resnets[1]."""
    model._hedit_pnp_conv = None if injection_schedule is None else {int(t) for t in injection_schedule}


def pnp_self_mask(layers_per_block: int = 2) -> int:
    """Bit mask over transformer blocks in forward order of the layers of `PNP_ATTN_BLOCKS`."""
    L, m = layers_per_block, 0
    for res, blocks in PNP_ATTN_BLOCKS.items():
        for b in blocks:
            m |= 1 << (3 * L + 1 + (res - 1) * (L + 1) + b)
    return m


def pnp_step_flags(model, after_skip_steps: int):
    """(qk_on, feat_on) per executed timestep: the injected pair call of step i runs at tt = op[i+1] (0 after the last step) and the
    patched forwards test `self.t in injection_schedule or self.t == 1000` (pnp_utils.py:51-52,138)."""
    op = [int(t) for t in model.scheduler.timesteps[-after_skip_steps:]]
    tts = op[1:] + [0]
    on = lambda sched: [int(sched is not None and (tt in sched or tt == 1000)) for tt in tts]
    return on(getattr(model, "_hedit_pnp_qk", None)), on(getattr(model, "_hedit_pnp_conv", None))


def h_Edit_PnP_implicit(model, xT, eta=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, optimization_steps=1, after_skip_steps=35,
                        is_ddim_inversion=True, schedule=1):
    """Reference signature (pnp_h_edit.py:33); `schedule` = 0 reproduces the reference's 8 UNet sample-forwards per step, 1 (default)
    the exact-reuse 7."""
    qk_on, feat_on = pnp_step_flags(model, after_skip_steps)
    L = getattr(getattr(model.unet, "cfg", None), "layers_per_block", 2)
    B = 1
    dev = xT.device
    eng = get_engine(model, max_samples=5 * B, device=_device_index(xT, None) if xT.is_cuda else None)
    eng.set_splitk(True)       # one image per call: see h_edit_p2p_batch
    x = xT.reshape(1, *xT.shape[-3:])
    z = zs[:after_skip_steps].reshape(1, after_skip_steps, *xT.shape[-3:])
    use_cuda = torch.device(dev).type == "cuda"
    x, z = (x, z.to(dev)) if use_cuda else (x.cpu(), z.cpu())
    edited, recon = h_edit_p2p_batch(model, x, z, [prompts[:2]], cfg_scales, None, eta, 0.0, optimization_steps, after_skip_steps,
                                     is_ddim_inversion, False, schedule=schedule, engine=eng, mos_pull=False,
                                     pnp=(pnp_self_mask(L), qk_on, feat_on))
    return edited.to(dev), recon.to(dev)


class HEditStepper:
    """Per-timestep interface to the same native loop (`h_edit_step`): the body of the reference's `for i, t in enumerate(op)`
    (p2p_h_edit.py:598-699) as one call, with the controller state (P2P step counter, LocalBlend word maps) carried here.
    Uses the reference's UNet call pattern (schedule 0), since the exact-reuse schedule carries UNet outputs across steps."""

    def __init__(self, model, prompt_pairs, cfg_scales, controllers=None, eta=1.0, weight_reconstruction=0.075, optimization_steps=1,
                 after_skip_steps=None, is_ddim_inversion=False, explicit_form=False, engine: Optional[UNetEngine] = None):
        self.model = model
        self.B = len(prompt_pairs)
        self.steps = after_skip_steps if after_skip_steps is not None else model.scheduler.num_inference_steps
        self.eng = engine or get_engine(model, max_samples=5 * self.B)
        dev = torch.device("cuda", self.eng.device)
        ctx = [encode_text(model, [""])] + [encode_text(model, list(pp)) for pp in prompt_pairs]
        self.ctx = torch.cat(ctx).float().to(dev)
        self.ts, self.coef = step_tables(model.scheduler, self.steps, eta, is_ddim_inversion)
        self.plan = compile_edit_plan(controllers, self.steps) if controllers is not None else None
        self.cfg_scales, self.w_rec, self.K, self.explicit = cfg_scales, weight_reconstruction, optimization_steps, explicit_form
        self.blend_state = self.eng.new_blend_state(self.B) if self.plan is not None and self.plan.has_blend.any() else None
        self.i = 0

    def step(self, xt: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """xt (B,2,C,h,w) = rows (x_orig, x_edit) at the current timestep; z (B,C,h,w) = zs[:, idx] of this step.
        Returns xt at the previous timestep."""
        import dataclasses
        i = self.i
        plan = None
        if self.plan is not None:
            plan = dataclasses.replace(self.plan, steps=1, c_base=self.plan.c_base[i:i + 2], c_tar=self.plan.c_tar[i:i + 2])
        dev = self.ctx.device
        ed, rc = self.eng.edit(xt.to(dev).contiguous(), z.to(dev)[:, None].contiguous(), self.ctx, self.ts[i:i + 2], self.coef[i:i + 1],
                               self.cfg_scales, plan, self.w_rec, self.K, self.explicit, schedule=0, xt_is_pair=True, ctrl_step0=i,
                               blend_state=self.blend_state)
        self.i += 1
        return torch.stack([rc, ed], dim=1)


def h_edit_step(stepper: HEditStepper, xt: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """One reverse-time bridge step [x_orig_t, x_edit_t] -> [x_orig_{t-1}, x_edit_{t-1}] (incl. P2P injection and LocalBlend)."""
    return stepper.step(xt, z)


# ---- baseline samplers of the reference's drivers (main_p2p.py --mode ef / ef_p2p / pnp_inv_p2p, main_masactrl.py) ----------------
def _baseline(model, xT, etas, prompts, cfg_scales, zs, controller, is_ddim_inversion, masactrl=None, pnp=None, null_prompt=None):
    """One native call of loop variant 2 (csrc/edit_loop.cu): per timestep ONE attention-controlled launch [xo,null] [xe,null] [xo,src]
    [xe,tar], the orig row stepped with the source-guided noise and the edit row with the target-guided noise."""
    assert len(prompts) >= 2 and len(cfg_scales) >= 2
    dev = xT.device
    steps = zs.shape[0]
    x = xT.reshape(1, *xT.shape[-3:])
    z = zs.reshape(1, steps, *xT.shape[-3:])
    from .compat import controller_kind
    kind = controller_kind(controller)
    if kind == "custom":
        raise NotImplementedError("the baseline samplers run stock P2P controllers (or none) on the fused path")
    use_cuda = torch.device(dev).type == "cuda"
    x, z = (x, z.to(dev)) if use_cuda else (x.cpu(), z.cpu())
    # PnP Inversion steps the edit row deterministically (eta = 0) while the orig row keeps eta (p2p_baselines.py:176-184)
    coef_edit = step_tables(model.scheduler, steps, 0.0, True)[1] if is_ddim_inversion else None
    w_src, w_tar = float(cfg_scales[0]), float(cfg_scales[1])
    edited, recon = h_edit_p2p_batch(model, x, z, [prompts[:2]], [w_src, w_src, w_tar], [controller] if kind == "stock" else None, etas, 0.0, 1, steps,
                                     is_ddim_inversion, False, variant=2, masactrl=masactrl, mos_pull=False, coef_edit=coef_edit, pnp=pnp,
                                     null_prompts=null_prompt)
    return edited.to(dev), recon.to(dev)


def ef_or_pnp_inv_w_p2p(model, xT, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None, is_ddim_inversion=False):
    """Reference signature (inversion/p2p_baselines.py:103): Edit Friendly (is_ddim_inversion=False) / PnP Inversion (True) with P2P.
    Returns (edited, reconstructed), each (1,C,h,w)."""
    return _baseline(model, xT, etas, prompts, cfg_scales, zs, controller, is_ddim_inversion)


def ef_wo_p2p(model, xT, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None, is_ddim_inversion=False):
    """Reference signature (inversion/p2p_baselines.py:19): Edit Friendly without P2P -- only the target prompt is denoised from xT with the
    inverted noise maps.  Returns the edited sample (1,C,h,w) like the reference (its `controller` only ever stores maps)."""
    assert len(prompts) == 1 and len(cfg_scales) == 1, "ef_wo_p2p takes the target prompt only (p2p_baselines.py:35)"
    edited, _ = _baseline(model, xT, etas, [prompts[0], prompts[0]], [cfg_scales[0], cfg_scales[0]], zs, None, is_ddim_inversion)
    return edited


def ef_or_pnp_inv_w_masactrl(model, xT, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, is_ddim_inversion=False):
    """Reference signature (inversion/masactrl_baselines.py:15): Edit Friendly / PnP Inversion with the registered MasaCtrl editor."""
    ed = getattr(model, "_hedit_masactrl_editor", None)
    assert ed is not None, "call regiter_attention_editor_diffusers(model, MutualSelfAttentionControl(...)) first"
    if type(ed).__name__ != "MutualSelfAttentionControl":
        raise NotImplementedError("the baseline samplers run the stock MutualSelfAttentionControl editor on the fused path")
    steps = zs.shape[0]
    out = _baseline(model, xT, etas, prompts, cfg_scales, zs, None, is_ddim_inversion,
                    masactrl=ed.launch_plan(steps, get_engine(model).n_transformer_blocks()))
    ed.cur_step += steps
    return out


def pnp_step_flags_at_t(model, after_skip_steps: int):
    """(qk_on, feat_on) per executed timestep for samplers whose injected pair call runs at the CURRENT timestep t (pnp_baselines.py:367
    register_time(model, t)): `self.t in injection_schedule or self.t == 1000` (pnp_utils.py:43-44,132)."""
    op = [int(t) for t in model.scheduler.timesteps[-after_skip_steps:]]
    on = lambda sched: [int(sched is not None and (t in sched or t == 1000)) for t in op]
    return on(getattr(model, "_hedit_pnp_qk", None)), on(getattr(model, "_hedit_pnp_conv", None))


def _pnp_tuple(model, steps):
    qk_on, feat_on = pnp_step_flags_at_t(model, steps)
    return (pnp_self_mask(getattr(getattr(model.unet, "cfg", None), "layers_per_block", 2)), qk_on, feat_on)


def ef_or_pnp_inv_w_pnp(model, xT, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, is_ddim_inversion=False):
    """Reference signature (inversion/pnp_baselines.py:317): Edit Friendly / PnP Inversion with Plug-and-Play injection (register the
    injection schedules with register_attention_control_efficient / register_conv_control_efficient first).  The reference's two
    single-sample unconditional calls and its injected pair call are one 4-sample launch here (injection only ever writes the pair's
    target sample)."""
    assert len(prompts) >= 2 and etas == 0, "PnP requires source and target prompts, with eta is set to 0"      # reference assert (:340)
    return _baseline(model, xT, etas, prompts, cfg_scales, zs, None, is_ddim_inversion, pnp=_pnp_tuple(model, zs.shape[0]))


def negative_prompt_pnp(model, xT, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None):
    """Reference signature (inversion/pnp_baselines.py:244): Negative-Prompt inversion with Plug-and-Play -- the unconditional embedding is
    replaced by the SOURCE prompt's, both rows use the target guidance scale (:290-291) and step deterministically (eta = 0)."""
    assert len(prompts) >= 2 and etas == 0, "PnP requires source and target prompts, with eta is set to 0"      # reference assert (:263)
    return _baseline(model, xT, 0, prompts, [cfg_scales[1], cfg_scales[1]], zs, None, False, pnp=_pnp_tuple(model, zs.shape[0]), null_prompt=prompts[0])


def _unet_sample(unet, x, t, emb):
    """One forward of the CALLER's differentiable torch UNet (diffusers call convention)."""
    try:
        return unet(x, t, encoder_hidden_states=emb, cross_attention_kwargs={"use_controller": False}).sample
    except TypeError:                                             # a UNet with stock attention processors takes no P2P keyword
        return unet(x, t, encoder_hidden_states=emb).sample


def _grad_guided(model, xT, xT_ori, prompts, cfg_scales, zs, controller, pnp_on, mode, guidance_noise_map=10.0, grad_scale=5e+3,
                 optimization_steps=10, epsilon=1e-5):
    """Shared body of the three reference samplers that differentiate THROUGH the UNet once (or a few times) per timestep
    (p2p_baselines.py:195 nmg_p2p, pnp_baselines.py:32 nmg_pnp, :134 nulltext_pnp).  HYBRID: there is no native UNet backward yet
    (DESIGN 10.1), so those differentiable forwards run on the caller's own `model.unet` torch module under autograd, exactly as the
    reference does; everything else -- the attention-controlled / feature-injected 4-sample launch, both reverse steps, LocalBlend -- is a
    single-step call of the native loop (variant 2, controller state carried across calls).
      mode "nmg":      the recon row is first moved by noise-map guidance (gradient of an L1 loss w.r.t. the latent);
      mode "nulltext": the unconditional embedding is optimised with Adam for this step (gradient w.r.t. the embedding) and replaces
                       context 0 of the native launch."""
    import dataclasses
    import torch.nn.functional as F
    from .compat import controller_kind
    kind = controller_kind(controller)
    if kind == "custom":
        raise NotImplementedError("the baseline samplers run stock P2P controllers (or none) on the fused path")
    steps = zs.shape[0]
    eng = get_engine(model, max_samples=5, device=_device_index(xT, None) if xT.is_cuda else None)
    eng.set_splitk(True)
    dev = torch.device("cuda", eng.device)
    ctx = encode_prompts(model, ["", prompts[0], prompts[1]]).float().to(dev)
    ts, coef = step_tables(model.scheduler, steps, 0.0, False)          # reverse_step(..., eta = 0.0, variance_noise = None) everywhere
    plan = compile_edit_plan([controller], steps) if kind == "stock" else None
    blend_state = eng.new_blend_state(1) if plan is not None and plan.has_blend.any() else None
    pnp_mask, (qk_on, feat_on) = (pnp_self_mask(getattr(getattr(model.unet, "cfg", None), "layers_per_block", 2)), pnp_step_flags_at_t(model, steps)) \
        if pnp_on else (0, (None, None))
    unet = model.unet
    udev = next(unet.parameters()).device
    uncond = encode_text(model, [""]).to(udev).float()
    cond_src = encode_text(model, [prompts[0]]).to(udev).float()
    ac = model.scheduler.alphas_cumprod
    w_tar = float(cfg_scales[1])                                          # both rows use the TARGET scale (p2p_baselines.py:243-244)
    x = xT.reshape(1, *xT.shape[-3:]).to(dev, torch.float32)
    xt = torch.stack([x, x], dim=1).contiguous()                          # (1, 2, C, h, w): rows (recon, target)
    zero = torch.zeros_like(x)[:, None]

    def reverse0(eps, sample, c):                                         # reverse_step with eta = 0 from the step's scalar row
        return c[2] * ((sample - c[0] * eps) / c[1]) + c[3] * eps

    for i in range(steps):
        t = int(ts[i])
        c = [float(v) for v in coef[i]]
        x_ori = xT_ori[len(xT_ori) - i - 2].reshape(1, *xT.shape[-3:]).to(udev, torch.float32)
        step_ctx = ctx
        if mode == "nmg":
            with torch.enable_grad():
                x_in = xt[:, 0].detach().to(udev).requires_grad_(True)
                eps_u = _unet_sample(unet, x_in, t, uncond)
                loss = F.l1_loss(reverse0(eps_u, x_in, c), x_ori)
                grad = -torch.autograd.grad(loss, x_in)[0]
            eps_u = eps_u.detach()
            eps_c = eps_u - (1 - ac[t]).sqrt().to(udev) * grad * grad_scale
            eps = eps_u + guidance_noise_map * (eps_c - eps_u)
            xt[:, 0] = reverse0(eps, xt[:, 0].to(udev), c).to(dev)
        else:
            x_rec = xt[:, 0].detach().to(udev)
            with torch.no_grad():
                eps_cond = _unet_sample(unet, x_rec, t, cond_src)
            with torch.enable_grad():
                emb = uncond.detach().clone().requires_grad_(True)
                opt = torch.optim.Adam([emb], lr=1e-2 * (1. - i / 100.))
                for _ in range(optimization_steps):
                    eps_u = _unet_sample(unet, x_rec, t, emb)
                    loss = F.mse_loss(reverse0(eps_u + w_tar * (eps_cond - eps_u), x_rec, c), x_ori)
                    opt.zero_grad()
                    loss.backward()
                    opt.step()
                    if loss.item() < epsilon + i * 2e-5:
                        break
            step_ctx = torch.cat([emb.detach().to(dev), ctx[1:]])
        p = None if plan is None else dataclasses.replace(plan, steps=1, c_base=plan.c_base[i:i + 2], c_tar=plan.c_tar[i:i + 2])
        pnp = (pnp_mask, qk_on[i:i + 1], feat_on[i:i + 1]) if pnp_on else None
        ed, rc = eng.edit(xt.contiguous(), zero, step_ctx, ts[i:i + 2], coef[i:i + 1], [w_tar, w_tar, w_tar], p, 0.0, 1, False, 1, variant=2,
                          xt_is_pair=True, ctrl_step0=i, blend_state=blend_state, mos_pull=False, pnp=pnp)
        xt = torch.stack([rc, ed], dim=1)
    if kind == "stock":
        controller.cur_step = getattr(controller, "cur_step", 0) + steps
        if getattr(controller, "local_blend", None) is not None:
            controller.local_blend.counter += steps
    return xt[:, 1].to(xT.device), xt[:, 0].to(xT.device)


def nmg_p2p(model, xT, xT_ori, etas: float = 0.0, prompts="", cfg_scales=None, prog_bar=False, zs=None, controller=None,
            guidance_noise_map: float = 10.0, grad_scale: float = 5e+3):
    """Reference signature (inversion/p2p_baselines.py:195): Noise Map Guidance with P2P (hybrid, see _grad_guided).  Returns (edited, reconstructed)."""
    assert len(prompts) >= 2 and etas == 0, "P2P requires source and target prompts, with eta is set to 0 for NMG"      # reference assert (:216)
    return _grad_guided(model, xT, xT_ori, prompts, cfg_scales, zs, controller, False, "nmg", guidance_noise_map, grad_scale)


def nmg_pnp(model, xT, xT_ori, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, guidance_noise_map=10.0, grad_scale: float = 5e+3):
    """Reference signature (inversion/pnp_baselines.py:32): Noise Map Guidance with Plug-and-Play injection (hybrid, see _grad_guided)."""
    assert len(prompts) >= 2 and etas == 0, "PnP requires source and target prompts, with eta is set to 0 for NMG"      # reference assert (:53)
    return _grad_guided(model, xT, xT_ori, prompts, cfg_scales, zs, None, True, "nmg", guidance_noise_map, grad_scale)


def nulltext_pnp(model, xT, xT_ori, etas=0, prompts="", cfg_scales=None, prog_bar=False, zs=None, optimization_steps: int = 10, epsilon: float = 1e-5):
    """Reference signature (inversion/pnp_baselines.py:134): Null-Text Inversion with Plug-and-Play injection (hybrid, see _grad_guided)."""
    assert len(prompts) >= 2 and etas == 0, "PnP requires source and target prompts, with eta is set to 0"
    return _grad_guided(model, xT, xT_ori, prompts, cfg_scales, zs, None, True, "nulltext", optimization_steps=optimization_steps, epsilon=epsilon)
