#!/usr/bin/env python
"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/text-guided:
`inversion_forward_process_ddpm` -> `make_controller` -> `register_attention_control` -> `h_Edit_p2p_implicit`)
on the oracle's seeded random-init SD-1.x UNet restatement (no pretrained weights exist offline).

Runs only in the build container (the GPU box has no /root/reference); the resulting tensors are committed so
that tests on the GPU box can compare the CUDA path and the oracle port against the reference's own outputs.

    python tests/make_golden.py --config tiny      # seconds
    python tests/make_golden.py --config sd15      # BASELINE.json configs[0]: full SD-1.5 UNet, T=10, ~10 min CPU
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.pipeline import OraclePipeline  # noqa: E402
from oracle.sd_unet import UNetConfig  # noqa: E402
from refload import load_reference  # noqa: E402

PROMPTS = ["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"]
BLEND = ("lizard", "lizard")

# sampler variants beyond the north-star implicit+P2P loop: name -> (unet cfg, T, K, mode)
VARIANTS = {
    "tiny_p2p_explicit": (UNetConfig.tiny(sample_size=64), 6, 1, "p2p_explicit"),
    "tiny_R_implicit_mos2": (UNetConfig.tiny(sample_size=64), 5, 2, "R_implicit"),
    "tiny_R_explicit": (UNetConfig.tiny(sample_size=64), 6, 1, "R_explicit"),
    "tiny_masactrl_mos2": (UNetConfig.tiny(sample_size=64), 5, 2, "masactrl"),
    "tiny_R_implicit_skip2": (UNetConfig.tiny(sample_size=64), 6, 1, "R_implicit_skip"),
    "tiny_pnp": (UNetConfig.tiny(sample_size=64), 6, 1, "pnp"),
    "small32_pnp_mos2": (UNetConfig.tiny(sample_size=32), 4, 2, "pnp"),
    # editor built with total_steps = T while K = 2: the reference's step_idx = range(start, total_steps) ENDS the injection after
    # editor step 5 of 10 (masactrl.py:36,58)
    "tiny_masactrl_mos2_stop": (UNetConfig.tiny(sample_size=64), 5, 2, "masactrl"),
    # explicit layer_idx / step_idx lists (masactrl.py:33-36)
    "tiny_masactrl_lists": (UNetConfig.tiny(sample_size=64), 6, 1, "masactrl"),
    # BASELINE.json configs[2] at full SD-1.5 geometry: implicit h-Edit + MasaCtrl (the sampler the reference ships), T = 50
    "sd15_config3_T50_masactrl": (UNetConfig.sd15(), 50, 1, "masactrl"),
    # baseline samplers of main_p2p.py --mode ef_p2p / pnp_inv_p2p / ef and main_masactrl.py (inversion/p2p_baselines.py, masactrl_baselines.py)
    "tiny_ef_p2p": (UNetConfig.tiny(sample_size=64), 6, 1, "ef_p2p"),
    "tiny_pnpinv_p2p": (UNetConfig.tiny(sample_size=64), 6, 1, "pnpinv_p2p"),
    "tiny_ef": (UNetConfig.tiny(sample_size=64), 6, 1, "ef"),
    "tiny_ef_masactrl": (UNetConfig.tiny(sample_size=64), 6, 1, "ef_masactrl"),
    # full SD-1.5 geometry, T = 10, for the samplers that had tiny-width goldens only
    "sd15_pnp_T10": (UNetConfig.sd15(), 10, 1, "pnp"),
    "sd15_p2p_explicit_T10": (UNetConfig.sd15(), 10, 1, "p2p_explicit"),
    "sd15_ef_p2p_T10": (UNetConfig.sd15(), 10, 1, "ef_p2p"),
    # Plug-and-Play baselines (inversion/pnp_baselines.py:317,244)
    "tiny_ef_pnp": (UNetConfig.tiny(sample_size=64), 6, 1, "ef_pnp"),
    # Noise Map Guidance with P2P (p2p_baselines.py:195): differentiates through one UNet forward per step
    "tiny_nmg_p2p": (UNetConfig.tiny(sample_size=64), 6, 1, "nmg_p2p"),
    "tiny_nmg_pnp": (UNetConfig.tiny(sample_size=64), 6, 1, "nmg_pnp"),
    "tiny_nulltext_pnp": (UNetConfig.tiny(sample_size=64), 6, 1, "nulltext_pnp"),
    "tiny_np_pnp": (UNetConfig.tiny(sample_size=64), 6, 1, "np_pnp"),
}
# MutualSelfAttentionControl arguments per masactrl variant (default: start_step 2, start_layer 10, total_steps T*K)
MASA_ARGS = {
    "tiny_masactrl_mos2_stop": dict(start_step=2, start_layer=10, total_steps=5),
    "tiny_masactrl_lists": dict(start_step=0, start_layer=0, layer_idx=[3, 8, 9, 12, 15], step_idx=[1, 2, 4]),
    "sd15_config3_T50_masactrl": dict(start_step=4, start_layer=10, total_steps=50),
}

CASES = {
    # name: (unet cfg, T, K, is_replace, blend)
    "tiny_refine_blend": (UNetConfig.tiny(sample_size=64), 10, 1, False, True),
    "tiny_replace_mos2": (UNetConfig.tiny(sample_size=64), 6, 2, True, True),
    "tiny_refine_noblend": (UNetConfig.tiny(sample_size=64), 6, 1, False, False),
    "sd15_config1": (UNetConfig.sd15(), 10, 1, False, True),
    # LocalBlend with a NON-TRIVIAL mask: at the reference's default threshold 0.3 the word maps of a random-init UNet are so flat that the
    # mask covers the whole latent (4096 of 4096 pixels in every golden above); th = 0.9 cuts it to a partial region that grows step by step,
    # which is what makes "number of mask pixels that differ" a real check (ptp_classes.py:17,40-42: `th` is a LocalBlend argument)
    "tiny_refine_blend_th09": (UNetConfig.tiny(sample_size=64), 10, 1, False, True),
    "sd15_config2_T50_refine_blend_th09": (UNetConfig.sd15(), 50, 1, False, True),
    # LocalBlend with substruct_words (ptp_classes.py:28-38,66-67): the region of "branch" is excluded from the blend mask
    "tiny_refine_blend_substruct": (UNetConfig.tiny(sample_size=64), 6, 1, False, True),
    # BASELINE.json configs[1] (the headline): full SD-1.5 geometry, T = 50; two images with different prompts / controllers / noise
    # (Refine + Reweight + LocalBlend, and Replace + Reweight + LocalBlend) that the GPU test runs inside ONE mixed batch of 8
    "sd15_config2_T50_refine_blend": (UNetConfig.sd15(), 50, 1, False, True),
    "sd15_config2_T50_replace": (UNetConfig.sd15(), 50, 1, True, True),
    # fast fixtures for the CPU (-m "not gpu") suite: 32x32 latent (LocalBlend needs 64x64, so it is off here)
    "small32_refine": (UNetConfig.tiny(sample_size=32), 4, 1, False, False),
    "small32_replace_mos2": (UNetConfig.tiny(sample_size=32), 3, 2, True, False),
}


# per-case overrides of (prompts, blend words, seed of w0 / inversion noise)
CASE_INPUTS = {
    "sd15_config2_T50_replace": (["a photo of a cat sitting on a bench", "a photo of a dog sitting on a bench"], ("cat", "dog"), 1),
}


SUBSTRUCT = {"tiny_refine_blend_substruct": (("branch",), ("branch",))}
BLEND_TH = {"tiny_refine_blend_th09": (0.9, 0.9), "sd15_config2_T50_refine_blend_th09": (0.9, 0.9)}


def run_case(ref, name, cfg, T, K, is_replace, blend, xa=0.4, sa=0.35):
    global PROMPTS, BLEND
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", os.cpu_count())))
    seed = 0
    if name in CASE_INPUTS:
        PROMPTS, BLEND, seed = CASE_INPUTS[name]
    model = OraclePipeline(cfg, seed=0)
    model.scheduler.set_timesteps(T)
    g = torch.Generator(device="cpu").manual_seed(seed)
    w0 = torch.randn(1, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g) * 0.18215 * 5
    torch.manual_seed(seed)    # the reference draws its inversion noise from the global RNG (ddpm_inversion.py:48)
    t0 = time.time()
    _, zs, wts, _ = ref.ddpm_inversion.inversion_forward_process_ddpm(
        model, w0, etas=1.0, prog_bar=False, prompt=PROMPTS[0], cfg_scale_src=1.0, num_inference_steps=T)
    t_inv = time.time() - t0
    blend_word = ((BLEND[0],), (BLEND[1],)) if blend else None
    eq = {"words": (BLEND[1],), "values": (1.25 if K > 1 else 2.0,)} if blend else None    # main_p2p.py:196-201
    controller = ref.ptp_controller_utils.make_controller(
        prompts=PROMPTS, is_replace_controller=is_replace, cross_replace_steps=xa, self_replace_steps=sa,
        blend_word=blend_word, equilizer_params=eq, num_steps=T, tokenizer=model.tokenizer, device=model.device)
    substruct = SUBSTRUCT.get(name)
    blend_th = BLEND_TH.get(name)
    if substruct is not None or blend_th is not None:   # the reference's make_controller never passes substruct_words / th; its LocalBlend class takes them
        controller.local_blend = ref.ptp_classes.LocalBlend(PROMPTS, T, blend_word, substruct_words=substruct, th=blend_th or (0.3, 0.3),
                                                            tokenizer=model.tokenizer, device=model.device)
    ref.ptp_utils.register_attention_control(model, controller)
    trace = []
    orig_cb = controller.step_callback

    def cb(x):
        y = orig_cb(x)
        trace.append(y.detach().clone())
        return y

    controller.step_callback = cb
    t0 = time.time()
    edited, recon = ref.p2p_h_edit.h_Edit_p2p_implicit(
        model, xT=wts[T], eta=1.0, prompts=PROMPTS, cfg_scales=[1.0, 5.0, 7.5], prog_bar=False, zs=zs[:T],
        controller=controller, weight_reconstruction=0.1, optimization_steps=K, after_skip_steps=T,
        is_ddim_inversion=False)
    t_edit = time.time() - t0
    enc = ref.inversion_utils.encode_text
    out = {
        "meta": dict(name=name, T=T, K=K, is_replace=is_replace, blend=blend, xa=xa, sa=sa, prompts=PROMPTS, substruct_words=substruct, blend_th=blend_th,
                     blend_words=BLEND, cfg_scales=[1.0, 5.0, 7.5], eta=1.0, weight_reconstruction=0.1,
                     unet=dict(block_out_channels=list(cfg.block_out_channels), sample_size=cfg.sample_size,
                               cross_attention_dim=cfg.cross_attention_dim, heads=cfg.attention_head_dim),
                     weights="oracle.sd_unet.seeded_init_(seed=0)", w0=f"randn(seed {seed})*0.18215*5", seed=seed,
                     generator="tests/make_golden.py", seconds=dict(inversion=t_inv, edit=t_edit),
                     torch=torch.__version__, threads=torch.get_num_threads()),
        "w0": w0, "zs": zs[:T].clone(), "xT": wts[T].clone(),
        "ctx_uncond": enc(model, [""]), "ctx_src": enc(model, [PROMPTS[0]]), "ctx_tar": enc(model, [PROMPTS[1]]),
        "edited": edited.detach().clone(), "recon": recon.detach().clone(),
        "trace": torch.stack(trace),
        "tables": {
            "alpha_words": controller.cross_replace_alpha.reshape(T + 1, 77).clone(),
            "self_window": list(controller.num_self_replace),
        },
    }
    inner = controller.prev_controller if hasattr(controller, "prev_controller") and controller.prev_controller is not None else controller
    if is_replace:
        out["tables"]["replace_matrix"] = inner.mapper[0].clone()
    else:
        out["tables"]["mapper"] = inner.mapper[0].clone()
        out["tables"]["refine_alpha"] = inner.alphas.reshape(77).clone()
    if hasattr(controller, "equalizer"):
        out["tables"]["equalizer"] = controller.equalizer.reshape(77).clone()
    if controller.local_blend is not None:
        out["tables"]["blend_alpha"] = controller.local_blend.alpha_layers.reshape(2, 77).clone()
        out["tables"]["start_blend"] = controller.local_blend.start_blend
    return out


def run_variant(ref, name, cfg, T, K, mode, xa=0.4, sa=0.35):
    """Other reference samplers on the same set-up: h_Edit_p2p_explicit (p2p_h_edit.py:380), h_Edit_R_implicit/explicit (:162,:21),
    h_Edit_masactrl_implicit (masactrl_h_edit.py:14)."""
    import importlib
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", os.cpu_count())))
    t_start = time.time()
    model = OraclePipeline(cfg, seed=0)
    model.scheduler.set_timesteps(T)
    g = torch.Generator(device="cpu").manual_seed(0)
    w0 = torch.randn(1, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g) * 0.18215 * 5
    prompts = list(PROMPTS)
    if mode in ("masactrl", "ef_masactrl"):
        prompts[0] = ""                      # main_masactrl.py:180
    torch.manual_seed(0)
    _, zs, wts, _ = ref.ddpm_inversion.inversion_forward_process_ddpm(
        model, w0, etas=1.0, prog_bar=False, prompt=prompts[0], cfg_scale_src=1.0, num_inference_steps=T)
    meta_extra = {}
    kw = dict(xT=wts[T], eta=1.0, prompts=prompts, cfg_scales=[1.0, 5.0, 7.5], prog_bar=False, zs=zs[:T], after_skip_steps=T,
              is_ddim_inversion=False)
    if mode == "p2p_explicit":
        controller = ref.ptp_controller_utils.make_controller(
            prompts=prompts, is_replace_controller=False, cross_replace_steps=xa, self_replace_steps=sa,
            blend_word=((BLEND[0],), (BLEND[1],)), equilizer_params={"words": (BLEND[1],), "values": (2.0,)}, num_steps=T,
            tokenizer=model.tokenizer, device=model.device)
        ref.ptp_utils.register_attention_control(model, controller)
        edited, recon = ref.p2p_h_edit.h_Edit_p2p_explicit(model, controller=controller, **kw)
    elif mode == "R_implicit":
        edited, recon = ref.p2p_h_edit.h_Edit_R_implicit(model, controller=None, weight_reconstruction=0.1, optimization_steps=K, **kw)
    elif mode == "R_implicit_skip":
        # main_p2p.py:220-236 with --skip 2: start from wts[T-2], use the first T-2 noise maps
        S = T - 2
        kw.update(xT=wts[S], zs=zs[:S], after_skip_steps=S)
        edited, recon = ref.p2p_h_edit.h_Edit_R_implicit(model, controller=None, weight_reconstruction=0.1, optimization_steps=K, **kw)
        meta_extra = dict(after_skip_steps=S)
    elif mode == "R_explicit":
        edited, recon = ref.p2p_h_edit.h_Edit_R_explicit(model, controller=None, **kw)
    elif mode == "masactrl":
        masa_pkg = importlib.import_module("masactrl")
        sys.modules.setdefault("masa_ctrl", masa_pkg)                       # reference typo: masactrl.py:8 imports `masa_ctrl`
        sys.modules.setdefault("masa_ctrl.masactrl_utils", importlib.import_module("masactrl.masactrl_utils"))
        masa = importlib.import_module("masactrl.masactrl")
        mh = importlib.import_module("inversion.masactrl_h_edit")
        margs = dict(MASA_ARGS.get(name, dict(start_step=2, start_layer=10, total_steps=T * K)))
        margs.setdefault("total_steps", T * K)
        editor = masa.MutualSelfAttentionControl(**margs)
        importlib.import_module("masactrl.masactrl_utils").regiter_attention_editor_diffusers(model, editor)
        edited, recon = mh.h_Edit_masactrl_implicit(model, optimization_steps=K, **kw)
        meta_extra = dict(masa_start_step=margs["start_step"], masa_start_layer=margs["start_layer"], masa_total_steps=margs["total_steps"],
                          masa_layer_idx=margs.get("layer_idx"), masa_step_idx=margs.get("step_idx"), seconds=time.time() - t_start)
    elif mode == "pnp":
        # main_plugnplay.py:186-208 (h_edit_R_pnp), with the injection fractions raised so that short schedules have both
        # injected and un-injected steps
        pu = importlib.import_module("plug_n_play.pnp_utils")
        ph = importlib.import_module("inversion.pnp_h_edit")
        f_t, attn_t = int(T * 0.8), int(T * 0.5)
        qk_ts, conv_ts = model.scheduler.timesteps[:attn_t], model.scheduler.timesteps[:f_t]
        pu.register_attention_control_efficient(model, qk_ts)
        pu.register_conv_control_efficient(model, conv_ts)
        edited, recon = ph.h_Edit_PnP_implicit(model, optimization_steps=K, **kw)
        meta_extra = dict(pnp_qk_timesteps=[int(t) for t in qk_ts], pnp_conv_timesteps=[int(t) for t in conv_ts])
    elif mode == "nmg_p2p":
        pb = importlib.import_module("inversion.p2p_baselines")
        controller = ref.ptp_controller_utils.make_controller(
            prompts=prompts, is_replace_controller=False, cross_replace_steps=xa, self_replace_steps=sa,
            blend_word=((BLEND[0],), (BLEND[1],)), equilizer_params={"words": (BLEND[1],), "values": (2.0,)}, num_steps=T,
            tokenizer=model.tokenizer, device=model.device)
        ref.ptp_utils.register_attention_control(model, controller)
        edited, recon = pb.nmg_p2p(model, xT=wts[T], xT_ori=wts[:T + 1], etas=0.0, prompts=prompts, cfg_scales=[1.0, 7.5], prog_bar=False,
                                   zs=zs[:T], controller=controller, guidance_noise_map=10.0, grad_scale=5e+3)
        meta_extra = dict(baseline_cfg_scales=[1.0, 7.5], is_ddim_inversion=False, guidance_noise_map=10.0, grad_scale=5e+3)
        nmg_extra = {"xT_ori": torch.stack([w for w in wts[:T + 1]]).clone()}
    elif mode in ("nmg_pnp", "nulltext_pnp"):
        pu = importlib.import_module("plug_n_play.pnp_utils")
        pnb = importlib.import_module("inversion.pnp_baselines")
        f_t, attn_t = int(T * 0.8), int(T * 0.5)
        qk_ts, conv_ts = model.scheduler.timesteps[:attn_t], model.scheduler.timesteps[:f_t]
        pu.register_attention_control_efficient(model, qk_ts)
        pu.register_conv_control_efficient(model, conv_ts)
        bkw = dict(xT=wts[T], xT_ori=wts[:T + 1], etas=0.0, prompts=prompts, cfg_scales=[1.0, 7.5], prog_bar=False, zs=zs[:T])
        if mode == "nmg_pnp":
            edited, recon = pnb.nmg_pnp(model, guidance_noise_map=10.0, grad_scale=5e+3, **bkw)
        else:
            edited, recon = pnb.nulltext_pnp(model, optimization_steps=3, epsilon=1e-5, **bkw)
        meta_extra = dict(pnp_qk_timesteps=[int(t) for t in qk_ts], pnp_conv_timesteps=[int(t) for t in conv_ts], baseline_cfg_scales=[1.0, 7.5],
                          is_ddim_inversion=False, guidance_noise_map=10.0, grad_scale=5e+3, nulltext_steps=3)
        nmg_extra = {"xT_ori": torch.stack([w for w in wts[:T + 1]]).clone()}
    elif mode in ("ef_pnp", "np_pnp"):
        pu = importlib.import_module("plug_n_play.pnp_utils")
        pnb = importlib.import_module("inversion.pnp_baselines")
        f_t, attn_t = int(T * 0.8), int(T * 0.5)
        qk_ts, conv_ts = model.scheduler.timesteps[:attn_t], model.scheduler.timesteps[:f_t]
        pu.register_attention_control_efficient(model, qk_ts)
        pu.register_conv_control_efficient(model, conv_ts)
        if mode == "ef_pnp":
            edited, recon = pnb.ef_or_pnp_inv_w_pnp(model, xT=wts[T], etas=0, prompts=prompts, cfg_scales=[1.0, 7.5], prog_bar=False, zs=zs[:T], is_ddim_inversion=False)
        else:
            edited, recon = pnb.negative_prompt_pnp(model, xT=wts[T], etas=0, prompts=prompts, cfg_scales=[1.0, 7.5], prog_bar=False, zs=zs[:T])
        meta_extra = dict(pnp_qk_timesteps=[int(t) for t in qk_ts], pnp_conv_timesteps=[int(t) for t in conv_ts], baseline_cfg_scales=[1.0, 7.5],
                          is_ddim_inversion=False)
    elif mode in ("ef_p2p", "pnpinv_p2p", "ef", "ef_masactrl"):
        # main_p2p.py:245-255 / main_masactrl.py: the Edit Friendly and PnP Inversion baselines
        pb = importlib.import_module("inversion.p2p_baselines")
        bkw = dict(xT=wts[T], etas=1.0, prog_bar=False, zs=zs[:T], is_ddim_inversion=(mode == "pnpinv_p2p"))
        if mode in ("ef_p2p", "pnpinv_p2p"):
            controller = ref.ptp_controller_utils.make_controller(
                prompts=prompts, is_replace_controller=False, cross_replace_steps=xa, self_replace_steps=sa,
                blend_word=((BLEND[0],), (BLEND[1],)), equilizer_params={"words": (BLEND[1],), "values": (2.0,)}, num_steps=T,
                tokenizer=model.tokenizer, device=model.device)
            ref.ptp_utils.register_attention_control(model, controller)
            edited, recon = pb.ef_or_pnp_inv_w_p2p(model, prompts=prompts, cfg_scales=[1.0, 7.5], controller=controller, **bkw)
        elif mode == "ef":
            ref.ptp_utils.register_attention_control(model, ref.ptp_classes.EmptyControl())
            edited = pb.ef_wo_p2p(model, prompts=[prompts[1]], cfg_scales=[7.5], controller=None, **bkw)
            recon = edited
        else:
            masa_pkg = importlib.import_module("masactrl")
            sys.modules.setdefault("masa_ctrl", masa_pkg)
            sys.modules.setdefault("masa_ctrl.masactrl_utils", importlib.import_module("masactrl.masactrl_utils"))
            masa = importlib.import_module("masactrl.masactrl")
            mb = importlib.import_module("inversion.masactrl_baselines")
            editor = masa.MutualSelfAttentionControl(start_step=2, start_layer=10, total_steps=T)
            importlib.import_module("masactrl.masactrl_utils").regiter_attention_editor_diffusers(model, editor)
            edited, recon = mb.ef_or_pnp_inv_w_masactrl(model, prompts=prompts, cfg_scales=[1.0, 7.5], **bkw)
            meta_extra = dict(masa_start_step=2, masa_start_layer=10, masa_total_steps=T, masa_layer_idx=None, masa_step_idx=None)
        meta_extra = dict(meta_extra, baseline_cfg_scales=[1.0, 7.5], is_ddim_inversion=(mode == "pnpinv_p2p"))
    else:
        raise ValueError(mode)
    enc = ref.inversion_utils.encode_text
    return {
        **(nmg_extra if mode in ("nmg_p2p", "nmg_pnp", "nulltext_pnp") else {}),
        "meta": dict(name=name, mode=mode, T=T, K=K, xa=xa, sa=sa, prompts=prompts, blend_words=BLEND, cfg_scales=[1.0, 5.0, 7.5], eta=1.0,
                     weight_reconstruction=0.1, is_replace=False, blend=(mode in ("p2p_explicit", "ef_p2p", "pnpinv_p2p", "nmg_p2p")),
                     unet=dict(block_out_channels=list(cfg.block_out_channels), sample_size=cfg.sample_size,
                               cross_attention_dim=cfg.cross_attention_dim, heads=cfg.attention_head_dim),
                     weights="oracle.sd_unet.seeded_init_(seed=0)", generator="tests/make_golden.py", torch=torch.__version__, **meta_extra),
        "w0": w0, "zs": kw["zs"].clone(), "xT": kw["xT"].clone(),
        "ctx_uncond": enc(model, [""]), "ctx_src": enc(model, [prompts[0]]), "ctx_tar": enc(model, [prompts[1]]),
        "edited": edited.detach().clone(), "recon": recon.detach().clone(),
    }


def run_ddim_inversion(ref, name="tiny_ddim_inversion", T=6):
    """Reference ddim_inversion (inversion/ddim_inversion.py:55) with the eta = 0 scheduler (steps_offset 0, main_p2p.py:139-141)."""
    import importlib
    cfg = UNetConfig.tiny(sample_size=64)
    model = OraclePipeline(cfg, seed=0, steps_offset=0)
    model.scheduler.set_timesteps(T)
    g = torch.Generator(device="cpu").manual_seed(0)
    w0 = torch.randn(1, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g) * 0.18215 * 5
    di = importlib.import_module("inversion.ddim_inversion")
    latent, zs, latents = di.ddim_inversion(model, w0, PROMPTS[0], 1.0)
    return {"meta": dict(name=name, T=T, prompt=PROMPTS[0], cfg_scale=1.0, steps_offset=0,
                         unet=dict(block_out_channels=list(cfg.block_out_channels), sample_size=cfg.sample_size,
                                   cross_attention_dim=cfg.cross_attention_dim, heads=cfg.attention_head_dim)),
            "w0": w0, "latent": latent.clone(), "zs": zs.clone(), "latents": torch.cat(latents).clone()}


def run_style(name="tiny_style_mos2", T=4, K=2, weight=0.5, full=False):
    """The UNMODIFIED style sampler (text-guided-n-style/inversion/h_edit.py:14 `h_Edit_p2p_implicit(model, image_encoder, ...)`) with the
    reference's own `CLIPEncoder.get_gram_matrix_residual` (clip_guidance/base_clip.py:55) on a small seeded CLIP ViT and the oracle's
    small VAE decoder.  Must run in its own process: the style tree has its own `inversion` / `p2p` packages."""
    import importlib
    style_root = "/root/reference/text-guided-n-style"
    for p in (os.path.join(ROOT, "tests", "refshim"), style_root):
        sys.path.insert(0, p)
    from oracle.clip_visual import tiny_style_encoder
    from oracle.vae import AutoencoderKLDecoder, VAEConfig
    he = importlib.import_module("inversion.h_edit")
    inv = importlib.import_module("inversion.ddpm_inversion")
    iu = importlib.import_module("inversion.inversion_utils")
    pcu = importlib.import_module("p2p.ptp_controller_utils")
    ptu = importlib.import_module("p2p.ptp_utils")
    bc = importlib.import_module("clip_guidance.base_clip")
    clip_model = importlib.import_module("clip_guidance.clip.model")
    import torchvision

    torch.set_num_threads(os.cpu_count())
    # full = True: BASELINE.json configs[3] geometry -- SD-1.5 UNet, SD VAE decoder (128, 256, 512, 512), CLIP ViT-B/16 width (768, 12 heads;
    # the Gram residual only reads the first three transformer blocks, base_clip.py:55-66, so three are built)
    cfg = UNetConfig.sd15() if full else UNetConfig.tiny(sample_size=64)
    vw = 768 if full else 64
    model = OraclePipeline(cfg, seed=0)
    model.vae = AutoencoderKLDecoder(VAEConfig() if full else VAEConfig.tiny())
    model.scheduler.set_timesteps(T)
    tiny = tiny_style_encoder(width=vw)
    # the reference's CLIPEncoder, constructed without its weight download (base_clip.py:31-52), carrying the same seeded weights
    enc = bc.CLIPEncoder.__new__(bc.CLIPEncoder)
    torch.nn.Module.__init__(enc)
    enc.clip_model = clip_model.CLIP(embed_dim=32, image_resolution=224, vision_layers=3, vision_width=vw, vision_patch_size=16, context_length=8,
                                     vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=1)
    enc.clip_model.visual.load_state_dict(tiny.visual.state_dict())
    enc.preprocess = torchvision.transforms.Normalize((0.48145466 * 2 - 1, 0.4578275 * 2 - 1, 0.40821073 * 2 - 1),
                                                      (0.26862954 * 2, 0.26130258 * 2, 0.27577711 * 2))
    enc.ref = tiny.ref
    for q in enc.parameters():
        q.requires_grad_(False)
    g = torch.Generator(device="cpu").manual_seed(0)
    w0 = torch.randn(1, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g) * 0.18215 * 5
    torch.manual_seed(0)
    _, zs, wts, _ = inv.inversion_forward_process_ddpm(model, w0, etas=1.0, prog_bar=False, prompt=PROMPTS[0], cfg_scale_src=1.0,
                                                       num_inference_steps=T)
    # main_edit.py:179-195: blend words are forced off for the style path; Refine controller
    controller = pcu.make_controller(prompts=PROMPTS, is_replace_controller=False, cross_replace_steps=0.4, self_replace_steps=0.35,
                                     blend_word=None, equilizer_params=None, num_steps=T, tokenizer=model.tokenizer, device=model.device)
    ptu.register_attention_control(model, controller)
    edited, recon = he.h_Edit_p2p_implicit(model, enc, xT=wts[T], eta=1.0, prompts=PROMPTS, cfg_scales=[1.0, 5.0, 7.5], prog_bar=False,
                                           zs=zs[:T], controller=controller, weight_edit_clip=weight, optimization_steps=K,
                                           after_skip_steps=T, is_ddim_inversion=False)
    # the same run without the reward (image_encoder=None) shows how much of the edit the style term accounts for
    controller2 = pcu.make_controller(prompts=PROMPTS, is_replace_controller=False, cross_replace_steps=0.4, self_replace_steps=0.35,
                                      blend_word=None, equilizer_params=None, num_steps=T, tokenizer=model.tokenizer, device=model.device)
    ptu.register_attention_control(model, controller2)
    edited_ns, _ = he.h_Edit_p2p_implicit(model, None, xT=wts[T], eta=1.0, prompts=PROMPTS, cfg_scales=[1.0, 5.0, 7.5], prog_bar=False,
                                          zs=zs[:T], controller=controller2, weight_edit_clip=weight, optimization_steps=K,
                                          after_skip_steps=T, is_ddim_inversion=False)
    enc_t = iu.encode_text
    out = {"meta": dict(name=name, mode="style", T=T, K=K, xa=0.4, sa=0.35, prompts=PROMPTS, cfg_scales=[1.0, 5.0, 7.5], eta=1.0,
                        weight_edit_clip=weight, is_replace=False, blend=False,
                        unet=dict(block_out_channels=list(cfg.block_out_channels), sample_size=cfg.sample_size,
                                  cross_attention_dim=cfg.cross_attention_dim, heads=cfg.attention_head_dim),
                        vae="oracle.vae.AutoencoderKLDecoder(VAEConfig%s, seed 7)" % ("()" if full else ".tiny()"),
                        clip="oracle.clip_visual.tiny_style_encoder(seed 11, width %d)" % vw, full_geometry=bool(full), clip_width=vw,
                        generator="tests/make_golden.py --config style", torch=torch.__version__),
           "w0": w0, "zs": zs[:T].clone(), "xT": wts[T].clone(),
           "ctx_uncond": enc_t(model, [""]), "ctx_src": enc_t(model, [PROMPTS[0]]), "ctx_tar": enc_t(model, [PROMPTS[1]]),
           "edited": edited.detach().clone(), "recon": recon.detach().clone(), "edited_no_style": edited_ns.detach().clone()}
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(out, path)
    print(name, "->", path, "|edited| %.4f" % out["edited"].abs().mean().item(),
          "| style term moved the edit by %.4f (rel)" % ((edited - edited_ns).norm() / edited_ns.norm()).item(), flush=True)


def run_face_full(name="face256_irse50_lpips_k2", T=4, K=2, weight=1500.0, lin_gain=100.0, celebahq=False):
    """The UNMODIFIED face-swapping sampler `h_Edit_R` (face-swapping/inversion/h_edit_R.py:7) driving the reference's OWN reward classes:
    `IDLoss` (arcface/arcface_model.py:12) around the reference `Backbone(112, 50, 'ir_se')` and `LPIPS_Loss` (:72) around `lpips.LPIPS`
    (tests/refshim/lpips: the package itself is not available offline), both with seeded random weights and seeded images instead of
    model_ir_se50.pth / the jpg files their constructors open.  256 x 256 images (the crop [35:223, 32:220] the identity network sees
    needs them), a 4-level DDPM UNet carrying the oracle's seeded weights."""
    import importlib
    import numpy as np
    for p in (os.path.join(ROOT, "tests", "refshim"), "/root/reference/face-swapping"):
        sys.path.insert(0, p)
    from oracle.face_unet import FaceUNet, FaceUNetConfig
    from hedit_b200 import reward_nets
    ref_model = importlib.import_module("diffusion.diffusion")
    du = importlib.import_module("diffusion.diffusion_utils")
    inv = importlib.import_module("inversion.sde_inversion")
    he = importlib.import_module("inversion.h_edit_R")
    am = importlib.import_module("arcface.arcface_model")
    irse = importlib.import_module("arcface.facial_recognition.model_irse")
    import lpips
    torch.set_num_threads(os.cpu_count())
    # celebahq = True: the denoiser at BASELINE.json configs[4] geometry (ch 128 x (1, 1, 2, 2, 4, 4), attention at 16 x 16)
    cfg = FaceUNetConfig() if celebahq else FaceUNetConfig(ch=64, ch_mult=(1, 1, 2, 2), image_size=256, attn_resolutions=(32,))
    model = ref_model.Model(cfg.as_reference_dict())
    model.load_state_dict(FaceUNet(cfg).state_dict())
    model.eval()
    betas = torch.from_numpy(du.get_beta_schedule(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    seq = (np.arange(0, 1000, 1000 // T) + 1)[::-1]
    g = torch.Generator(device="cpu").manual_seed(0)
    smooth = lambda: torch.nn.functional.interpolate(torch.randn(1, 3, 32, 32, generator=g), size=(256, 256), mode="bicubic", align_corners=False).mul(0.6).clamp(-1, 1)
    x0, ref_img = smooth(), smooth()
    # the reference classes without their file-reading constructors, same attributes (arcface_model.py:16-39, 76-90)
    idloss = am.IDLoss.__new__(am.IDLoss)
    torch.nn.Module.__init__(idloss)
    idloss.facenet = irse.Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode="ir_se")
    idloss.facenet.load_state_dict(reward_nets._seed_init(reward_nets.IRSE50(), 3).state_dict())
    idloss.facenet.eval().requires_grad_(False)
    idloss.pool, idloss.face_pool, idloss.ref = torch.nn.AdaptiveAvgPool2d((256, 256)), torch.nn.AdaptiveAvgPool2d((112, 112)), ref_img
    lpipsloss = am.LPIPS_Loss.__new__(am.LPIPS_Loss)
    torch.nn.Module.__init__(lpipsloss)
    lpipsloss.lpips_loss = lpips.LPIPS(net="vgg")
    vsd = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 4).state_dict()
    for k in vsd:                                   # random `lin` heads are tiny (0..0.02): scaled so that the LPIPS move is visible next to the identity move
        if k.startswith("lins."):
            vsd[k] = vsd[k] * lin_gain
    lpipsloss.lpips_loss.net.load_state_dict(vsd)
    lpipsloss.lpips_loss.eval().requires_grad_(False)
    lpipsloss.src = x0.clone()
    with torch.no_grad():
        _, zs, xts, _ = inv.inversion_forward_process_sde(model, x0, betas, seq, etas=1.0, num_inference_steps=T, device="cpu")
    kw = dict(eta=1.0, zs=zs[:T], weight_edit_face=weight, optimization_steps=K, after_skip_steps=T, num_inference_steps=T, soft_face_mask=None)
    edited = he.h_Edit_R(model, lpipsloss, idloss, xts[T].clone(), betas, seq, **kw)
    only_id = he.h_Edit_R(model, None, idloss, xts[T].clone(), betas, seq, **kw)
    recon = he.h_Edit_R(model, None, None, xts[T].clone(), betas, seq, **kw)
    out = {"meta": dict(name=name, mode="face_full", T=T, K=K, weight_edit_face=weight, seq=[int(v) for v in seq],
                        unet=dict(ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=list(cfg.attn_resolutions),
                                  image_size=cfg.image_size),
                        rewards="reference IDLoss around Backbone(112, 50, 'ir_se') with reward_nets._seed_init(IRSE50(), 3) weights; reference LPIPS_Loss "
                                "around tests/refshim/lpips.LPIPS('vgg') with reward_nets._seed_init(LPIPSVGG16(), 4) weights, lin heads x lin_gain",
                        lin_gain=lin_gain, irse_seed=3, vgg_seed=4,
                        generator="tests/make_golden.py --config face_full", torch=torch.__version__),
           "x0": x0, "ref_img": ref_img, "zs": zs[:T].clone(), "xT": xts[T].clone(), "betas": betas,
           "edited": edited.detach().clone(), "edited_id_only": only_id.detach().clone(), "no_reward": recon.detach().clone()}
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(out, path)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print(name, "->", path, "| both rewards moved the result by %.4f (rel), identity alone by %.4f | no-reward run returns x0 to %.2e" %
          (rel(edited, recon), rel(only_id, recon), (recon - x0).abs().max().item()), flush=True)


def run_face(name="tiny_face_k2", T=5, K=2, weight=50.0):
    """The UNMODIFIED face-swapping sampler (face-swapping/inversion/h_edit_R.py:7 `h_Edit_R`) and inversion
    (inversion/sde_inversion.py:54) on the reference's own `Model` class (diffusion/diffusion.py:193) in a small configuration carrying
    the oracle's seeded weights, with seeded stand-ins for the ArcFace / LPIPS reward models (their weights do not exist offline).
    Must run in its own process (the face tree has its own `inversion` package)."""
    import importlib
    import numpy as np
    sys.path.insert(0, "/root/reference/face-swapping")
    from oracle.face_unet import FaceUNet, FaceUNetConfig, TinyIDLoss, TinyLPIPSLoss
    ref_model = importlib.import_module("diffusion.diffusion")
    du = importlib.import_module("diffusion.diffusion_utils")
    inv = importlib.import_module("inversion.sde_inversion")
    he = importlib.import_module("inversion.h_edit_R")
    torch.set_num_threads(os.cpu_count())
    cfg = FaceUNetConfig.tiny()
    oracle_model = FaceUNet(cfg)
    model = ref_model.Model(cfg.as_reference_dict())
    model.load_state_dict(oracle_model.state_dict())
    model.eval()
    betas = torch.from_numpy(du.get_beta_schedule(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000)).float()
    skip = 1000 // T
    seq = (np.arange(0, 1000, skip) + 1)[::-1]                        # main_edit.py:140-142
    g = torch.Generator(device="cpu").manual_seed(0)
    x0 = torch.tanh(torch.randn(1, 3, cfg.image_size, cfg.image_size, generator=g))
    ref_img = torch.tanh(torch.randn(1, 3, cfg.image_size, cfg.image_size, generator=g))
    idloss, lpipsloss = TinyIDLoss(ref_img), TinyLPIPSLoss(x0.clone())
    with torch.no_grad():
        _, zs, xts, _ = inv.inversion_forward_process_sde(model, x0, betas, seq, etas=1.0, num_inference_steps=T, device="cpu")
    edited = he.h_Edit_R(model, lpipsloss, idloss, xts[T].clone(), betas, seq, eta=1.0, zs=zs[:T], weight_edit_face=weight, optimization_steps=K,
                         after_skip_steps=T, num_inference_steps=T, soft_face_mask=None)
    recon = he.h_Edit_R(model, None, None, xts[T].clone(), betas, seq, eta=1.0, zs=zs[:T], weight_edit_face=weight, optimization_steps=K,
                        after_skip_steps=T, num_inference_steps=T, soft_face_mask=None)
    out = {"meta": dict(name=name, mode="face", T=T, K=K, weight_edit_face=weight, seq=[int(v) for v in seq],
                        unet=dict(ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=list(cfg.attn_resolutions),
                                  image_size=cfg.image_size),
                        rewards="oracle.face_unet.TinyIDLoss(seed 21) / TinyLPIPSLoss(seed 22)", generator="tests/make_golden.py --config face",
                        torch=torch.__version__),
           "x0": x0, "ref_img": ref_img, "zs": zs[:T].clone(), "xT": xts[T].clone(), "betas": betas,
           "edited": edited.detach().clone(), "no_reward": recon.detach().clone()}
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(out, path)
    print(name, "->", path, "|edited| %.4f" % out["edited"].abs().mean().item(),
          "| rewards moved the result by %.4f (rel) | no-reward run returns x0 to %.2e" %
          (((edited - recon).norm() / recon.norm()).item(), (recon - x0).abs().max().item()), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="tiny", choices=["face_full", "face_full_celebahq", "style_sd15", "baselines", "tiny", "sd15", "sd15_config1", "sd15_config2", "sd15_config2_T50_refine_blend_th09", "tiny_refine_blend_th09", "masa", "small32", "variants", "inversion", "pnp", "style", "face", "all"])
    ap.add_argument("--only", default=None, help="generate just this VARIANTS entry")
    args = ap.parse_args()
    if args.only:
        ref = load_reference()
        cfg, T, K, mode = VARIANTS[args.only]
        out = run_variant(ref, args.only, cfg, T, K, mode)
        path = os.path.join(ROOT, "tests", "golden", f"{args.only}.pt")
        torch.save(out, path)
        print(args.only, "->", path, "|edited| %.4f" % out["edited"].abs().mean().item(), flush=True)
        return
    if args.config == "style":
        run_style()
        return
    if args.config == "style_sd15":
        run_style(name="sd15_config4_T10_style_k3", T=10, K=3, weight=0.5, full=True)
        return
    if args.config == "face":
        run_face()
        return
    if args.config == "face_full":
        run_face_full()
        return
    if args.config == "face_full_celebahq":
        run_face_full(name="celebahq_config5_T5_irse50_lpips_k3", T=5, K=3, celebahq=True)
        return
    ref = load_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    if args.config in ("inversion", "all"):
        out = run_ddim_inversion(ref)
        torch.save(out, os.path.join(ROOT, "tests", "golden", "tiny_ddim_inversion.pt"))
        print("tiny_ddim_inversion |zs|", out["zs"].abs().mean().item(), flush=True)
    if args.config in ("variants", "all", "pnp", "masa", "baselines"):
        for name, (cfg, T, K, mode) in VARIANTS.items():
            if args.config == "baselines" and os.path.exists(os.path.join(ROOT, "tests", "golden", f"{name}.pt")):
                continue
            if args.config == "baselines" and mode not in ("ef_p2p", "pnpinv_p2p", "ef", "ef_masactrl", "ef_pnp", "np_pnp"):
                continue
            if args.config == "masa" and (mode != "masactrl" or os.path.exists(os.path.join(ROOT, "tests", "golden", f"{name}.pt"))):
                continue
            if args.config in ("variants", "all") and name.startswith("sd15") and os.path.exists(os.path.join(ROOT, "tests", "golden", f"{name}.pt")):
                continue
            if args.config == "pnp" and mode not in ("pnp", "R_implicit_skip"):
                continue
            if args.config == "pnp" and os.path.exists(os.path.join(ROOT, "tests", "golden", f"{name}.pt")):
                continue
            out = run_variant(ref, name, cfg, T, K, mode)
            path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
            torch.save(out, path)
            print(name, "->", path, "|edited| %.4f" % out["edited"].abs().mean().item(), flush=True)
    for name, (cfg, T, K, rep, blend) in CASES.items():
        if args.config in ("variants", "inversion", "pnp", "masa", "baselines") or (args.config != "all" and not name.startswith(args.config)):
            continue
        out = run_case(ref, name, cfg, T, K, rep, blend)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        print(name, "edit %.1fs" % out["meta"]["seconds"]["edit"], "->", path,
              "|edited| %.4f" % out["edited"].abs().mean().item(), flush=True)


if __name__ == "__main__":
    main()
