"""GPU: combined text-guided + style editing -- the native loop with the reward hook against the outputs of the
UNMODIFIED reference sampler text-guided-n-style/inversion/h_edit.py:14 (tests/make_golden.py --config style)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.clip_visual import tiny_style_encoder  # noqa: E402
from oracle.pipeline import OraclePipeline  # noqa: E402
from oracle.vae import AutoencoderKLDecoder, VAEConfig  # noqa: E402
from oracle_run import cfg_from_meta, load_golden  # noqa: E402

import hedit_b200  # noqa: E402
from gpu_util import rel_err  # noqa: E402

TOL_STYLE = 4e-2      # whole-loop tolerance of the fp16-operand UNet (tests/test_gpu_unet.py TOL_LOOP); the reward branch is fp32


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = load_golden("tiny_style_mos2")
    meta = g["meta"]
    model = OraclePipeline(cfg_from_meta(meta), seed=0)
    model.scheduler.set_timesteps(meta["T"])
    model.vae = AutoencoderKLDecoder(VAEConfig.tiny()).cuda()
    enc = tiny_style_encoder().cuda()
    return g, meta, model, enc


@pytest.mark.parametrize("native_vae,native_clip", [(True, True), (True, False), (False, False)])
def test_style_sampler_matches_reference_golden(native_vae, native_clip):
    g, meta, model, enc = _setup()
    T, K = meta["T"], meta["K"]
    ctrl = hedit_b200.make_controller(meta["prompts"], False, meta["xa"], meta["sa"], blend_word=None, equilizer_params=None, num_steps=T,
                                      tokenizer=model.tokenizer)
    # reference signature; autocast off because the golden was produced on CPU, where the reference's autocast("cuda") is inert
    ed, rc = hedit_b200.style.h_Edit_p2p_implicit(model, enc, xT=g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"],
                                                  cfg_scales=meta["cfg_scales"], zs=g["zs"].cuda(), controller=ctrl,
                                                  weight_edit_clip=meta["weight_edit_clip"], optimization_steps=K, after_skip_steps=T,
                                                  is_ddim_inversion=False, autocast=False, native_vae=native_vae, native_clip=native_clip)
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    r_rc, m_rc = rel_err(rc.cpu(), g["recon"])
    r_ns, _ = rel_err(ed.cpu(), g["edited_no_style"])
    print(f"style (native VAE {native_vae}, native CLIP {native_clip}): edited rel {r_ed:.3e} max {m_ed:.3e} | recon rel {r_rc:.3e} | distance to the no-style edit {r_ns:.3e}")
    assert r_ed < TOL_STYLE and r_rc < TOL_STYLE
    assert r_ns > 5 * TOL_STYLE            # the reward term is a large part of the result, so the check above is meaningful
    # without an image encoder the sampler reduces to the text-guided loop (h_edit.py:157 `if image_encoder:`)
    ctrl2 = hedit_b200.make_controller(meta["prompts"], False, meta["xa"], meta["sa"], blend_word=None, equilizer_params=None, num_steps=T,
                                       tokenizer=model.tokenizer)
    ed0, _ = hedit_b200.style.h_Edit_p2p_implicit(model, None, xT=g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"],
                                                  cfg_scales=meta["cfg_scales"], zs=g["zs"].cuda(), controller=ctrl2,
                                                  weight_edit_clip=meta["weight_edit_clip"], optimization_steps=K, after_skip_steps=T)
    assert rel_err(ed0.cpu(), g["edited_no_style"])[0] < TOL_STYLE


def test_style_batch_is_per_image():
    """B = 2 copies of the same edit give the single-image result for both (per-image RMS, per-image Gram loss)."""
    g, meta, model, enc = _setup()
    T, K = meta["T"], meta["K"]
    mk = lambda: hedit_b200.make_controller(meta["prompts"], False, meta["xa"], meta["sa"], blend_word=None, equilizer_params=None,
                                            num_steps=T, tokenizer=model.tokenizer)
    xT = g["xT"].reshape(1, *g["xT"].shape[-3:]).cuda().repeat(2, 1, 1, 1)
    zs = g["zs"].reshape(1, *g["zs"].shape).cuda().repeat(2, 1, 1, 1, 1)
    ed, rc = hedit_b200.style.h_edit_style_batch(model, enc, xT, zs, [meta["prompts"]] * 2, meta["cfg_scales"], [mk(), mk()], eta=1.0,
                                                 weight_edit_clip=meta["weight_edit_clip"], optimization_steps=K, after_skip_steps=T, autocast=False)
    for b in range(2):
        assert rel_err(ed[b:b + 1].cpu(), g["edited"])[0] < TOL_STYLE
    # torch's backward kernels (cuDNN dgrad, atomics) are not run-to-run bit-stable and the random-init UNet amplifies last-bit
    # differences across steps, so the two slots agree to loop tolerance, not bitwise
    assert rel_err(ed[0], ed[1])[0] < TOL_STYLE / 2


def test_guidance_callback_errors_surface():
    g, meta, model, enc = _setup()
    T = meta["T"]

    def bad(x0):
        raise ValueError("boom")

    xT = g["xT"].reshape(1, *g["xT"].shape[-3:]).cuda()
    zs = g["zs"].reshape(1, *g["zs"].shape).cuda()
    with pytest.raises(ValueError, match="boom"):
        hedit_b200.style.h_edit_style_batch(model, None, xT, zs, [meta["prompts"]], meta["cfg_scales"], None, after_skip_steps=T, guidance_fn=bad)


def test_style_sampler_full_geometry_golden():
    """BASELINE.json configs[3] geometry (SD-1.5 UNet, SD VAE decoder, CLIP ViT-B/16 width), T = 10, K = 3 Langevin steps per timestep: the
    native loop + native VAE decode / backward + native CLIP-Gram reward against the UNMODIFIED reference sampler's output
    (tests/make_golden.py --config style_sd15)."""
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", "sd15_config4_T10_style_k3.pt")):
        pytest.skip("full-size golden missing")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = load_golden("sd15_config4_T10_style_k3")
    meta = g["meta"]
    T, K = meta["T"], meta["K"]
    model = OraclePipeline(cfg_from_meta(meta), seed=0)
    model.scheduler.set_timesteps(T)
    model.vae = AutoencoderKLDecoder(VAEConfig()).cuda()
    enc = tiny_style_encoder(width=meta["clip_width"]).cuda()
    ctrl = hedit_b200.make_controller(meta["prompts"], False, meta["xa"], meta["sa"], blend_word=None, equilizer_params=None, num_steps=T,
                                      tokenizer=model.tokenizer)
    ed, rc = hedit_b200.style.h_Edit_p2p_implicit(model, enc, xT=g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"],
                                                  cfg_scales=meta["cfg_scales"], zs=g["zs"].cuda(), controller=ctrl,
                                                  weight_edit_clip=meta["weight_edit_clip"], optimization_steps=K, after_skip_steps=T,
                                                  is_ddim_inversion=False, autocast=False, native_vae=True, native_clip=True)
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    r_rc, m_rc = rel_err(rc.cpu(), g["recon"])
    r_ns, _ = rel_err(ed.cpu(), g["edited_no_style"])
    print(f"sd15_config4_T10_style_k3: edited rel {r_ed:.3e} max {m_ed:.3e} (|latent| max {float(g['edited'].abs().max()):.1f}) | recon rel {r_rc:.3e} max {m_rc:.3e} | "
          f"distance to the no-style edit {r_ns:.3e}")
    assert r_ed < TOL_STYLE and r_rc < TOL_STYLE
    assert r_ns > 3 * r_ed            # the style term is live at this geometry too
