"""-m gpu: parity of the CUDA path on the HEADLINE configuration (BASELINE.json configs[1]: SD-1.5 geometry, 64x64 latent, T = 50, implicit
h-Edit-R + P2P, a mixed batch of 8) against goldens produced by the UNMODIFIED reference sampler (tests/make_golden.py --config
sd15_config2: `h_Edit_p2p_implicit`, p2p_h_edit.py:529, on the oracle's seeded SD-1.5 UNet, fp32 CPU, ~25 CPU-minutes per image).

Two golden images ride in ONE native call together with six other images (different prompts, controllers and noise):
  image 0 = Refine + Reweight(2.0) + LocalBlend ("green lizard" -> "brown lizard"), image 1 = Replace + Reweight + LocalBlend (cat -> dog).

What is asserted, per golden image, on the reconstruction row AND the edited row:
  * relative L2 and PER-PIXEL max-abs of the final latents;
  * the per-step trajectory (relative L2 of xt after every one of the 50 timesteps);
  * the number of LocalBlend mask pixels that differ at every step (the mask is recovered from the trajectory itself: outside the mask
    the blended edit row equals the reconstruction row bit for bit, ptp_classes.py:71).
The BOUNDS are calibrated, not guessed: the oracle loop itself is run on the same inputs (a) in fp32 on the GPU -- this shows
how far two fp32 evaluations with different summation orders drift apart over 50 chained steps -- and (b) with every contraction's
operands rounded to the CUDA path's 16-bit operand type (tests/fp16_emulation.py).  The CUDA path, a different realisation of the same
operand-rounding noise, must stay within FACTOR x the emulated run's deviation from the golden (plus a small absolute floor), metric by
metric.  The measured table goes to gpurun_out/parity_headline.json and DESIGN.md section 6.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from fp16_emulation import operand_rounding  # noqa: E402
from gpu_util import opdtype, rel_err  # noqa: E402
from oracle import h_edit as oh  # noqa: E402
from oracle.pipeline import OraclePipeline  # noqa: E402
from oracle.sd_unet import UNetConfig  # noqa: E402
from oracle_run import GOLDEN_DIR, load_golden, spec_from_meta  # noqa: E402

import hedit_b200  # noqa: E402
from hedit_b200 import UNetEngine  # noqa: E402

NAMES = ["sd15_config2_T50_refine_blend", "sd15_config2_T50_replace"]
FILLERS = [   # (prompts, blend words, is_replace)
    (["a cat sitting next to a mirror", "a silver cat sculpture sitting next to a mirror"], ("cat", "cat"), False),
    (["two birds on a wire", "two parrots on a wire"], ("birds", "parrots"), True),
    (["a photo of a house on a hill", "a photo of a castle on a hill in winter"], ("house", "castle"), False),
    (["a red car parked on a street", "a blue car parked on a street"], ("car", "car"), True),
    (["a bowl of apples on a table", "a bowl of oranges on a wooden table"], ("apples", "oranges"), False),
    (["a man riding a horse", "a man riding a camel"], ("horse", "camel"), True),
]
FACTOR = 3.0          # CUDA-path deviation <= FACTOR x emulated-oracle deviation (+ floor)
FLOOR_REL = 2e-3      # relative-L2 floor (one UNet call alone differs by ~1.3e-3)
FLOOR_ABS = 2e-2      # per-pixel floor on latents of magnitude ~1-10
FLOOR_FLIPS = 8       # LocalBlend mask pixels (of 4096)

have = all(os.path.exists(os.path.join(GOLDEN_DIR, n + ".pt")) for n in NAMES)


def _masks(trace):
    """(T, 64, 64) bool: where the edit row differs from the reconstruction row after each step (== inside the LocalBlend mask once
    blending has started)."""
    return (trace[:, 1] != trace[:, 0]).any(dim=1)


def _metrics(ed, rc, tr, g):
    """deviation of a run (edited, recon, trace (T,2,C,h,w)) from the golden"""
    T = g["trace"].shape[0]
    m = {}
    m["ed_rel"], m["ed_max"] = rel_err(ed, g["edited"])
    m["rc_rel"], m["rc_max"] = rel_err(rc, g["recon"])
    m["step_rel_rc"] = [rel_err(tr[i, 0], g["trace"][i, 0])[0] for i in range(T)]
    m["step_rel_ed"] = [rel_err(tr[i, 1], g["trace"][i, 1])[0] for i in range(T)]
    m["step_max_ed"] = [rel_err(tr[i, 1], g["trace"][i, 1])[1] for i in range(T)]
    mk, mg = _masks(tr), _masks(g["trace"])
    m["mask_flips"] = [int((mk[i] != mg[i]).sum()) for i in range(T)]
    m["mask_pixels_golden"] = [int(mg[i].sum()) for i in range(T)]
    return m


def _oracle_on_gpu(model, g, emulate, spec_fn=spec_from_meta):
    """The oracle port of the reference loop (oracle/h_edit.py) evaluated with torch on the GPU, fp32 (TF32 off), optionally with
    operand rounding."""
    meta = g["meta"]
    dev = torch.device("cuda")
    spec = spec_fn(meta, model.tokenizer)
    for k in ("alpha_words", "mapper", "refine_alpha", "replace_matrix", "equalizer", "blend_alpha"):
        v = getattr(spec, k, None)
        if torch.is_tensor(v):
            setattr(spec, k, v.to(dev))
    sched = model.scheduler
    sched.set_timesteps(meta["T"])
    sched.alphas_cumprod = sched.alphas_cumprod.to(dev)
    sched.final_alpha_cumprod = sched.final_alpha_cumprod.to(dev)
    trace = []
    args = (model.unet, sched, g["ctx_uncond"].to(dev), g["ctx_src"].to(dev), g["ctx_tar"].to(dev), g["xT"].to(dev), g["zs"].to(dev), spec,
            meta["cfg_scales"])
    kw = dict(eta=meta["eta"], weight_reconstruction=meta["weight_reconstruction"], optimization_steps=meta["K"], after_skip_steps=meta["T"],
              is_ddim_inversion=False, trace=trace)
    if emulate:
        with operand_rounding(opdtype()):
            ed, rc = oh.h_edit_p2p_implicit(*args, **kw)
    else:
        ed, rc = oh.h_edit_p2p_implicit(*args, **kw)
    sched.alphas_cumprod = sched.alphas_cumprod.cpu()
    sched.final_alpha_cumprod = sched.final_alpha_cumprod.cpu()
    return ed.cpu(), rc.cpu(), torch.stack(trace).cpu()


@pytest.fixture(scope="module")
def headline():
    if not have:
        pytest.skip("headline goldens missing (python tests/make_golden.py --config sd15_config2)")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    gs = [load_golden(n) for n in NAMES]
    T = gs[0]["meta"]["T"]
    model = OraclePipeline(UNetConfig.sd15(), seed=0)
    model.scheduler.set_timesteps(T)
    B = 8
    eng = UNetEngine.from_unet(model.unet, max_samples=5 * B, max_contexts=1 + 2 * B)
    # ---- the mixed batch: images 0, 1 = the goldens; 2..7 = fillers
    ctrls, ctx, xT, zs = [], [gs[0]["ctx_uncond"]], [], []
    for g in gs:
        meta = g["meta"]
        bw = meta["blend_words"]
        ctrls.append(hedit_b200.make_controller(meta["prompts"], meta["is_replace"], meta["xa"], meta["sa"], blend_word=((bw[0],), (bw[1],)),
                                                equilizer_params={"words": (bw[1],), "values": (2.0,)}, num_steps=T, tokenizer=model.tokenizer))
        ctx += [g["ctx_src"], g["ctx_tar"]]
        xT.append(g["xT"].reshape(1, 4, 64, 64))
        zs.append(g["zs"].reshape(1, T, 4, 64, 64))
        assert torch.equal(g["ctx_uncond"], gs[0]["ctx_uncond"])
    gen = torch.Generator().manual_seed(77)
    enc = lambda p: model.text_encoder(model.tokenizer(p).input_ids)[0]
    for prompts, bw, rep in FILLERS:
        ctrls.append(hedit_b200.make_controller(prompts, rep, 0.4, 0.35, blend_word=((bw[0],), (bw[1],)),
                                                equilizer_params={"words": (bw[1],), "values": (2.0,)}, num_steps=T, tokenizer=model.tokenizer))
        ctx += [enc([prompts[0]]), enc([prompts[1]])]
        xT.append(torch.randn(1, 4, 64, 64, generator=gen))
        zs.append(torch.randn(1, T, 4, 64, 64, generator=gen))
    plan = hedit_b200.compile_edit_plan(ctrls, T)
    assert plan.is_replace.tolist() == [0, 1, 0, 1, 0, 1, 0, 1] and plan.has_blend.all()
    ts, coef = hedit_b200.step_tables(model.scheduler, T, 1.0, False)
    meta = gs[0]["meta"]
    ed, rc, tr = eng.edit(torch.cat(xT).cuda(), torch.cat(zs).cuda(), torch.cat(ctx).cuda(), ts, coef, meta["cfg_scales"], plan,
                          meta["weight_reconstruction"], 1, False, 1, trace=True)
    stats = dict(eng.last_stats)
    ed, rc, tr = ed.cpu(), rc.cpu(), tr.cpu()          # tr: (T, B, 2, C, h, w)
    del eng
    torch.cuda.empty_cache()
    out = {"stats": stats, "images": []}
    unet = model.unet.cuda()
    for b, g in enumerate(gs):
        cuda_m = _metrics(ed[b:b + 1], rc[b:b + 1], tr[:, b], g)
        f32_m = _metrics(*_oracle_on_gpu(model, g, emulate=False), g)
        emu_m = _metrics(*_oracle_on_gpu(model, g, emulate=True), g)
        out["images"].append({"name": NAMES[b], "cuda": cuda_m, "oracle_fp32_gpu": f32_m, "oracle_16bit_operands": emu_m,
                              "recon_vs_w0": rel_err(rc[b:b + 1], g["w0"]), "golden_recon_vs_w0": rel_err(g["recon"], g["w0"]),
                              "latent_absmax": g["edited"].abs().max().item(), "latent_rms": g["edited"].pow(2).mean().sqrt().item()})
    unet.cpu()
    out["finite_all"] = bool(torch.isfinite(ed).all() and torch.isfinite(rc).all())
    short = lambda v: [round(x, 5) if isinstance(x, float) else x for x in v[::7]]
    for im in out["images"]:
        print(f"\n== {im['name']}: |latent| max {im['latent_absmax']:.2f} rms {im['latent_rms']:.2f}; recon vs w0 (known answer): cuda {im['recon_vs_w0'][0]:.3e}, "
              f"golden {im['golden_recon_vs_w0'][0]:.3e}")
        for k in ("cuda", "oracle_16bit_operands", "oracle_fp32_gpu"):
            m = im[k]
            print(f"  {k:22s} edited rel {m['ed_rel']:.3e} max-abs {m['ed_max']:.3e} | recon rel {m['rc_rel']:.3e} max-abs {m['rc_max']:.3e} | "
                  f"mask flips max {max(m['mask_flips'])} (golden mask {max(m['mask_pixels_golden'])} px)")
            print(f"     per-step edited rel (every 7th): {short(m['step_rel_ed'])}")
            print(f"     per-step recon  rel (every 7th): {short(m['step_rel_rc'])}")
            print(f"     per-step mask flips (every 7th): {m['mask_flips'][::7]}")
    d = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_headline.json"), "w") as f:
            json.dump(out, f)
    return out


def test_headline_batch_runs_the_headline_schedule(headline):
    assert headline["stats"]["sample_forwards"] == 8 * 50 * 7        # exact-reuse schedule, B = 8, T = 50
    assert headline["finite_all"]


@pytest.mark.parametrize("b", [0, 1])
def test_headline_final_latents_per_pixel(headline, b):
    im = headline["images"][b]
    c, e = im["cuda"], im["oracle_16bit_operands"]
    for k in ("ed_rel", "rc_rel"):
        assert c[k] <= FACTOR * e[k] + FLOOR_REL, (k, c[k], e[k])
    for k in ("ed_max", "rc_max"):
        assert c[k] <= FACTOR * e[k] + FLOOR_ABS, (k, c[k], e[k])
    # intrinsic known answer: the reconstruction row returns the inverted latent w0 (as far as the golden itself does)
    assert im["recon_vs_w0"][0] <= FACTOR * e["rc_rel"] + im["golden_recon_vs_w0"][0] + FLOOR_REL


@pytest.mark.parametrize("b", [0, 1])
def test_headline_per_step_trajectory(headline, b):
    im = headline["images"][b]
    c, e = im["cuda"], im["oracle_16bit_operands"]
    for k in ("step_rel_rc", "step_rel_ed"):
        # the emulated run's curve, made monotone (a step where it happens to dip does not tighten the bound)
        env, run = [], 0.0
        for v in e[k]:
            run = max(run, v)
            env.append(run)
        bad = [(i, cv, ev) for i, (cv, ev) in enumerate(zip(c[k], env)) if cv > FACTOR * ev + FLOOR_REL]
        assert not bad, (k, bad[:5])


@pytest.mark.parametrize("b", [0, 1])
def test_headline_localblend_mask_flips(headline, b):
    im = headline["images"][b]
    c, e = im["cuda"], im["oracle_16bit_operands"]
    assert max(c["mask_pixels_golden"]) > 0                             # LocalBlend was active in the golden
    worst_e = max(e["mask_flips"])
    bad = [(i, f) for i, f in enumerate(c["mask_flips"]) if f > FACTOR * worst_e + FLOOR_FLIPS]
    assert not bad, (bad[:5], worst_e)


# ---------------------------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2] at full geometry: implicit h-Edit + MasaCtrl (the sampler the reference ships), SD-1.5 UNet, T = 50, against the
# golden of the unmodified `h_Edit_masactrl_implicit` + `MutualSelfAttentionControl(4, 10)` (tests/make_golden.py --config masa).
MASA = "sd15_config3_T50_masactrl"


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN_DIR, MASA + ".pt")), reason="golden missing")
def test_config3_masactrl_full_geometry():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = load_golden(MASA)
    meta = g["meta"]
    T = meta["T"]
    model = OraclePipeline(UNetConfig.sd15(), seed=0)
    model.scheduler.set_timesteps(T)
    hedit_b200.regiter_attention_editor_diffusers(model, hedit_b200.MutualSelfAttentionControl(meta["masa_start_step"], meta["masa_start_layer"],
                                                                                              total_steps=meta["masa_total_steps"]))
    ed, rc = hedit_b200.h_Edit_masactrl_implicit(model, g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"], cfg_scales=meta["cfg_scales"],
                                                 zs=g["zs"].cuda(), optimization_steps=meta["K"], after_skip_steps=T, is_ddim_inversion=False)
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    r_rc, m_rc = rel_err(rc.cpu(), g["recon"])
    print(f"{MASA}: edited rel {r_ed:.3e} max-abs {m_ed:.3e} (|latent| max {g['edited'].abs().max().item():.1f}) | recon rel {r_rc:.3e} max-abs {m_rc:.3e}")
    assert hedit_b200.get_engine(model).last_stats["sample_forwards"] == 7 * T
    # same operand-rounding process as the headline loop (50 chained steps, 16-bit operands): the calibrated headline levels x 3
    assert r_ed < 2e-2 and r_rc < 1.5e-1 and m_rc < 0.6


TH09 = "sd15_config2_T50_refine_blend_th09"


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN_DIR, TH09 + ".pt")), reason="golden missing")
def test_headline_partial_localblend_mask_full_geometry():
    """The headline loop with LocalBlend th = 0.9 (a genuinely partial, step-dependent mask at SD-1.5 geometry; at the default 0.3 the mask
    of a random-init UNet covers the whole latent): mask pixels on the other side than in the reference are counted against the
    16-bit-operand-emulated oracle run."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = load_golden(TH09)
    meta = g["meta"]
    T = meta["T"]
    model = OraclePipeline(UNetConfig.sd15(), seed=0)
    model.scheduler.set_timesteps(T)
    eng = UNetEngine.from_unet(model.unet, max_samples=5, max_contexts=4)
    bw = meta["blend_words"]
    ctrl = hedit_b200.make_controller(meta["prompts"], False, meta["xa"], meta["sa"], blend_word=((bw[0],), (bw[1],)),
                                      equilizer_params={"words": (bw[1],), "values": (2.0,)}, num_steps=T, tokenizer=model.tokenizer)
    ctrl.local_blend.th = tuple(meta["blend_th"])
    plan = hedit_b200.compile_edit_plan([ctrl], T)
    ts, coef = hedit_b200.step_tables(model.scheduler, T, 1.0, False)
    ctx = torch.cat([g["ctx_uncond"], g["ctx_src"], g["ctx_tar"]]).cuda()
    ed, rc, tr = eng.edit(g["xT"].reshape(1, 4, 64, 64).cuda(), g["zs"].reshape(1, T, 4, 64, 64).cuda(), ctx, ts, coef, meta["cfg_scales"], plan,
                          meta["weight_reconstruction"], 1, False, 1, trace=True)
    del eng
    torch.cuda.empty_cache()
    c = _metrics(ed.cpu(), rc.cpu(), tr.cpu()[:, 0], g)
    model.unet.cuda()
    import oracle_run
    spec0 = oracle_run.spec_from_meta
    oracle_run_spec = lambda m, tok: _with_th(spec0(m, tok), meta["blend_th"])
    e = _metrics(*_oracle_on_gpu(model, g, emulate=True, spec_fn=oracle_run_spec), g)
    model.unet.cpu()
    print(f"{TH09}: golden mask px (every 7th) {c['mask_pixels_golden'][::7]}")
    print(f"  cuda : edited rel {c['ed_rel']:.3e} max {c['ed_max']:.3e} recon rel {c['rc_rel']:.3e} | mask flips (every 7th) {c['mask_flips'][::7]} max {max(c['mask_flips'])}")
    print(f"  16-bit-operand oracle: edited rel {e['ed_rel']:.3e} max {e['ed_max']:.3e} recon rel {e['rc_rel']:.3e} | mask flips max {max(e['mask_flips'])}")
    assert 0 < min(c["mask_pixels_golden"][12:]) and max(c["mask_pixels_golden"][12:]) < 4096
    assert max(c["mask_flips"]) <= FACTOR * max(e["mask_flips"]) + 2 * 16          # + two 16x16 cells
    assert c["rc_rel"] <= FACTOR * e["rc_rel"] + FLOOR_REL
    assert c["ed_rel"] <= FACTOR * e["ed_rel"] + FLOOR_REL


def _with_th(spec, th):
    spec.blend_th = float(th[0])
    return spec
