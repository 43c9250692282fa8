"""CPU: pins the oracle.  (1) against the committed golden fixtures produced by the UNMODIFIED reference loop
(tests/make_golden.py) -- runs anywhere; (2) live against the reference's own set-up classes when /root/reference
is present (build container only)."""
import pytest
import torch

from oracle import p2p as op
from oracle.pipeline import OraclePipeline, ToyTokenizer
from oracle_run import load_golden, run_oracle_on_golden
from refload import load_reference, reference_available

PAIRS = [
    (["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"], "lizard", "lizard"),
    (["a cat sitting next to a mirror", "a silver cat sculpture sitting next to a mirror"], "cat", "cat"),
    (["a photo of a house on a hill", "a photo of a castle on a hill in winter"], "house", "castle"),
    (["two birds on a wire", "two parrots on a wire"], "birds", "parrots"),
]


@pytest.mark.parametrize("name", ["small32_refine", "small32_replace_mos2"])
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    ed, rc, tr, spec = run_oracle_on_golden(g)
    # same torch build, same thread count, same op order as the reference loop -> bit-exact (0.0) here; a different CPU thread count
    # or torch build changes the fp32 summation order inside the convolutions (measured 1.2e-4 with 2 instead of 8 threads)
    assert (ed - g["edited"]).abs().max().item() <= 5e-4
    assert (rc - g["recon"]).abs().max().item() <= 5e-4
    assert (tr - g["trace"]).abs().max().item() <= 5e-4
    # intrinsic known answer: the reconstruction row returns the inverted latent
    assert (rc - g["w0"]).abs().max().item() < 1e-3
    tb = g["tables"]
    assert torch.equal(tb["alpha_words"], spec.alpha_words) and tuple(tb["self_window"]) == tuple(spec.self_window)
    for k in ("mapper", "refine_alpha", "replace_matrix", "equalizer", "blend_alpha"):
        if k in tb:
            assert torch.equal(tb[k], getattr(spec, k)), k


def test_oracle_pnp_matches_reference_golden():
    """oracle/pnp.py against the outputs of the unmodified reference h_Edit_PnP_implicit + pnp_utils registration (MOS K=2)."""
    from oracle import pnp as opnp
    from oracle_run import cfg_from_meta
    g = load_golden("small32_pnp_mos2")
    meta = g["meta"]
    model = OraclePipeline(cfg_from_meta(meta), seed=0)
    model.scheduler.set_timesteps(meta["T"])
    st = opnp.PnPState(set(meta["pnp_qk_timesteps"]), set(meta["pnp_conv_timesteps"]))
    opnp.install_oracle_pnp(model.unet, st)
    ed, rc = opnp.h_edit_pnp_implicit(model.unet, model.scheduler, g["ctx_uncond"], g["ctx_src"], g["ctx_tar"], g["xT"], g["zs"], st,
                                      meta["cfg_scales"], eta=meta["eta"], optimization_steps=meta["K"], after_skip_steps=meta["T"])
    assert (ed - g["edited"]).abs().max().item() <= 1e-4 and (rc - g["recon"]).abs().max().item() <= 1e-4
    assert (rc - g["w0"]).abs().max().item() < 1e-3          # intrinsic known answer: the reconstruction row returns the inverted latent


@pytest.mark.slow
def test_oracle_matches_reference_golden_blend64():
    g = load_golden("tiny_replace_mos2")
    ed, rc, tr, _ = run_oracle_on_golden(g)
    assert (ed - g["edited"]).abs().max().item() <= 1e-4 and (tr - g["trace"]).abs().max().item() <= 1e-4


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("prompts,bw_src,bw_tar", PAIRS)
@pytest.mark.parametrize("is_replace", [False, True])
def test_setup_tables_match_reference_classes(prompts, bw_src, bw_tar, is_replace):
    ref = load_reference()
    tok = ToyTokenizer()
    if is_replace and len(prompts[0].split(" ")) != len(prompts[1].split(" ")):
        pytest.skip("replace controller needs equal word counts")
    T = 12
    kw = dict(cross_replace_steps=0.4, self_replace_steps=0.35, blend_word=((bw_src,), (bw_tar,)),
              equilizer_params={"words": (bw_tar,), "values": (2.0,)}, num_steps=T, tokenizer=tok)
    c = ref.ptp_controller_utils.make_controller(prompts=prompts, is_replace_controller=is_replace, device="cpu", **kw)
    s = op.make_edit_spec(prompts, is_replace, kw["cross_replace_steps"], kw["self_replace_steps"], kw["blend_word"],
                          kw["equilizer_params"], T, tok)
    assert torch.equal(c.cross_replace_alpha.reshape(T + 1, 77), s.alpha_words)
    assert tuple(c.num_self_replace) == tuple(s.self_window)
    assert torch.equal(c.equalizer.reshape(77), s.equalizer)
    assert torch.equal(c.local_blend.alpha_layers.reshape(2, 77), s.blend_alpha) and c.local_blend.start_blend == s.start_blend
    inner = c.prev_controller
    if is_replace:
        assert torch.equal(inner.mapper[0], s.replace_matrix)
    else:
        assert torch.equal(inner.mapper[0], s.mapper) and torch.equal(inner.alphas.reshape(77), s.refine_alpha)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_hook_matches_reference_controller():
    """One attention-layer call of the oracle hook vs the reference controller object on the same probabilities
    (edit region, in-place semantics, stored views, step counter)."""
    ref = load_reference()
    tok = ToyTokenizer()
    prompts = PAIRS[1][0]
    T = 6
    c = ref.ptp_controller_utils.make_controller(prompts=prompts, is_replace_controller=False, cross_replace_steps=0.4, self_replace_steps=0.5,
                                                 blend_word=None, equilizer_params={"words": ("silver",), "values": (2.0,)}, num_steps=T,
                                                 tokenizer=tok, device="cpu")
    c.num_att_layers = 2
    s = op.make_edit_spec(prompts, False, 0.4, 0.5, None, {"words": ("silver",), "values": (2.0,)}, T, tok)
    st = op.P2PState(num_att_layers=2)
    g = torch.Generator().manual_seed(3)
    for step in range(T):
        for is_cross, M in ((False, 64), (True, 77)):
            p = torch.softmax(torch.randn(32, 64, M, generator=g), -1)
            a, b = p.clone(), p.clone()
            c(a, is_cross, "down", True)
            op.p2p_hook(st, s, b, is_cross, "down", True)
            assert torch.equal(a, b), (step, is_cross)
        assert c.cur_step == st.cur_step == step + 1
    for k in c.attention_store:
        for x, y in zip(c.attention_store[k], st.attention_store[k]):
            assert torch.equal(x, y)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_clip_gram_restatement_matches_reference_classes():
    """oracle/clip_visual.py vs the reference's own CLIP class + CLIPEncoder.get_gram_matrix_residual (base_clip.py:55) on the same seeded
    weights.  Runs in a subprocess: the style tree has its own `inversion` / `p2p` packages that would shadow the text-guided ones."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, importlib, torch
sys.path[:0] = [%r, %r, "/root/reference/text-guided-n-style"]
from oracle.clip_visual import tiny_style_encoder
bc = importlib.import_module("clip_guidance.base_clip")
cm = importlib.import_module("clip_guidance.clip.model")
import torchvision
tiny = tiny_style_encoder()
enc = bc.CLIPEncoder.__new__(bc.CLIPEncoder); torch.nn.Module.__init__(enc)
enc.clip_model = cm.CLIP(embed_dim=32, image_resolution=224, vision_layers=3, vision_width=64, vision_patch_size=16, context_length=8,
                         vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=1)
enc.clip_model.visual.load_state_dict(tiny.visual.state_dict())
enc.preprocess = torchvision.transforms.Normalize((0.48145466*2-1, 0.4578275*2-1, 0.40821073*2-1), (0.26862954*2, 0.26130258*2, 0.27577711*2))
enc.ref = tiny.ref
img = torch.randn(1, 3, 96, 96, generator=torch.Generator().manual_seed(5)).clamp(-1, 1)
a, b = enc.get_gram_matrix_residual(img), tiny.get_gram_matrix_residual(img)
err = ((a - b).norm() / a.norm()).item()
print("GRAM_REL_ERR", err)
assert err < 1e-5, err
""" % (root, os.path.join(root, "tests", "refshim"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GRAM_REL_ERR" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_face_unet_restatement_matches_reference_class():
    """oracle/face_unet.py vs the reference's own `Model` (face-swapping/diffusion/diffusion.py:193) on the same seeded weights, and the face
    golden's intrinsic known answer (the no-reward run of h_Edit_R returns the inverted image).  Subprocess: own `inversion` package."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, importlib, torch
sys.path[:0] = [%r, "/root/reference/face-swapping"]
from oracle.face_unet import FaceUNet, FaceUNetConfig
cfg = FaceUNetConfig.tiny()
m = FaceUNet(cfg)
ref = importlib.import_module("diffusion.diffusion").Model(cfg.as_reference_dict())
ref.load_state_dict(m.state_dict())
x = torch.randn(2, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1)); t = torch.tensor([981.0, 11.0])
with torch.no_grad():
    err = ((m(x, t) - ref(x, t)).norm() / ref(x, t).norm()).item()
print("FACE_REL_ERR", err)
assert err < 1e-5, err
""" % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FACE_REL_ERR" in r.stdout, r.stderr[-2000:]
    g = load_golden("tiny_face_k2")
    assert (g["no_reward"] - g["x0"]).abs().max().item() < 1e-4


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("is_replace,K", [(False, 2), (True, 1)])
def test_compat_loop_and_protocol_controller_match_reference_sampler(is_replace, K):
    """The compat sampler's loop arithmetic (hedit_b200/compat.py) and the user-side protocol controller of the GPU tests
    (tests/protocol_controller.py), both on CPU on the oracle UNet with the REFERENCE's processors installed, against the unmodified
    reference sampler driving the reference's own controller: same latents, same controller bookkeeping."""
    import hedit_b200
    from hedit_b200.compat import h_edit_p2p_implicit_compat
    from oracle.sd_unet import UNetConfig
    from protocol_controller import UserController
    ref = load_reference()
    torch.manual_seed(0)
    T = 4
    cfg = UNetConfig.tiny(sample_size=16)
    model = OraclePipeline(cfg, seed=0)
    model.scheduler.set_timesteps(T)
    prompts = ["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"]
    g = torch.Generator().manual_seed(5)
    xT = torch.randn(1, cfg.in_channels, 16, 16, generator=g)
    zs = torch.randn(T, cfg.in_channels, 16, 16, generator=g)
    kw = dict(cross_replace_steps=0.4, self_replace_steps=0.5, blend_word=None, equilizer_params={"words": ("brown",), "values": (2.0,)},
              num_steps=T, tokenizer=model.tokenizer)
    args = dict(eta=1.0, prompts=prompts, cfg_scales=[1.0, 5.0, 7.5], zs=zs, weight_reconstruction=0.1, optimization_steps=K,
                after_skip_steps=T, is_ddim_inversion=False)

    c_ref = ref.ptp_controller_utils.make_controller(prompts=prompts, is_replace_controller=is_replace, device="cpu", **kw)
    ref.ptp_utils.register_attention_control(model, c_ref)
    ed_ref, rc_ref = ref.p2p_h_edit.h_Edit_p2p_implicit(model, xT=xT, prog_bar=False, controller=c_ref, **args)

    user = UserController(hedit_b200.make_controller(prompts, is_replace, kw["cross_replace_steps"], kw["self_replace_steps"], None,
                                                     kw["equilizer_params"], T, model.tokenizer), "cpu")
    assert hedit_b200.controller_kind(user) == "custom"
    ref.ptp_utils.register_attention_control(model, user)          # the reference's processors now call the user object

    class TorchUNet:                                               # `model.unet` protocol on the oracle UNet
        def __call__(self, sample, t, encoder_hidden_states=None, cross_attention_kwargs=None):
            return model.unet(sample, t, encoder_hidden_states=encoder_hidden_states, cross_attention_kwargs=cross_attention_kwargs)

    ed, rc = h_edit_p2p_implicit_compat(model, xT, controller=user, unet=TorchUNet(), **args)
    assert (ed - ed_ref).abs().max().item() < 2e-5 and (rc - rc_ref).abs().max().item() < 2e-5
    assert (ed_ref - rc_ref).abs().max().item() > 1e-2            # the edit really moved the latent
    assert user.cur_step == c_ref.cur_step == T and user.cur_att_layer == c_ref.cur_att_layer == 0
    assert user.num_att_layers == c_ref.num_att_layers
    for key, items in c_ref.attention_store.items():
        assert len(items) == len(user.attention_store[key])
        for a, b in zip(items, user.attention_store[key]):
            assert torch.allclose(a, b, atol=1e-5)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_compat_masactrl_loop_and_protocol_editor_match_reference_sampler():
    """Editor form of the compat sampler (hedit_b200/compat.py, editor_mode) and the user-side editor of the GPU tests, on CPU on the
    oracle UNet patched by the REFERENCE's `regiter_attention_editor_diffusers`, against the unmodified reference MasaCtrl sampler driving
    the reference's own MutualSelfAttentionControl."""
    import importlib
    import sys
    from hedit_b200.compat import h_edit_masactrl_implicit_compat
    from oracle.sd_unet import UNetConfig
    from protocol_controller import UserMutualSelfAttention
    load_reference()
    masa_pkg = importlib.import_module("masactrl")
    sys.modules.setdefault("masa_ctrl", masa_pkg)                       # reference typo: masactrl.py imports `masa_ctrl`
    mu = importlib.import_module("masactrl.masactrl_utils")
    sys.modules.setdefault("masa_ctrl.masactrl_utils", mu)
    masa = importlib.import_module("masactrl.masactrl")
    mh = importlib.import_module("inversion.masactrl_h_edit")
    T, K, start_step, start_layer = 3, 2, 1, 10
    cfg = UNetConfig.tiny(sample_size=16)
    model = OraclePipeline(cfg, seed=0)
    model.scheduler.set_timesteps(T)
    prompts = ["", "a brown lizard is sitting on a branch"]
    g = torch.Generator().manual_seed(7)
    xT = torch.randn(1, cfg.in_channels, 16, 16, generator=g)
    zs = torch.randn(T, cfg.in_channels, 16, 16, generator=g)
    args = dict(eta=1.0, prompts=prompts, cfg_scales=[1.0, 5.0, 7.5], zs=zs, optimization_steps=K, after_skip_steps=T, is_ddim_inversion=False)

    e_ref = masa.MutualSelfAttentionControl(start_step, start_layer, total_steps=T * K)
    mu.regiter_attention_editor_diffusers(model, e_ref)
    ed_ref, rc_ref = mh.h_Edit_masactrl_implicit(model, xT=xT, prog_bar=False, **args)

    user = UserMutualSelfAttention(start_step, start_layer)
    mu.regiter_attention_editor_diffusers(model, user)               # the reference's patched forwards now call the user object
    assert user.num_att_layers == e_ref.num_att_layers == 32

    class TorchUNet:
        def __call__(self, sample, t, encoder_hidden_states=None, cross_attention_kwargs=None):
            return model.unet(sample, t, encoder_hidden_states=encoder_hidden_states, cross_attention_kwargs=cross_attention_kwargs)

    ed, rc = h_edit_masactrl_implicit_compat(model, xT, editor=user, unet=TorchUNet(), **args)
    assert (ed - ed_ref).abs().max().item() < 2e-5 and (rc - rc_ref).abs().max().item() < 2e-5
    assert user.cur_step == e_ref.cur_step == T * K and user.controlled == (T * K - start_step) * (16 - start_layer)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_irse50_restatement_matches_reference_backbone():
    """hedit_b200/reward_nets.py `IRSE50` (the torch yardstick of the native ArcFace kernels, tests/test_gpu_reward.py) vs the reference's own
    `Backbone(112, 50, mode='ir_se')` (face-swapping/arcface/facial_recognition/model_irse.py:9): identical state_dict keys and shapes --
    the native loader (csrc/reward.cu) consumes exactly these keys -- and bit-identical embeddings on the same seeded weights; and
    `id_features` = `IDLoss.extract_feats` (arcface_model.py:41-47: crop [35:223, 32:220], adaptive pool to 112)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, torch
sys.path[:0] = [%r, "/root/reference/face-swapping"]
from hedit_b200 import reward_nets as R
from arcface.facial_recognition.model_irse import Backbone
ours = R._seed_init(R.IRSE50(), 0)
ref = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode='ir_se')
a, b = ours.state_dict(), ref.state_dict()
assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
ref.load_state_dict(a); ref.eval()
x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(1)).clamp(-1, 1)
pool = torch.nn.AdaptiveAvgPool2d((112, 112))
with torch.no_grad():
    want = ref(pool(x[:, :, 35:223, 32:220]))
    got = R.id_features(ours, x)
print("IRSE_MAX_ABS", (want - got).abs().max().item())
assert torch.equal(want, got)
""" % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "IRSE_MAX_ABS" in r.stdout, r.stderr[-2000:]


def test_lpips_vgg_restatement_matches_torchvision_vgg16_slices():
    """The `lpips` package (0.1.x; not installed offline, no copy under /root/reference) taps torchvision's vgg16().features after modules
    3 / 8 / 15 / 22 / 29 (relu1_2 .. relu5_3).  `reward_nets.LPIPSVGG16` must be that network: same `features.N` keys, same five tensors."""
    tv = pytest.importorskip("torchvision")
    from hedit_b200 import reward_nets as R
    ours = R._seed_init(R.LPIPSVGG16(), 1)
    vgg = tv.models.vgg16(weights=None).features.eval()
    vgg.load_state_dict({k[len("features."):]: v for k, v in ours.state_dict().items() if k.startswith("features.")})
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        h, want = (x - ours.shift) / ours.scale, []
        for i, m in enumerate(vgg):
            h = m(h)
            if i in (3, 8, 15, 22, 29):
                want.append(h)
        got = ours.taps(x)
    assert len(got) == 5 and all(torch.equal(a, b) for a, b in zip(got, want))
