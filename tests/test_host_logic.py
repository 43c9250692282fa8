"""CPU: host-side logic of the product package (no GPU, no compute calls into the library)."""
import os
import re

import numpy as np
import pytest
import torch

import hedit_b200
from hedit_b200 import _lib, p2p
from oracle import h_edit as oh
from oracle import p2p as op
from oracle.pipeline import DDIMSchedulerTables, OraclePipeline, ToyTokenizer
from refload import load_reference, reference_available
from test_oracle_pin import PAIRS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hedit_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hedit_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hedit_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and header disagree"


def test_engine_creation_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        hedit_b200.UNetEngine(dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(64, 128, 256, 256),
                                   layers_per_block=2, heads=8, cross_attention_dim=64, norm_groups=32, ctx_len=77), max_samples=2)


@pytest.mark.parametrize("T,skip,eta,ddim", [(50, 0, 1.0, False), (10, 0, 1.0, False), (50, 15, 1.0, False), (20, 0, 1.0, True), (10, 2, 0.6, False)])
def test_step_tables_match_oracle_scheduler_algebra(T, skip, eta, ddim):
    """schedule.step_tables reproduces reverse_step / compute_full_coeff (oracle restatement of inversion_utils.py)."""
    sched = DDIMSchedulerTables(steps_offset=0 if ddim else 1)
    sched.set_timesteps(T)
    S = T - skip
    ts, coef = hedit_b200.step_tables(sched, S, eta, ddim)
    op_ts = [int(t) for t in sched.timesteps[-S:]]
    assert ts == op_ts + [0]
    g = torch.Generator().manual_seed(0)
    for i, t in enumerate(op_ts):
        x, eps, z = (torch.randn(2, 4, 8, 8, generator=g) for _ in range(3))
        ref = oh.reverse_step(sched, eps, t, x, eta, z, ddim)
        k = coef[i]
        x0 = (x - k[0] * eps) / k[1]
        mine = (k[2] * x0 + k[3] * eps) + k[4] * z
        assert (mine - ref).abs().max().item() < 2e-5
        tt = ts[i + 1]
        ab = sched.alphas_cumprod
        c_ref = oh.full_coeff(sched, t, tt, eta, ddim) - (1 - ab[t]) ** 0.5 * (ab[tt] ** 0.5 / ab[t] ** 0.5)
        assert abs(float(c_ref) - float(k[5])) < 1e-6


@pytest.mark.parametrize("prompts,bw_src,bw_tar", PAIRS)
@pytest.mark.parametrize("is_replace", [False, True])
def test_product_controller_tables_match_oracle(prompts, bw_src, bw_tar, is_replace):
    tok = ToyTokenizer()
    if is_replace and len(prompts[0].split(" ")) != len(prompts[1].split(" ")):
        pytest.skip("replace controller needs equal word counts")
    T = 10
    eqp = {"words": (bw_tar,), "values": (2.0,)}
    c = hedit_b200.make_controller(prompts, is_replace, 0.4, 0.35, blend_word=((bw_src,), (bw_tar,)), equilizer_params=eqp, num_steps=T, tokenizer=tok)
    s = op.make_edit_spec(prompts, is_replace, 0.4, 0.35, ((bw_src,), (bw_tar,)), eqp, T, tok)
    assert torch.equal(c.cross_replace_alpha.reshape(T + 1, 77), s.alpha_words)
    assert tuple(c.num_self_replace) == tuple(s.self_window)
    assert torch.equal(c.equalizer.reshape(77), s.equalizer)
    assert torch.equal(c.local_blend.alpha_layers.reshape(2, 77), s.blend_alpha)
    if is_replace:
        assert torch.equal(c.mapper[0], s.replace_matrix)
    else:
        assert torch.equal(c.mapper[0], s.mapper) and torch.equal(c.alphas.reshape(77), s.refine_alpha)
    # compiled plan == the oracle's edit formula evaluated on random probabilities
    plan = hedit_b200.compile_edit_plan([c], T)
    g = torch.Generator().manual_seed(1)
    base, tar = torch.rand(3, 5, 77, generator=g), torch.rand(3, 5, 77, generator=g)
    for step in (0, 3, 4, 9):
        aw = s.alpha_words[step]
        want = op._mapped_base(s, base, tar) * aw + (1 - aw) * tar
        cb, ct = torch.from_numpy(plan.c_base[step, 0, :77]), torch.from_numpy(plan.c_tar[step, 0, :77])
        if is_replace:
            mapped = torch.einsum("hpw,wn->hpn", base, torch.from_numpy(plan.replace_m[0, :, :77]))
        else:
            mapped = base[:, :, torch.from_numpy(plan.mapper[0, :77]).long()]
        got = mapped * cb + tar * ct
        assert (got - want).abs().max().item() < 1e-6
    assert plan.has_blend.tolist() == [1] and plan.start_blend == s.start_blend


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_compile_edit_plan_accepts_reference_controller_objects():
    ref = load_reference()
    tok = ToyTokenizer()
    prompts = PAIRS[0][0]
    T = 8
    kw = dict(cross_replace_steps=0.4, self_replace_steps=0.35, blend_word=(("lizard",), ("lizard",)),
              equilizer_params={"words": ("lizard",), "values": (2.0,)}, num_steps=T, tokenizer=tok)
    c_ref = ref.ptp_controller_utils.make_controller(prompts=prompts, is_replace_controller=False, device="cpu", **kw)
    c_own = hedit_b200.make_controller(prompts, False, **kw)
    a, b = hedit_b200.compile_edit_plan([c_ref], T), hedit_b200.compile_edit_plan([c_own], T)
    for f in ("mapper", "is_replace", "c_base", "c_tar", "has_blend", "blend_alpha"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.self_window == b.self_window and a.start_blend == b.start_blend and a.blend_th == b.blend_th


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_controller_kind_routes_reference_classes():
    """Stock reference controllers (Reweight -> Refine chain, Replace) compile to the fused path; a bare AttentionStore / EmptyControl
    carries no edit; a user subclass that overrides the cross replacement must go through the compat path."""
    ref = load_reference()
    tok = ToyTokenizer()
    prompts = PAIRS[0][0]
    kw = dict(cross_replace_steps=0.4, self_replace_steps=0.35, num_steps=8, tokenizer=tok)
    chain = ref.ptp_controller_utils.make_controller(prompts=prompts, is_replace_controller=False, device="cpu", blend_word=(("lizard",), ("lizard",)),
                                                     equilizer_params={"words": ("lizard",), "values": (2.0,)}, **kw)
    assert type(chain).__name__ == "AttentionReweight" and hedit_b200.controller_kind(chain) == "stock"
    rep = ref.ptp_controller_utils.make_controller(prompts=prompts, is_replace_controller=True, device="cpu", blend_word=None, equilizer_params=None, **kw)
    assert hedit_b200.controller_kind(rep) == "stock"
    assert hedit_b200.controller_kind(ref.ptp_classes.AttentionStore()) == "none"
    assert hedit_b200.controller_kind(ref.ptp_classes.EmptyControl()) == "none"
    assert hedit_b200.controller_kind(None) == "none"

    class HalfStrength(ref.ptp_classes.AttentionRefine):
        def replace_cross_attention(self, attn_base, att_replace):
            return 0.5 * super().replace_cross_attention(attn_base, att_replace) + 0.5 * att_replace

    user = HalfStrength(prompts, 8, cross_replace_steps=0.4, self_replace_steps=0.35, tokenizer=tok, device="cpu")
    assert hedit_b200.controller_kind(user) == "custom"
    assert hedit_b200.controller_kind(lambda probs, is_cross, place, save_attn: probs) == "custom"


def test_batched_plan_stacks_images():
    tok = ToyTokenizer()
    ctrls = [hedit_b200.make_controller(p, False, 0.4, 0.35, blend_word=((s,), (t,)) if i % 2 == 0 else None,
                                        equilizer_params=None, num_steps=6, tokenizer=tok) for i, (p, s, t) in enumerate(PAIRS)]
    plan = hedit_b200.compile_edit_plan(ctrls, 6)
    assert plan.mapper.shape == (4, 80) and plan.c_base.shape == (7, 4, 80) and plan.has_blend.tolist() == [1, 0, 1, 0]
    with pytest.raises(ValueError):
        hedit_b200.compile_edit_plan(ctrls, 7)


# ---- host-side tables of the samplers added after the north-star loop --------------------------------------------------------------
def test_other_engines_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        hedit_b200.VaeDecoderEngine(dict(latent_channels=4, out_channels=3, block_out_channels=(64, 64, 128, 128), layers_per_block=2, norm_groups=32))
    with pytest.raises(RuntimeError):
        hedit_b200.ClipGramEngine(224, 16, 64, 1)
    with pytest.raises(RuntimeError):
        hedit_b200.FaceUNetEngine(dict(ch=64, ch_mult=(1, 2, 2), num_res_blocks=2, attn_resolution=16, image_size=64, in_channels=3, out_ch=3))


def test_pnp_flags_and_layer_mask():
    """Plug-and-Play: the injected pair call of step i runs at tt = op[i+1] (0 after the last step) and the patched forwards test
    `t in schedule` (plug_n_play/pnp_utils.py:51,138); layers = decoder self-attention blocks 4-11 (pnp_utils.py:88)."""
    model = OraclePipeline(build_unet=False)
    T = 10
    model.scheduler.set_timesteps(T)
    ts = [int(t) for t in model.scheduler.timesteps]
    hedit_b200.register_attention_control_efficient(model, ts[:4])
    hedit_b200.register_conv_control_efficient(model, ts[:7])
    qk, ft = hedit_b200.pnp_step_flags(model, T)
    assert qk == [1, 1, 1, 0, 0, 0, 0, 0, 0, 0] and ft == [1, 1, 1, 1, 1, 1, 0, 0, 0, 0]
    qk, ft = hedit_b200.pnp_step_flags(model, T - 3)            # skipped schedule: op = timesteps[-7:]
    assert qk == [0] * 7 and ft == [1, 1, 1, 0, 0, 0, 0]
    assert hedit_b200.pnp_self_mask(2) == 0xFF00
    hedit_b200.register_attention_control_efficient(model, None)
    assert hedit_b200.pnp_step_flags(model, T)[0] == [0] * T


def test_skip_pre_coeff_and_x0_tables_follow_the_reference_formulas():
    model = OraclePipeline(build_unet=False)
    T, S = 20, 14
    model.scheduler.set_timesteps(T)
    sched = model.scheduler
    assert hedit_b200.skip_pre_coeff(sched, T, 1.0) is None
    ta, t = int(sched.timesteps[-(S + 1)]), int(sched.timesteps[-S])
    want = oh.full_coeff(sched, ta, t, 1.0, False) - (1 - sched.alphas_cumprod[ta]) ** 0.5 * (sched.alphas_cumprod[t] ** 0.5 / sched.alphas_cumprod[ta] ** 0.5)
    assert abs(hedit_b200.skip_pre_coeff(sched, S, 1.0) - float(want)) < 1e-7
    x0c = hedit_b200.x0_tables(sched, S)
    tts = [int(v) for v in sched.timesteps[-S:]][1:] + [0]
    for (a, b), tt in zip(x0c, tts):
        assert abs(a - float((1 - sched.alphas_cumprod[tt]) ** 0.5)) < 1e-7 and abs(b - float(sched.alphas_cumprod[tt] ** 0.5)) < 1e-7


def test_face_step_tables_follow_the_reference_formulas():
    """face-swapping/inversion/h_edit_R.py:68-88,106 with the beta schedule of main_edit.py:130-142."""
    import numpy as np
    T = 10
    betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64).float()
    seq = (np.arange(0, 1000, 1000 // T) + 1)[::-1]
    tab = hedit_b200.face.face_step_tables(betas, seq, T, T - 2, eta=1.0)
    ab = (1 - betas).cumprod(0)
    op = [int(v) for v in seq[-(T - 2):]]
    assert tab.shape == (T - 2, 8)
    for i, t in enumerate(op):
        tm1 = op[i + 1] if i + 1 < len(op) else 0
        c1 = (1 - ab[tm1]).sqrt() * 0.5
        c2 = (1 - ab[tm1]).sqrt() * ((1 - 0.25) ** 0.5)
        want = [t, tm1, (1 - ab[t]) ** 0.5, ab[t] ** 0.5, (1 - ab[tm1]) ** 0.5, ab[tm1].sqrt(), c2, 1.0 * c1]
        assert np.allclose(tab[i], np.asarray([float(v) for v in want], dtype=np.float32), rtol=1e-6, atol=0)


def test_compat_unet_routes_cross_attention_kwargs_like_the_reference_processors():
    """`CompatUNet` must honour the reference's kwargs: P2P processors run the controller unless `use_controller` is False and pass
    `save_attn` through (p2p/ptp_utils.py:38-46, 102-104); MasaCtrl's patched forward calls the editor unless `use_editor` is False
    (masactrl/masactrl_utils.py:40, 72-78).  Checked on a stand-in engine that records which native entry point would run."""
    from hedit_b200.compat import CompatUNet, PLACES

    class FakeEngine:
        config = {"in_channels": 4, "sample_size": 64}

        def __init__(self):
            self.log = []

        def forward(self, x, t, ctx):
            self.log.append(("fused", t))
            return x

        def forward_compat(self, x, t, ctx, hook):
            self.log.append(("probs", t))
            hook(3, True, 2, torch.zeros(16, 4, 77))
            return x

        def forward_editor(self, x, t, ctx, hook):
            self.log.append(("editor", t))
            out = hook(0, False, 0, torch.zeros(16, 4, 8), torch.zeros(16, 4, 8), torch.zeros(16, 4, 8), torch.zeros(16, 4, 4), torch.zeros(16, 4, 4), 8)
            assert out.shape == (2, 4, 64)
            return x

    calls = []
    ctrl = lambda probs, is_cross, place, save_attn: calls.append((tuple(probs.shape), is_cross, place, save_attn))
    x, ctx = torch.zeros(2, 4, 64, 64), torch.zeros(2, 77, 8)
    eng = FakeEngine()
    unet = CompatUNet(eng, controller=ctrl)
    assert unet(x, torch.tensor(481), encoder_hidden_states=ctx, cross_attention_kwargs={"use_controller": False}).sample is x
    unet(x, 461, encoder_hidden_states=ctx, cross_attention_kwargs={"save_attn": False})
    unet(x, 441, encoder_hidden_states=ctx)
    assert [k for k, _ in eng.log] == ["fused", "probs", "probs"]
    assert calls == [((16, 4, 77), True, PLACES[2], False), ((16, 4, 77), True, "up", True)]
    assert CompatUNet(FakeEngine(), controller=None)(x, 1, encoder_hidden_states=ctx).sample is x

    seen = []

    def editor(q, k, v, sim, attn, is_cross, place, heads, scale=None):
        seen.append((is_cross, place, heads, scale))
        return torch.zeros(q.shape[0] // heads, q.shape[1], heads * q.shape[2])

    eng2 = FakeEngine()
    unet2 = CompatUNet(eng2, editor=editor)
    unet2(x, 481, encoder_hidden_states=ctx, cross_attention_kwargs={"use_editor": False})
    unet2(x, 461, encoder_hidden_states=ctx)
    assert [k for k, _ in eng2.log] == ["fused", "editor"]
    assert seen == [(False, "down", 8, 8 ** -0.5)]


def test_controller_kind_checks_the_defining_module_not_only_the_name():
    """A user class that merely reuses a stock class NAME (or overrides a hook in a subclass) only promises the protocol."""
    tok = ToyTokenizer()
    own = hedit_b200.make_controller(PAIRS[0][0], False, 0.4, 0.35, num_steps=8, tokenizer=tok)
    assert hedit_b200.controller_kind(own) == "stock"

    class AttentionRefine:          # same name as the reference's class, defined elsewhere
        cross_replace_alpha = own.cross_replace_alpha
        prev_controller = None

        def __call__(self, attn, is_cross, place, save_attn):
            return attn

    assert hedit_b200.controller_kind(AttentionRefine()) == "custom"

    class AttentionStore:
        pass
    assert hedit_b200.controller_kind(AttentionStore()) == "custom"      # a user-side recorder: must see its maps (compat path)


def test_masactrl_launch_plan_follows_the_reference_schedule():
    """masactrl.py:33-36,57: active iff cur_step in step_idx and block in layer_idx; step_idx defaults to range(start, total_steps)."""
    ed = hedit_b200.MutualSelfAttentionControl(4, 10, total_steps=50)
    mask, on = ed.launch_plan(70)                     # 35 timesteps x K = 2 controlled launches
    assert mask == sum(1 << l for l in range(10, 16))
    assert on == [0] * 4 + [1] * 46 + [0] * 20        # the injection ENDS at total_steps
    ed.cur_step = 48
    assert ed.launch_plan(4)[1] == [1, 1, 0, 0]
    ed2 = hedit_b200.MutualSelfAttentionControl(layer_idx=[3, 8, 15], step_idx=[1, 2, 4], total_steps=6)
    mask, on = ed2.launch_plan(6)
    assert mask == (1 << 3) | (1 << 8) | (1 << 15) and on == [0, 1, 1, 0, 1, 0]


def test_substruct_words_and_sparse_replacement_plan():
    tok = ToyTokenizer()
    c = hedit_b200.make_controller(PAIRS[0][0], False, 0.4, 0.35, blend_word=(("lizard",), ("lizard",)), substruct_words=(("branch",), ("branch",)),
                                   num_steps=8, tokenizer=tok)
    plan = hedit_b200.compile_edit_plan([c], 8)
    assert plan.blend_alpha.shape == (1, 4, 80) and plan.blend_alpha[0, 2:].sum() == 2 and plan.blend_alpha[0, :2].sum() == 2
    assert not np.array_equal(plan.blend_alpha[0, 0], plan.blend_alpha[0, 2])
    # replacement mapper: the sparse (index, weight) rows reproduce the dense 77x77 product exactly
    rep = hedit_b200.make_controller(PAIRS[3][0], True, 0.4, 0.35, num_steps=8, tokenizer=tok)
    plan = hedit_b200.compile_edit_plan([rep, c], 8)
    assert plan.map_w is not None and plan.mapper.shape == plan.map_w.shape and plan.mapper.shape[1] <= 4
    P = np.random.default_rng(0).random(77).astype(np.float32)
    dense = P @ plan.replace_m[0, :, :77]
    sparse = sum(P[plan.mapper[0, k, :77]] * plan.map_w[0, k, :77] for k in range(plan.map_w.shape[1]))
    assert np.array_equal(dense.astype(np.float32), sparse.astype(np.float32))
    # the Refine image rides along as (mapper[j], 1)
    single = hedit_b200.compile_edit_plan([c], 8)
    assert np.array_equal(plan.mapper[1, 0], single.mapper[0]) and (plan.map_w[1, 0, :77] == 1).all() and (plan.map_w[1, 1:] == 0).all()


def test_edit_controller_implements_the_reference_protocol_and_lazy_store():
    """EditController is itself a protocol controller (used by the compat replay): counters, store, cross edit on materialised maps."""
    from hedit_b200.p2p import LazyAttentionStore
    tok = ToyTokenizer()
    c = hedit_b200.make_controller(PAIRS[0][0], False, 0.4, 0.35, num_steps=4, tokenizer=tok)
    c.num_att_layers = 2
    g = torch.Generator().manual_seed(0)
    cross = torch.softmax(torch.randn(4 * 8, 256, 77, generator=g), -1)
    before = cross.clone()
    c(cross, True, "down", True)
    assert torch.equal(cross[:24], before[:24]) and not torch.equal(cross[24:], before[24:])      # only the cond-target rows change
    c(torch.softmax(torch.randn(32, 256, 256, generator=g), -1), False, "down", True)
    assert c.cur_step == 1 and c.cur_att_layer == 0
    assert len(c.attention_store["down_cross"]) == 1 and c.attention_store["down_cross"][0].shape == (16, 256, 77)
    assert torch.equal(c.attention_store["down_cross"][0], cross[16:])                              # the stored map is the EDITED one
    calls = []
    lazy = LazyAttentionStore(lambda: calls.append(1) or {"down_cross": [torch.ones(1)]})
    assert calls == []
    assert len(lazy["down_cross"]) == 1 and "down_cross" in lazy and len(lazy) == 1 and calls == [1]


def test_ctypes_struct_layouts_match_the_library():
    """The ctypes mirrors of the C-ABI structs (hedit_b200/_lib.py) have exactly the sizes the loaded library was compiled with
    (`hedit_abi_sizeof`): a field added on one side only would shift every later field of hedit_edit_args silently."""
    import ctypes as C
    from hedit_b200 import _lib
    lib = _lib.load()
    pairs = {"hedit_edit_args": _lib.EditArgsC, "hedit_step_coef": _lib.StepCoefC, "hedit_unet_config": _lib.UNetConfigC,
             "hedit_vae_config": _lib.VaeConfigC, "hedit_clip_config": _lib.ClipConfigC, "hedit_text_config": _lib.TextConfigC,
             "hedit_face_config": _lib.FaceConfigC, "hedit_face_step_coef": _lib.FaceStepCoefC, "hedit_face_args": _lib.FaceArgsC}
    for name, cls in pairs.items():
        assert lib.hedit_abi_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.hedit_abi_sizeof(b"no_such_struct") == -1


def test_reward_module_recognition_and_name_mapping():
    """Host side of the native reward networks (hedit_b200/reward.py): the IR-SE50 state_dict test, and the LPIPS tensor-name mapping for both
    key layouts (the lpips package's `net.slice3.12.weight` / `lin2.model.1.weight` / `scaling_layer.shift` and reward_nets' own)."""
    import torch
    from hedit_b200 import reward, reward_nets
    irse = reward_nets.IRSE50()
    assert reward.ArcFaceEngine.is_irse50_state_dict(irse.state_dict())
    sd = dict(irse.state_dict()); sd.pop("body.23.res_layer.5.fc2.weight")
    assert not reward.ArcFaceEngine.is_irse50_state_dict(sd)                         # another depth / mode: not IR-SE50
    net = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 1)
    ours = reward.lpips_native_tensors(net.state_dict())
    want = {f"conv{i}.{k}" for i in range(13) for k in ("weight", "bias")} | {f"lin{k}.weight" for k in range(5)} | {"shift", "scale"}
    assert set(ours) == want
    slices = {0: 1, 2: 1, 5: 2, 7: 2, 10: 3, 12: 3, 14: 3, 17: 4, 19: 4, 21: 4, 24: 5, 26: 5, 28: 5}
    pkg = {}
    for k, v in net.state_dict().items():
        if k.startswith("features."):
            idx = int(k.split(".")[1])
            pkg[f"net.slice{slices[idx]}.{idx}.{k.split('.')[2]}"] = v
        elif k.startswith("lins."):
            pkg[f"lin{k.split('.')[1]}.model.1.weight"] = v
            pkg[f"lins.{k.split('.')[1]}.model.1.weight"] = v
        else:
            pkg[f"scaling_layer.{k}"] = v
    theirs = reward.lpips_native_tensors(pkg)
    assert set(theirs) == want and all(torch.equal(theirs[k].reshape(-1), ours[k].reshape(-1)) for k in want)
    # seeded construction is reproducible, incl. the `lin` heads (drawn from the global RNG by the constructor)
    again = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 1)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), again.state_dict().values()))
    # a reward object whose loss method is overridden is NOT taken over by the native path (it keeps the autograd plug-in route)
    class MyID(reward_nets.SyntheticIDLoss):
        def get_cosine_loss(self, image):
            return super().get_cosine_loss(image) * 2
    assert reward._stock_class(MyID.__new__(MyID), "SyntheticIDLoss") is reward_nets.SyntheticIDLoss


def test_pnp_flags_at_current_timestep():
    """Plug-and-Play baselines inject at the CURRENT timestep (pnp_baselines.py:367), h_Edit_PnP_implicit at the previous one."""
    import types
    import hedit_b200
    model = types.SimpleNamespace(scheduler=types.SimpleNamespace(timesteps=torch.tensor([801, 601, 401, 201, 1])))
    hedit_b200.register_attention_control_efficient(model, [801, 601])
    hedit_b200.register_conv_control_efficient(model, [801, 601, 401])
    from hedit_b200.samplers import pnp_step_flags_at_t
    assert pnp_step_flags_at_t(model, 5) == ([1, 1, 0, 0, 0], [1, 1, 1, 0, 0])
    assert hedit_b200.pnp_step_flags(model, 5) == ([1, 0, 0, 0, 0], [1, 1, 0, 0, 0])
    assert pnp_step_flags_at_t(model, 3) == ([0, 0, 0], [1, 0, 0])
