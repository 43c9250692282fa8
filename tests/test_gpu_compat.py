"""GPU: the compat path -- attention probabilities materialised in fp32 and handed to ANY controller object that
follows the reference's call protocol.  Checked (a) layer by layer against the fused kernels with a controller that edits nothing and
(b) end to end: a user-side controller (tests/protocol_controller.py, not a stock class) driven through the public sampler must
reproduce what the UNMODIFIED reference produced with its own controllers (tests/golden/, tests/make_golden.py)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.pipeline import OraclePipeline  # noqa: E402
from oracle_run import cfg_from_meta, load_golden  # noqa: E402
from protocol_controller import UserController, UserMutualSelfAttention  # noqa: E402

import hedit_b200  # noqa: E402
from gpu_util import rel_err  # noqa: E402

TOL_LOOP = 2.5e-2


def _model(meta, cache={}):
    cfg = cfg_from_meta(meta)
    key = (tuple(cfg.block_out_channels), cfg.sample_size)
    if key not in cache:
        cache[key] = OraclePipeline(cfg, seed=0)
    model = cache[key]
    model.scheduler.set_timesteps(meta["T"])
    return model


def test_compat_forward_equals_fused_forward():
    g = load_golden("tiny_refine_noblend")
    model = _model(g["meta"])
    eng = hedit_b200.get_engine(model, max_samples=5)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(4, 4, 64, 64, generator=gen).cuda()
    ctx = torch.cat([g["ctx_uncond"], g["ctx_uncond"], g["ctx_src"], g["ctx_tar"]]).cuda()
    seen = []

    def hook(layer, is_cross, place, probs):
        seen.append((layer, is_cross, place, tuple(probs.shape)))
        s = probs.sum(-1)
        assert torch.allclose(s, torch.ones_like(s), atol=1e-4)       # rows are softmax outputs

    fused = eng.forward(x, 481.0, ctx)
    compat = eng.forward_compat(x, 481.0, ctx, hook)
    r, m = rel_err(compat, fused)
    print(f"compat vs fused forward: rel {r:.3e} max {m:.3e}; {len(seen)} hook calls, launches {eng.last_stats['kernel_launches']}")
    assert r < 3e-3
    nb = eng.n_transformer_blocks()
    assert len(seen) == 2 * nb == 32
    assert [s[0] for s in seen] == [i // 2 for i in range(2 * nb)] and [s[1] for s in seen] == [False, True] * nb
    assert [s[2] for s in seen[::2]] == [0] * 6 + [1] + [2] * 9
    assert seen[0][3] == (4 * 8, 4096, 4096) and seen[1][3] == (4 * 8, 4096, 77)

    def boom(*_):
        raise ValueError("user hook failed")
    with pytest.raises(ValueError, match="user hook failed"):
        eng.forward_compat(x, 481.0, ctx, boom)
    assert rel_err(eng.forward(x, 481.0, ctx), fused)[0] == 0.0        # the engine is still usable, fused path unchanged


@pytest.mark.parametrize("name", ["tiny_refine_noblend", "tiny_refine_blend", "tiny_replace_mos2"])
def test_custom_controller_through_sampler_matches_reference_golden(name):
    g = load_golden(name)
    meta = g["meta"]
    model = _model(meta)
    bw = meta["blend_words"]
    tables = hedit_b200.make_controller(
        meta["prompts"], meta["is_replace"], meta["xa"], meta["sa"],
        blend_word=((bw[0],), (bw[1],)) if meta["blend"] else None,
        equilizer_params={"words": (bw[1],), "values": (1.25 if meta["K"] > 1 else 2.0,)} if meta["blend"] else None,
        num_steps=meta["T"], tokenizer=model.tokenizer)
    ctrl = UserController(tables, "cuda")
    assert hedit_b200.controller_kind(ctrl) == "custom" and hedit_b200.controller_kind(tables) == "stock"
    ed, rc = hedit_b200.h_Edit_p2p_implicit(model, g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"], cfg_scales=meta["cfg_scales"],
                                            zs=g["zs"].cuda(), controller=ctrl, weight_reconstruction=meta["weight_reconstruction"],
                                            optimization_steps=meta["K"], after_skip_steps=meta["T"], is_ddim_inversion=False)
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    r_rc, m_rc = rel_err(rc.cpu(), g["recon"])
    print(f"{name} via compat path: edited rel {r_ed:.3e} max {m_ed:.3e} | recon rel {r_rc:.3e} | controller calls {ctrl.calls}, cur_step {ctrl.cur_step}")
    assert r_rc < TOL_LOOP and r_ed < TOL_LOOP
    # the controller's own bookkeeping advanced exactly as under the reference: one step per timestep, 32 calls per controlled launch
    assert ctrl.cur_step == meta["T"] and ctrl.cur_att_layer == 0
    assert ctrl.calls == 32 * meta["T"] * meta["K"]
    assert len(ctrl.attention_store["down_cross"]) == 4 and len(ctrl.attention_store["up_cross"]) == 6


def test_custom_editor_through_masactrl_sampler_matches_reference_golden():
    """MasaCtrl's editor protocol on the compat path: a user-side editor (own class) registered with the reference's hook name and driven
    through `h_Edit_masactrl_implicit` must reproduce the golden of the unmodified reference sampler + reference editor."""
    g = load_golden("tiny_masactrl_mos2")
    meta = g["meta"]
    model = _model(meta)
    T, K = meta["T"], meta["K"]
    editor = UserMutualSelfAttention(meta["masa_start_step"], meta["masa_start_layer"])
    hedit_b200.regiter_attention_editor_diffusers(model, editor)
    ed, rc = hedit_b200.h_Edit_masactrl_implicit(model, g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"], cfg_scales=meta["cfg_scales"],
                                                 zs=g["zs"].cuda(), optimization_steps=K, after_skip_steps=T, is_ddim_inversion=False)
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    r_rc, m_rc = rel_err(rc.cpu(), g["recon"])
    print(f"tiny_masactrl_mos2 via the editor compat path: edited rel {r_ed:.3e} max {m_ed:.3e} | recon rel {r_rc:.3e} | calls {editor.calls}, "
          f"controlled {editor.controlled}, cur_step {editor.cur_step}")
    assert r_rc < TOL_LOOP and r_ed < TOL_LOOP
    assert editor.calls == 32 * T * K and editor.cur_step == T * K and editor.cur_att_layer == 0
    assert editor.controlled == (T * K - meta["masa_start_step"]) * (16 - meta["masa_start_layer"])

    # the plain editor protocol (attn @ v) must reproduce the fused forward
    eng = hedit_b200.get_engine(model, max_samples=5)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 4, 64, 64, generator=gen).cuda()
    ctx = torch.cat([g["ctx_src"], g["ctx_tar"]]).cuda()
    plain = lambda layer, is_cross, place, q, k, v, sim, attn, heads: UserMutualSelfAttention._merge(torch.bmm(torch.softmax(sim, -1), v), heads)
    r, _ = rel_err(eng.forward_editor(x, 301.0, ctx, plain), eng.forward(x, 301.0, ctx))
    assert r < 3e-3, r


def test_attention_store_is_materialised_lazily_after_a_fused_edit():
    """after an edit on the FUSED path the stock controller's `attention_store` / `get_average_attention()` must hold
    what the reference's AttentionStore would (ptp_classes.py:135-160).  The fused kernels never write the maps, so the store fills
    itself on first access by replaying the edit through the compat path; it must equal the store a user-side protocol controller
    accumulates when it drives the same edit itself."""
    g = load_golden("tiny_refine_blend")
    meta = g["meta"]
    model = _model(meta)
    bw = meta["blend_words"]
    mk = lambda: hedit_b200.make_controller(meta["prompts"], False, meta["xa"], meta["sa"], blend_word=((bw[0],), (bw[1],)),
                                            equilizer_params={"words": (bw[1],), "values": (2.0,)}, num_steps=meta["T"], tokenizer=model.tokenizer)
    kw = dict(eta=meta["eta"], prompts=meta["prompts"], cfg_scales=meta["cfg_scales"], zs=g["zs"].cuda(), weight_reconstruction=meta["weight_reconstruction"],
              optimization_steps=meta["K"], after_skip_steps=meta["T"], is_ddim_inversion=False)
    stock = mk()
    assert hedit_b200.controller_kind(stock) == "stock"
    ed, rc = hedit_b200.h_Edit_p2p_implicit(model, g["xT"].cuda(), controller=stock, **kw)
    assert hedit_b200.get_engine(model).last_stats["sample_forwards"] == 7 * meta["T"]        # ran fused (exact-reuse schedule)
    assert stock.cur_step == meta["T"]
    user = UserController(mk(), "cuda")
    hedit_b200.h_Edit_p2p_implicit(model, g["xT"].cuda(), controller=user, **kw)
    store = stock.attention_store                       # first access: replay through the compat path
    assert set(store.keys()) == set(user.attention_store.keys())
    for key, items in user.attention_store.items():
        assert len(store[key]) == len(items), key
        for a, b in zip(store[key], items):
            assert a.shape == b.shape and rel_err(a, b)[0] < 1e-5, key
    avg = stock.get_average_attention()
    assert rel_err(avg["down_cross"][0], user.attention_store["down_cross"][0] / meta["T"])[0] < 1e-5
    assert len(store["down_cross"]) == 4 and len(store["up_cross"]) == 6 and len(store["mid_self"]) == 1


def test_bare_attention_store_passed_to_the_p2p_sampler_is_filled():
    """A passive store handed to h_Edit_p2p_implicit (no edit tables) must not be silently ignored: it is served by the compat path and
    ends up holding the maps (the reference fills it in call C, p2p_h_edit.py:652)."""
    g = load_golden("tiny_refine_noblend")
    meta = g["meta"]
    model = _model(meta)

    class AttentionStore:               # user-side passive recorder with the reference's class name and protocol
        def __init__(self):
            self.num_att_layers, self.cur_att_layer, self.cur_step, self.maps = -1, 0, 0, 0

        def __call__(self, attn, is_cross, place, save_attn):
            self.maps += int(save_attn and attn.shape[1] <= 32 ** 2)
            self.cur_att_layer += 1
            if self.cur_att_layer == self.num_att_layers:
                self.cur_att_layer, self.cur_step = 0, self.cur_step + 1
            return attn

        def step_callback(self, x):
            return x

    store = AttentionStore()
    ed, rc = hedit_b200.h_Edit_p2p_implicit(model, g["xT"].cuda(), eta=meta["eta"], prompts=meta["prompts"], cfg_scales=meta["cfg_scales"], zs=g["zs"].cuda(),
                                            controller=store, weight_reconstruction=meta["weight_reconstruction"], optimization_steps=1,
                                            after_skip_steps=meta["T"], is_ddim_inversion=False)
    assert store.cur_step == meta["T"] and store.maps > 0
    assert torch.isfinite(ed).all() and rel_err(rc.cpu(), g["recon"])[0] < TOL_LOOP
