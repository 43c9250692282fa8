"""GPU: face-swapping path -- the native pixel-space DDPM UNet against the oracle restatement of the reference's
`Model` (pinned to the reference class on CPU), and the native `h_Edit_R` loop against outputs of the UNMODIFIED reference sampler
(tests/make_golden.py --config face)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.face_unet import FaceUNet, FaceUNetConfig, TinyIDLoss, TinyLPIPSLoss  # noqa: E402
from oracle_run import load_golden  # noqa: E402

import hedit_b200  # noqa: E402
from gpu_util import rel_err  # noqa: E402

TOL_UNET = 5e-3
TOL_LOOP = 4e-2


def _fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("cfg,S", [(FaceUNetConfig.tiny(), 3), (FaceUNetConfig(ch=64, ch_mult=(1, 1, 2, 2), image_size=256, attn_resolutions=(32,)), 1)])
def test_face_unet_forward_matches_torch(cfg, S):
    _fp32()
    model = FaceUNet(cfg).cuda()
    eng = hedit_b200.FaceUNetEngine.from_model(model)
    g = torch.Generator(device="cpu").manual_seed(4)
    x = torch.randn(S, 3, cfg.image_size, cfg.image_size, generator=g).cuda()
    t = torch.tensor([981.0, 401.0, 1.0][:S]).cuda()
    with torch.no_grad():
        ref = model(x, t)
    out = eng(x, t)
    r, m = rel_err(out, ref)
    print(f"face unet {cfg.image_size}px ch={cfg.ch} mult={cfg.ch_mult}: rel {r:.3e} max {m:.3e} launches {eng.last_stats['kernel_launches']}")
    assert r < TOL_UNET


def test_face_unet_graph_replay_is_bit_identical():
    """Same (x, out) buffers: 1st call launches directly, 2nd is captured, 3rd+ are CUDA-graph replays (netexec.h); a changed time step
    and changed input VALUES must flow through the replay, and all three ways must agree bitwise with direct launches on new buffers."""
    cfg = FaceUNetConfig.tiny()
    eng = hedit_b200.FaceUNetEngine.from_model(FaceUNet(cfg).cuda())
    g = torch.Generator(device="cpu").manual_seed(6)
    xa = torch.randn(2, 3, cfg.image_size, cfg.image_size, generator=g).cuda()
    xb = torch.randn(2, 3, cfg.image_size, cfg.image_size, generator=g).cuda()
    x, out = xa.clone(), torch.empty_like(xa)
    outs = [eng(x, 500.0, out=out).clone() for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    x.copy_(xb)
    replayed = eng(x, torch.tensor([37.0, 912.0]), out=out).clone()          # replay with new values and time steps
    direct = eng(xb.clone(), torch.tensor([37.0, 912.0]))                    # new buffers -> direct launches
    assert torch.equal(replayed, direct)
    assert not torch.equal(replayed, outs[0])
    big = eng(torch.cat([xa, xb, xa]), 500.0)                                # larger batch: arena regrows, graphs are dropped
    assert torch.equal(big[:2], outs[0])
    assert torch.equal(eng(x, torch.tensor([37.0, 912.0]), out=out), direct)


def test_face_unet_full_geometry_runs():
    """CelebA-HQ geometry (ch 128, mult (1,1,2,2,4,4), 256x256, attention at 16x16), random-init: forward against torch."""
    _fp32()
    model = FaceUNet(FaceUNetConfig()).cuda()
    eng = hedit_b200.FaceUNetEngine.from_model(model)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(2, 3, 256, 256, generator=g).cuda()
    t = torch.tensor([801.0, 201.0]).cuda()
    with torch.no_grad():
        ref = model(x, t)
    out = eng(x, t)
    r, m = rel_err(out, ref)
    print(f"face unet CelebA-HQ geometry: rel {r:.3e} max {m:.3e}")
    assert r < TOL_UNET


def test_face_h_edit_R_matches_reference_golden():
    _fp32()
    g = load_golden("tiny_face_k2")
    meta = g["meta"]
    u = meta["unet"]
    cfg = FaceUNetConfig(ch=u["ch"], ch_mult=tuple(u["ch_mult"]), num_res_blocks=u["num_res_blocks"], attn_resolutions=tuple(u["attn_resolutions"]),
                         image_size=u["image_size"])
    model = FaceUNet(cfg).cuda()
    idloss, lpipsloss = TinyIDLoss(g["ref_img"]).cuda(), TinyLPIPSLoss(g["x0"].clone()).cuda()
    T, K = meta["T"], meta["K"]
    kw = dict(xT=g["xT"].cuda(), betas=g["betas"], seq=meta["seq"], eta=1.0, zs=g["zs"].cuda(), weight_edit_face=meta["weight_edit_face"],
              optimization_steps=K, after_skip_steps=T, num_inference_steps=T, soft_face_mask=None)
    ed = hedit_b200.face.h_Edit_R(model, lpipsloss, idloss, **kw)
    eng = hedit_b200.face.get_face_engine(model)
    assert eng.last_stats["sample_forwards"] == T + 2 * K * (T - 1)          # the last step (tm1 == 0) runs no implicit loops
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    nr = hedit_b200.face.h_Edit_R(model, None, None, **kw)
    r_nr, _ = rel_err(nr.cpu(), g["no_reward"])
    r_x0, _ = rel_err(nr.cpu(), g["x0"])
    r_far, _ = rel_err(ed.cpu(), g["no_reward"])
    print(f"face h_Edit_R: edited rel {r_ed:.3e} max {m_ed:.3e} | no-reward rel {r_nr:.3e} (returns x0 to {r_x0:.3e}) | distance edited<->no-reward {r_far:.3e}")
    assert r_ed < TOL_LOOP and r_nr < TOL_LOOP and r_x0 < TOL_LOOP
    assert r_far > 5 * TOL_LOOP


def test_face_sde_inversion_matches_reference_golden():
    """inversion_forward_process_sde (sde_inversion.py:54) on the native denoiser: same seed-42 draws as the reference, all T noise
    predictions in one batched call; z_t against the tensors the reference produced for the face golden."""
    _fp32()
    g = load_golden("tiny_face_k2")
    meta = g["meta"]
    u = meta["unet"]
    cfg = FaceUNetConfig(ch=u["ch"], ch_mult=tuple(u["ch_mult"]), num_res_blocks=u["num_res_blocks"], attn_resolutions=tuple(u["attn_resolutions"]),
                         image_size=u["image_size"])
    model = FaceUNet(cfg)          # CPU module: only its weights are used
    T = meta["T"]
    _, zs, xts, _ = hedit_b200.face.inversion_forward_process_sde(model, g["x0"], g["betas"], meta["seq"], etas=1.0, num_inference_steps=T)
    r_x, _ = rel_err(xts[T], g["xT"])
    r_z, m_z = rel_err(zs, g["zs"])
    print(f"face sde inversion: xT rel {r_x:.3e} | zs rel {r_z:.3e} max {m_z:.3e}")
    assert r_x < 1e-6 and r_z < 2e-2
