"""GPU: the native face-swapping reward networks (csrc/reward.cu) -- ArcFace IR-SE50 identity loss and LPIPS-VGG16,
loss value and IMAGE GRADIENT through the C-ABI against torch fp32 autograd on the same weights (hedit_b200/reward_nets.py: the torch
restatement of both networks, whose IR-SE50 is pinned bit-exactly to the reference's `Backbone` on CPU in test_host_logic.py), and the
native h_Edit_R loop with native rewards against the same loop with the torch-autograd plug-in route.

Tolerances (16-bit conv operands, fp32 accumulate; ~50 / 13 convs forward and as many backward): embedding 5e-3, loss 5e-3 relative.
The image gradient passes ~100 layers and, above all, the PReLU / ReLU / max-pool DECISIONS of the forward pass: a pre-activation that
16-bit rounding moves across zero flips that element's derivative.  Its bound is therefore calibrated like the loop-level bounds
(DESIGN 6): the same torch network with both operands of every conv / linear rounded to 16 bits (tests/fp16_emulation.py) gives the
deviation the reference's own arithmetic shows at this operand precision; the native gradient must stay within 3x that + 5e-3."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hedit_b200  # noqa: E402
from hedit_b200 import reward, reward_nets  # noqa: E402
from gpu_util import rel_err  # noqa: E402
from fp16_emulation import operand_rounding  # noqa: E402

TOL_FEAT, TOL_LOSS, GRAD_FLOOR = 5e-3, 5e-3, 5e-3


def _fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _images(n, R, seed, noise=False):
    """smooth-ish images in [-1, 1] (a reward network sees Tweedie predictions of faces), or white noise of amplitude 1.5 (what a
    random-init denoiser's Tweedie prediction looks like in the loop test)"""
    g = torch.Generator().manual_seed(seed)
    if noise:
        return (1.5 * torch.randn(n, 3, R, R, generator=g)).cuda()
    x = F.interpolate(torch.randn(n, 3, R // 8, R // 8, generator=g), size=(R, R), mode="bicubic", align_corners=False)
    return (0.6 * x + 0.15 * torch.randn(n, 3, R, R, generator=g)).clamp(-1, 1).cuda()


def _autograd(loss_per_image, x, loss_scale=1.0):
    with torch.enable_grad():
        xx = x.detach().clone().requires_grad_(True)
        l = loss_per_image(xx)
        return l.detach(), torch.autograd.grad(l.sum() * loss_scale, xx)[0] / loss_scale


def _emulated_grad(loss_per_image, x, g_ref):
    """Gradient of the same torch network with 16-bit-rounded conv / linear operands.  torch back-propagates through the rounding casts in
    fp16, so the loss is scaled to bring the gradient to unit rms first (the CUDA path normalises its gradient operands per sample too);
    without it the LPIPS gradients (1e-6) simply underflow."""
    with operand_rounding(torch.float16):
        return _autograd(loss_per_image, x, loss_scale=1.0 / float(g_ref.pow(2).mean().sqrt()))[1]


@pytest.mark.parametrize("B,noise", [(1, False), (3, False), (2, True)])
def test_arcface_loss_and_gradient_match_torch_autograd(B, noise):
    _fp32()
    ref_img, x = _images(1, 256, 1), _images(B, 256, 2, noise)
    idl = reward_nets.SyntheticIDLoss(ref_img, seed=0).cuda()
    eng = reward.ArcFaceEngine.from_facenet(idl.facenet)
    eng.set_reference(ref_img)
    with torch.no_grad():
        f_ref = reward_nets.id_features(idl.facenet, x)
    f = eng.features(x)
    r_f, _ = rel_err(f, f_ref)
    with torch.no_grad():
        rf = reward_nets.id_features(idl.facenet, ref_img)
    l_ref, g_ref = _autograd(lambda t: 1 - F.cosine_similarity(rf, reward_nets.id_features(idl.facenet, t), dim=-1), x)
    g_emu = _emulated_grad(lambda t: 1 - F.cosine_similarity(rf, reward_nets.id_features(idl.facenet, t), dim=-1), x, g_ref)
    loss, grad = eng.loss_grad(x)
    r_l = float(((loss - l_ref).abs() / l_ref.abs().clamp_min(1e-3)).max())
    r_g = max(rel_err(grad[b], g_ref[b])[0] for b in range(B))
    r_e = max(rel_err(g_emu[b], g_ref[b])[0] for b in range(B))
    cos = float(F.cosine_similarity(grad.flatten(1), g_ref.flatten(1)).min())
    print(f"arcface B={B}: embedding rel {r_f:.3e} | loss {loss.tolist()} vs {l_ref.tolist()} rel {r_l:.3e} | grad rel {r_g:.3e} (16-bit-operand torch: {r_e:.3e}) "
          f"cos {cos:.5f} (|grad| max {float(g_ref.abs().max()):.3e}) launches {eng.last_stats['kernel_launches']} GFLOP {eng.last_stats['flops'] / 1e9:.1f}")
    assert r_f < TOL_FEAT and r_l < TOL_LOSS and r_g < 3 * r_e + GRAD_FLOOR and cos > 0.999
    # the gradient is zero outside the crop the reference feeds the network (arcface_model.py:44)
    outside = grad.clone(); outside[:, :, 35:223, 32:220] = 0
    assert float(outside.abs().max()) == 0.0
    # same buffers again: 2nd call is captured, 3rd replayed from the CUDA graph -- identical bits
    g2, l2 = torch.empty_like(grad), torch.empty_like(loss)
    outs = []
    for _ in range(3):
        eng.loss_grad(x, grad_out=g2, loss_out=l2)
        outs.append(g2.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]) and torch.equal(outs[0], grad)


@pytest.mark.parametrize("R,B,nsrc,noise", [(128, 2, 1, False), (256, 2, 2, False), (256, 1, 1, True)])
def test_lpips_loss_and_gradient_match_torch_autograd(R, B, nsrc, noise):
    _fp32()
    src, x = _images(nsrc, R, 3), _images(B, R, 4, noise)
    net = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 1).cuda()
    eng = reward.LpipsEngine.from_module(net)
    eng.set_source(src)
    with torch.no_grad():
        taps = [f / (f.pow(2).sum(1, keepdim=True).sqrt() + 1e-10) for f in net.taps(src)]
    l_ref, g_ref = _autograd(lambda t: net(t, taps), x)
    g_emu = _emulated_grad(lambda t: net(t, taps), x, g_ref)
    loss, grad = eng.loss_grad(x)
    r_l = float(((loss - l_ref).abs() / l_ref.abs().clamp_min(1e-6)).max())
    r_g = max(rel_err(grad[b], g_ref[b])[0] for b in range(B))
    r_e = max(rel_err(g_emu[b], g_ref[b])[0] for b in range(B))
    print(f"lpips R={R} B={B} nsrc={nsrc}: loss {loss.tolist()} vs {l_ref.tolist()} rel {r_l:.3e} | grad rel {r_g:.3e} (16-bit-operand torch: {r_e:.3e}) "
          f"(|grad| max {float(g_ref.abs().max()):.3e}) launches {eng.last_stats['kernel_launches']} GFLOP {eng.last_stats['flops'] / 1e9:.1f}")
    assert r_l < TOL_LOSS and r_g < 3 * r_e + GRAD_FLOOR
    # an image equal to its source: zero loss, zero gradient (the path's intrinsic known answer)
    if nsrc == B:
        l0, g0 = eng.loss_grad(src)
        assert float(l0.abs().max()) < 1e-6 and float(g0.abs().max()) < 1e-6 * float(g_ref.abs().max()) + 1e-12


def test_lpips_package_state_dict_names_are_recognised():
    """The lpips package's key layout (`net.slice3.12.weight`, `lin2.model.1.weight`, `scaling_layer.shift`) loads like ours."""
    net = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 1)
    sd, slices = {}, {0: 1, 2: 1, 5: 2, 7: 2, 10: 3, 12: 3, 14: 3, 17: 4, 19: 4, 21: 4, 24: 5, 26: 5, 28: 5}
    for k, v in net.state_dict().items():
        if k.startswith("features."):
            idx = int(k.split(".")[1])
            sd[f"net.slice{slices[idx]}.{idx}.{k.split('.')[2]}"] = v
        elif k.startswith("lins."):
            sd[f"lin{k.split('.')[1]}.model.1.weight"] = v
            sd[f"lins.{k.split('.')[1]}.model.1.weight"] = v
        else:
            sd[f"scaling_layer.{k}"] = v
    a, b = reward.lpips_native_tensors(sd), reward.lpips_native_tensors(net.state_dict())
    assert set(a) == set(b) and all(torch.equal(a[k].reshape(-1), b[k].reshape(-1)) for k in a)
    eng = reward.LpipsEngine()
    eng.load_state_dict(sd)


def test_face_h_edit_R_native_rewards_match_autograd_plugins():
    """h_Edit_R through the reference signature with IDLoss / LPIPS_Loss-shaped reward objects: the native reward networks (default) and
    the torch-autograd plug-in route (HEDIT_NATIVE_REWARD=0) steer the same native loop to the same image."""
    _fp32()
    from oracle.face_unet import FaceUNet, FaceUNetConfig
    import numpy as np
    cfg = FaceUNetConfig(ch=64, ch_mult=(1, 1, 2, 2), image_size=256, attn_resolutions=(32,))
    model = FaceUNet(cfg).cuda()
    T = 4
    betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64).float().cuda()
    seq = (np.arange(0, 1000, 1000 // T) + 1)[::-1]
    g = torch.Generator().manual_seed(11)
    xT = torch.randn(1, 3, 256, 256, generator=g).cuda()
    zs = torch.randn(T, 3, 256, 256, generator=g).cuda()
    idl = reward_nets.SyntheticIDLoss(_images(1, 256, 7), seed=0).cuda()
    lpl = reward_nets.SyntheticLPIPSLoss(_images(1, 256, 8), seed=1).cuda()
    outs = {}
    for native in ("1", "0"):
        os.environ["HEDIT_NATIVE_REWARD"] = native
        outs[native] = hedit_b200.face.h_Edit_R(model, lpl, idl, xT, betas, seq, eta=1.0, zs=zs, weight_edit_face=2000.0, optimization_steps=2,
                                                after_skip_steps=T, num_inference_steps=T)
        used = hedit_b200.face.get_face_engine(model).last_stats["native_rewards"]
        assert used == ((True, True) if native == "1" else (False, False)), used
    os.environ.pop("HEDIT_NATIVE_REWARD", None)
    none = hedit_b200.face.h_Edit_R(model, None, None, xT, betas, seq, eta=1.0, zs=zs, weight_edit_face=2000.0, optimization_steps=2,
                                    after_skip_steps=T, num_inference_steps=T)
    r, m = rel_err(outs["1"], outs["0"])
    moved = rel_err(outs["0"], none)[0]
    print(f"h_Edit_R native vs autograd rewards: rel {r:.3e} max {m:.3e} | the rewards move the image by {moved:.3e}")
    assert moved > 1e-3 and r < 0.1 * moved + 2e-3


@pytest.mark.parametrize("name", ["face256_irse50_lpips_k2", "celebahq_config5_T5_irse50_lpips_k3"])
def test_face_h_edit_R_against_reference_sampler_with_reference_reward_classes(name):
    """Golden `face256_irse50_lpips_k2` = the UNMODIFIED reference `h_Edit_R` driving the reference's own `IDLoss` (around its `Backbone(112,
    50, 'ir_se')`) and `LPIPS_Loss` classes on CPU (tests/make_golden.py --config face_full; seeded weights); `celebahq_config5_T5_irse50_lpips_k3` is the same at BASELINE.json configs[4] geometry (CelebA-HQ DDPM UNet, K = 3).  Here: the native loop, the
    native DDPM UNet and the NATIVE reward networks with their image gradients, fed the same weights through reward objects of the same
    layout.  The rewards move the result by 50 %; the bound is the loop tolerance plus the calibrated gradient deviation (3-7 %) of that move."""
    import numpy as np
    from oracle.face_unet import FaceUNet, FaceUNetConfig
    from oracle_run import load_golden
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", name + ".pt")):
        pytest.skip("golden missing")
    _fp32()
    g = load_golden(name)
    meta, u = g["meta"], g["meta"]["unet"]
    cfg = FaceUNetConfig(ch=u["ch"], ch_mult=tuple(u["ch_mult"]), image_size=u["image_size"], attn_resolutions=tuple(u["attn_resolutions"]))
    model = FaceUNet(cfg).cuda()
    T, K = meta["T"], meta["K"]
    idl = reward_nets.SyntheticIDLoss(g["ref_img"], seed=meta["irse_seed"]).cuda()
    lpl = reward_nets.SyntheticLPIPSLoss(g["x0"], seed=meta["vgg_seed"]).cuda()
    with torch.no_grad():
        for p in lpl.lpips_loss.lins:
            p.mul_(meta["lin_gain"])
    kw = dict(eta=1.0, zs=g["zs"].cuda(), weight_edit_face=meta["weight_edit_face"], optimization_steps=K, after_skip_steps=T, num_inference_steps=T)
    betas, seq = g["betas"].cuda(), np.asarray(meta["seq"])
    ed = hedit_b200.face.h_Edit_R(model, lpl, idl, g["xT"].cuda(), betas, seq, **kw)
    assert hedit_b200.face.get_face_engine(model).last_stats["native_rewards"] == (True, True)
    ed_id = hedit_b200.face.h_Edit_R(model, None, idl, g["xT"].cuda(), betas, seq, **kw)
    none = hedit_b200.face.h_Edit_R(model, None, None, g["xT"].cuda(), betas, seq, **kw)
    # the same loop with the reward modules differentiated by torch autograd (fp32): separates the reward networks' 16-bit operands from
    # the denoiser's in the deviation from the golden
    os.environ["HEDIT_NATIVE_REWARD"] = "0"
    try:
        ed_ag = hedit_b200.face.h_Edit_R(model, lpl, idl, g["xT"].cuda(), betas, seq, **kw)
    finally:
        os.environ.pop("HEDIT_NATIVE_REWARD", None)
    r_ag, _ = rel_err(ed_ag.cpu(), g["edited"])
    r_ed, m_ed = rel_err(ed.cpu(), g["edited"])
    r_id, _ = rel_err(ed_id.cpu(), g["edited_id_only"])
    r_no, _ = rel_err(none.cpu(), g["no_reward"])
    moved = rel_err(g["edited"], g["no_reward"])[0]
    lp_effect = rel_err(g["edited"], g["edited_id_only"])[0]
    print(f"{name}: edited rel {r_ed:.3e} max {m_ed:.3e} (fp32 autograd rewards on the same native loop: {r_ag:.3e}) | identity only rel {r_id:.3e} | no reward rel {r_no:.3e} | "
          f"rewards move the result by {moved:.3e}, LPIPS alone by {lp_effect:.3e}")
    assert r_no < 4e-2
    assert r_ed < 4e-2 + 0.07 * moved and r_id < 4e-2 + 0.07 * moved
    assert r_ed < 1.25 * r_ag + 5e-3          # the native reward networks add next to nothing to what the denoiser's 16-bit operands already cost
    assert lp_effect > 2 * r_ed or lp_effect > 1e-2          # the LPIPS term is visible in the golden
