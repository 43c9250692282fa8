"""GPU: native VAE decoder (forward + input-gradient backward, C ABI hedit_vae_*) against the oracle restatement of diffusers'
AutoencoderKL decoder evaluated with torch fp32 + autograd on the same seeded weights.  Tolerances: 16-bit (fp16) conv / linear
operands with fp32 accumulation through ~30 layers; gradients additionally pass the attention softmax backward in 16 bits."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.vae import AutoencoderKLDecoder, AutoencoderKLEncoder, VAEConfig  # noqa: E402

import hedit_b200  # noqa: E402
from gpu_util import rel_err  # noqa: E402

TOL_FWD = 1e-2      # relative L2 of the decoded image
TOL_BWD = 3e-2      # relative L2 of dLoss/dz


def _pair(cfg):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    vae = AutoencoderKLDecoder(cfg).cuda()
    eng = hedit_b200.VaeDecoderEngine.from_vae(vae)
    return vae, eng


@pytest.mark.parametrize("B,hw", [(1, 64), (2, 32)])
def test_vae_decode_and_backward_match_torch(B, hw):
    vae, eng = _pair(VAEConfig.tiny())
    g = torch.Generator(device="cpu").manual_seed(3)
    z = (torch.randn(B, 4, hw, hw, generator=g) * 5).cuda()
    zr = z.clone().requires_grad_(True)
    ref = vae.decode(zr).sample
    out = eng.decode(z).sample
    assert out.shape == ref.shape == (B, 3, 8 * hw, 8 * hw)
    r, m = rel_err(out, ref.detach())
    print(f"vae decode B={B} {hw}x{hw}: rel {r:.3e} max {m:.3e} launches {eng.last_stats['kernel_launches']} GFLOP {eng.last_stats['flops'] / 1e9:.1f}")
    assert r < TOL_FWD
    # a loss with a non-trivial image gradient: weighted sum + quadratic term
    wgt = torch.randn(ref.shape, generator=g).cuda()
    loss = (ref * wgt).sum() + 0.5 * (ref * ref).sum()
    gz_ref = torch.autograd.grad(loss, zr)[0]
    dimg = wgt + ref.detach()
    gz = eng.backward(dimg)
    rb, mb = rel_err(gz, gz_ref)
    print(f"vae backward: rel {rb:.3e} max {mb:.3e} (|grad| max {gz_ref.abs().max().item():.3e})")
    assert rb < TOL_BWD
    # per-image scale invariance of the wrapper and linearity of the backward in dimg
    gz2 = eng.backward(dimg * 1e-6)
    assert rel_err(gz2 * 1e6, gz)[0] < 1e-3


def test_vae_full_size_geometry():
    """SD-1.x decoder geometry (128,256,512,512), one 64x64 latent -> 512x512 image, random-init weights: forward against torch."""
    vae, eng = _pair(VAEConfig())
    g = torch.Generator(device="cpu").manual_seed(5)
    z = (torch.randn(1, 4, 64, 64, generator=g) * 5).cuda()
    with torch.no_grad():
        ref = vae.decode(z).sample
    out = eng.decode(z).sample
    r, m = rel_err(out, ref)
    print(f"vae decode SD geometry: rel {r:.3e} max {m:.3e} GFLOP {eng.last_stats['flops'] / 1e9:.1f}")
    assert r < TOL_FWD
    gz = eng.backward(torch.ones_like(out))
    assert torch.isfinite(gz).all() and gz.abs().max().item() > 0


@pytest.mark.parametrize("cfg,B,px", [(VAEConfig.tiny(), 2, 256), (VAEConfig(), 1, 512)])
def test_vae_encode_matches_torch(cfg, B, px):
    """`vae.encode(x).latent_dist.mode()` (main_p2p.py:154-159): wide-image convs and the (0,1,0,1)-padded stride-2 downsamplers."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    vae = AutoencoderKLEncoder(cfg).cuda()
    eng = hedit_b200.VaeEncoderEngine.from_vae(vae)
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.tanh(torch.randn(B, 3, px, px, generator=g)).cuda()
    with torch.no_grad():
        ref = vae.encode(x).latent_dist
    out = eng.encode(x).latent_dist
    r, m = rel_err(out.mode(), ref.mode())
    r2, _ = rel_err(out.logvar, ref.logvar)
    print(f"vae encode {px}px boc={cfg.block_out_channels}: mean rel {r:.3e} max {m:.3e} | logvar rel {r2:.3e}")
    assert out.mode().shape == (B, 4, px // 8, px // 8)
    assert r < TOL_FWD and r2 < TOL_FWD
