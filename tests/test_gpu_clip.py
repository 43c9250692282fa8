"""GPU: native CLIP-Gram style reward (C ABI hedit_clip_*) -- loss and image gradient against the oracle restatement of the reference's
CLIPEncoder.get_gram_matrix_residual (pinned to the reference classes on CPU by tests/test_oracle_pin.py) evaluated with torch fp32 +
autograd on the same seeded weights.  Tolerances: fp16 operands of the patch embedding and of the 3 ViT blocks' linear layers."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.clip_visual import tiny_style_encoder  # noqa: E402

import hedit_b200  # noqa: E402
from gpu_util import rel_err  # noqa: E402


@pytest.mark.parametrize("width,B,hw", [(64, 1, 512), (768, 2, 512), (128, 1, 96)])
def test_clip_gram_loss_and_gradient_match_torch(width, B, hw):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    enc = tiny_style_encoder(seed=11, width=width).cuda()
    eng = hedit_b200.ClipGramEngine.from_image_encoder(enc)
    g = torch.Generator(device="cpu").manual_seed(7)
    img = torch.tanh(torch.randn(B, 3, hw, hw, generator=g)).cuda()
    x = img.clone().requires_grad_(True)
    losses = torch.stack([torch.linalg.norm(enc.get_gram_matrix_residual(x[b:b + 1])) for b in range(B)])
    gref = torch.autograd.grad(losses.sum(), x)[0]
    out = eng.loss(img)
    dimg = eng.backward()
    rl, _ = rel_err(out, losses.detach())
    rg, mg = rel_err(dimg, gref)
    print(f"clip gram width={width} B={B} {hw}px: loss {losses.tolist()} rel {rl:.3e} | grad rel {rg:.3e} max {mg:.3e} (|grad| max {gref.abs().max().item():.3e})")
    assert rl < 5e-3 and rg < 3e-2
