"""GPU: native CLIP text tower (hedit_text_*) against transformers' CLIPTextModel (the class the reference's `model.text_encoder` is,
text-guided/inversion/inversion_utils.py:13-36) with seeded random-init weights: SD-1.x geometry (12 layers, width 768, 12 heads, 77
tokens, quick_gelu, causal mask)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import hedit_b200  # noqa: E402
from gpu_util import rel_err  # noqa: E402


@pytest.mark.parametrize("layers,width,heads,vocab", [(2, 128, 2, 1000), (12, 768, 12, 49408)])
def test_text_encoder_matches_transformers(layers, width, heads, vocab):
    transformers = pytest.importorskip("transformers")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg = transformers.CLIPTextConfig(vocab_size=vocab, hidden_size=width, intermediate_size=4 * width, num_hidden_layers=layers,
                                      num_attention_heads=heads, max_position_embeddings=77, hidden_act="quick_gelu")
    model = transformers.CLIPTextModel(cfg).eval()
    with torch.no_grad():       # default init is tiny (std 0.02): scale up so that the check exercises every layer
        for n, p in model.named_parameters():
            if p.dim() >= 2 and "embedding" not in n:
                p.mul_(3.0)
    eng = hedit_b200.TextEncoderEngine.from_text_encoder(model)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, vocab, (3, 77), generator=g)
    ids[:, 0] = vocab - 2
    with torch.no_grad():
        ref = model.cuda()(ids.cuda())[0]
    out = eng(ids)[0]
    r, m = rel_err(out, ref)
    print(f"text encoder layers={layers} width={width}: rel {r:.3e} max {m:.3e}")
    assert out.shape == (3, 77, width) and r < 5e-3
