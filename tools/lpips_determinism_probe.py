import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hedit_b200 import reward, reward_nets
h = lambda t: hashlib.md5(t.detach().cpu().numpy().tobytes()).hexdigest()[:10]
net = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 4).cuda()
g = torch.Generator().manual_seed(0)
for R in (128, 256):
    for B in (1, 2):
        x = (torch.randn(B, 3, R, R, generator=g) * 0.4).clamp(-1, 1).cuda(); src = (torch.randn(1, 3, R, R, generator=g) * 0.4).clamp(-1, 1).cuda()
        e = reward.LpipsEngine.from_module(net); e.set_source(src)
        l, gr = e.loss_grad(x)
        print(f"R={R} B={B} loss {h(l)} {l.tolist()} grad {h(gr)} nan {bool(torch.isnan(gr).any())} |grad| {float(gr.abs().max()):.4e}")
