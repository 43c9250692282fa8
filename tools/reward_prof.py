"""One ArcFace and one LPIPS loss+gradient call at batch 8 for an ncu launch list (run with HEDIT_NET_GRAPH=0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hedit_b200 import reward, reward_nets
B = 8
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, 3, 256, 256, generator=g) * 0.4).clamp(-1, 1).cuda()
ref = (torch.randn(1, 3, 256, 256, generator=g) * 0.4).clamp(-1, 1).cuda()
which = sys.argv[1] if len(sys.argv) > 1 else "arcface"
if which == "arcface":
    eng = reward.ArcFaceEngine.from_facenet(reward_nets._seed_init(reward_nets.IRSE50(), 0).cuda()); eng.set_reference(ref)
else:
    eng = reward.LpipsEngine.from_module(reward_nets._seed_init(reward_nets.LPIPSVGG16(), 1).cuda()); eng.set_source(ref)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("measured")
eng.loss_grad(x); torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
