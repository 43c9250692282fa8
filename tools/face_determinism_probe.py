"""Run the native face edit (native rewards) twice in this process and print checksums (run the script twice to compare processes)."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import hedit_b200
from hedit_b200 import reward, reward_nets
from oracle.face_unet import FaceUNet, FaceUNetConfig
from oracle_run import load_golden
g = load_golden("face256_irse50_lpips_k2"); meta, u = g["meta"], g["meta"]["unet"]
cfg = FaceUNetConfig(ch=u["ch"], ch_mult=tuple(u["ch_mult"]), image_size=u["image_size"], attn_resolutions=tuple(u["attn_resolutions"]))
model = FaceUNet(cfg).cuda()
idl = reward_nets.SyntheticIDLoss(g["ref_img"], seed=3).cuda(); lpl = reward_nets.SyntheticLPIPSLoss(g["x0"], seed=4).cuda()
with torch.no_grad():
    for p in lpl.lpips_loss.lins: p.mul_(100.0)
kw = dict(eta=1.0, zs=g["zs"].cuda(), weight_edit_face=1500.0, optimization_steps=2, after_skip_steps=4, num_inference_steps=4)
h = lambda t: hashlib.md5(t.detach().cpu().numpy().tobytes()).hexdigest()[:12]
betas, seq = g["betas"].cuda(), np.asarray(meta["seq"])
for which, (lp, idd) in {"none": (None, None), "id": (None, idl), "lpips": (lpl, None), "both": (lpl, idl)}.items():
    outs = [h(hedit_b200.face.h_Edit_R(model, lp, idd, g["xT"].cuda(), betas, seq, **kw)) for _ in range(3)]
    print(which, outs)
x = g["x0"].cuda()
arc = reward.ArcFaceEngine.from_facenet(idl.facenet); arc.set_reference(g["ref_img"].cuda())
print("arcface grad", [h(arc.loss_grad(x)[1]) for _ in range(3)], "feat", h(arc.features(x)))
lpe = reward.LpipsEngine.from_module(lpl.lpips_loss); lpe.set_source(g["ref_img"].cuda())
print("lpips grad", [h(lpe.loss_grad(x)[1]) for _ in range(3)])
