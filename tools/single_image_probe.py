"""Single-image path probe: time h_Edit_p2p_implicit (one image per call) with split-K on (default for one-image calls) / off."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = types.SimpleNamespace(batch=1, timesteps=50, schedule=1, steps=1, warmup=1, config=2)
wl = bench.Headline(args, 0, torch.device("cuda", 0))
for label, env in (("default", None), ("HEDIT_GEMM_SPLITK=0 equivalent (flag off)", False)):
    if env is False:
        import hedit_b200.samplers as S
        orig = S.UNetEngine.set_splitk
        S.UNetEngine.set_splitk = lambda self, on: orig(self, False)
    wl.single_image(2)
    t = wl.single_image(4)
    eng = wl.h.get_engine(wl.pipe)
    print(label, "images/s", round(t, 4), "launches of the last edit", eng.last_stats)
