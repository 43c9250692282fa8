#!/usr/bin/env python
"""Per-GPU throughput of the other BASELINE.json configurations (not the headline: bench.py keeps that), synthetic inputs, random-init
weights of the reference geometries.
  --config 3 : explicit h-Edit + MasaCtrl, SD-1.5 512^2, 50 steps, 8 images per GPU (configs[2])
  --config 5 : face swapping h_Edit_R, CelebA-HQ DDPM UNet 256^2, 100 steps, K = 3, 8 images per GPU (configs[4]); the ArcFace / LPIPS
               reward networks are replaced by small seeded conv nets (their weights and the lpips package do not exist offline), so the
               number is an upper bound dominated by the 700 denoiser calls per image"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(0)
B = a.batch


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


if a.config == 3:
    T = 50
    cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
               cross_attention_dim=768, norm_groups=32, ctx_len=77)
    eng = hedit_b200.UNetEngine(cfg, max_samples=5 * B, max_contexts=1 + 2 * B)
    eng.load_random_weights(0)
    sched = hedit_b200.DDIMTables(T, steps_offset=0)
    ts, coef = hedit_b200.step_tables(sched, T, 1.0, True)       # h-Edit-D: eta 0 inversion, is_ddim_inversion = True
    xT = torch.randn(B, 4, 64, 64, generator=g, device=dev); zs = torch.randn(B, T, 4, 64, 64, generator=g, device=dev) * 0.01
    ctx = torch.randn(1 + 2 * B, 77, 768, generator=g, device=dev)
    (ed, rc), ms = timed(lambda: eng.edit(xT, zs, ctx, ts, coef, [1.0, 5.0, 7.5], None, 0.0, 1, True, 1, masactrl=hedit_b200.MutualSelfAttentionControl(4, 10, total_steps=T).launch_plan(T), mos_pull=False))
    fwd = eng.last_stats["sample_forwards"]
    print(json.dumps({"workload": f"explicit h-Edit-D + MasaCtrl (step 4, layer 10), SD-1.5 512^2, {T} steps, batch {B}, 1 GPU", "images_per_s": B / (ms / 1e3),
                      "unet_sample_forwards_per_image": fwd / B, "achieved_tflops": fwd * 0.8033 / (ms / 1e3), "finite": bool(torch.isfinite(ed).all())}))
else:
    T, K = 100, 3
    eng = hedit_b200.FaceUNetEngine(dict(ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolution=16, image_size=256, in_channels=3, out_ch=3))
    eng.load_random_weights(0)
    betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64).float()
    seq = (np.arange(0, 1000, 1000 // T) + 1)[::-1]
    coef = hedit_b200.face.face_step_tables(betas, seq, T, T, 1.0)
    xT = torch.randn(B, 3, 256, 256, generator=g, device=dev); zs = torch.randn(B, T, 3, 256, 256, generator=g, device=dev)
    net_id = torch.nn.Sequential(torch.nn.Conv2d(3, 32, 3, 2, 1), torch.nn.PReLU(32), torch.nn.Conv2d(32, 64, 3, 2, 1), torch.nn.AdaptiveAvgPool2d(4), torch.nn.Flatten(), torch.nn.Linear(1024, 128)).to(dev)
    net_lp = torch.nn.Sequential(torch.nn.Conv2d(3, 32, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(32, 64, 3, 1, 1)).to(dev)
    ref_feat = torch.randn(1, 128, device=dev); src = torch.randn(1, 3, 256, 256, device=dev)

    def grad_of(loss):
        def fn(x0):
            with torch.enable_grad():
                x = x0.detach().clone().requires_grad_(True)
                return torch.autograd.grad(loss(x), x)[0]
        return fn
    id_grad = grad_of(lambda x: (1 - torch.nn.functional.cosine_similarity(net_id(x), ref_feat)).sum())
    lp_grad = grad_of(lambda x: (net_lp(x) - net_lp(src)).pow(2).mean(dim=(1, 2, 3)).sum())
    ed, ms = timed(lambda: eng.edit(xT, zs, coef, 50.0, K, id_grad=id_grad, lpips_grad=lp_grad))
    fwd = eng.last_stats["sample_forwards"]
    print(json.dumps({"workload": f"face swapping h_Edit_R, CelebA-HQ DDPM UNet 256^2, {T} steps, K={K}, batch {B}, 1 GPU, stand-in reward nets",
                      "images_per_s": B / (ms / 1e3), "unet_calls_per_image": fwd / B, "achieved_tflops": fwd * 0.497 / (ms / 1e3),
                      "finite": bool(torch.isfinite(ed).all())}))
