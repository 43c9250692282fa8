#!/usr/bin/env python
"""BASELINE.json configs[3] on one GPU: text-guided-n-style implicit h-Edit + P2P + CLIP-style reward, SD-1.5 geometry (64x64 latent,
512x512 decode), ViT-B/16 CLIP geometry, K Langevin/implicit iterations per timestep, random-init weights, synthetic inputs.
Prints one JSON line (not the headline bench: bench.py keeps BASELINE's metric).  python tools/style_bench.py --batch 8 --K 3"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--timesteps", type=int, default=50)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, T, K = a.batch, a.timesteps, a.K
    cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
               cross_attention_dim=768, norm_groups=32, ctx_len=77)
    eng = hedit_b200.UNetEngine(cfg, max_samples=5 * B, max_contexts=1 + 2 * B)
    eng.load_random_weights(0)
    vae = hedit_b200.VaeDecoderEngine(dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_groups=32))
    vae.load_random_weights(1)
    clip = hedit_b200.ClipGramEngine(224, 16, 768, 12, 3)
    g = torch.Generator(device="cuda").manual_seed(2)
    sd = {}
    W = 768
    def rnd(*s, scale=1.0):
        return (torch.rand(*s, generator=g, device=dev) * 2 - 1) * scale
    sd["conv1.weight"] = rnd(W, 3, 16, 16, scale=(3 / 768) ** 0.5)
    sd["class_embedding"] = rnd(W, scale=0.05); sd["positional_embedding"] = rnd(197, W, scale=0.05)
    sd["ln_pre.weight"] = 1 + rnd(W, scale=0.1); sd["ln_pre.bias"] = rnd(W, scale=0.1)
    for i in range(3):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = 1 + rnd(W, scale=0.1); sd[p + "ln_1.bias"] = rnd(W, scale=0.1)
        sd[p + "ln_2.weight"] = 1 + rnd(W, scale=0.1); sd[p + "ln_2.bias"] = rnd(W, scale=0.1)
        sd[p + "attn.in_proj_weight"] = rnd(3 * W, W, scale=(3 / W) ** 0.5); sd[p + "attn.in_proj_bias"] = rnd(3 * W, scale=0.1)
        sd[p + "attn.out_proj.weight"] = rnd(W, W, scale=(3 / W) ** 0.5); sd[p + "attn.out_proj.bias"] = rnd(W, scale=0.1)
        sd[p + "mlp.c_fc.weight"] = rnd(4 * W, W, scale=(3 / W) ** 0.5); sd[p + "mlp.c_fc.bias"] = rnd(4 * W, scale=0.1)
        sd[p + "mlp.c_proj.weight"] = rnd(W, 4 * W, scale=(3 / (4 * W)) ** 0.5); sd[p + "mlp.c_proj.bias"] = rnd(W, scale=0.1)
    clip.load_state_dict(sd)
    clip.set_reference(torch.randn(1, 3, 224, 224, generator=g, device=dev))
    tok = hedit_b200.WordTokenizer()
    sched = hedit_b200.DDIMTables(T, steps_offset=1)
    ts, coef = hedit_b200.step_tables(sched, T, 1.0, False)
    prompts = ["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"]
    ctrls = [hedit_b200.make_controller(prompts, False, 0.4, 0.35, blend_word=None, equilizer_params=None, num_steps=T, tokenizer=tok) for _ in range(B)]
    plan = hedit_b200.compile_edit_plan(ctrls, T)
    xT = torch.randn(B, 4, 64, 64, generator=g, device=dev)
    zs = torch.randn(B, T, 4, 64, 64, generator=g, device=dev)
    ctx = torch.randn(1 + 2 * B, 77, 768, generator=g, device=dev)
    fn = hedit_b200.style.clip_gram_guidance_fused(vae, clip, vae_batch=2)
    t_reward = [0.0]

    def timed_fn(x0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(x0); e1.record()
        ev.append((e0, e1))
        return out

    ev = []
    guidance = (timed_fn, 0.5, hedit_b200.x0_tables(sched, T))
    run = lambda steps: eng.edit(xT, zs[:, :steps], ctx, ts[:steps] + [0] if steps < T else ts, coef[:steps], [1.0, 5.0, 7.5], plan if steps == T else None,
                                 0.0, K, False, 1, mos_pull=False, guidance=(timed_fn, 0.5, hedit_b200.x0_tables(sched, T)[:steps]))
    run(1)      # warm-up (arenas, plans)
    torch.cuda.synchronize()
    ev.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        ed, rc = run(T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    ms_reward = sum(x.elapsed_time(y) for x, y in ev) / a.reps
    fwd = eng.last_stats["sample_forwards"]
    tf_unet = fwd * 0.8033
    tf_reward = B * T * K * (2.514 + 2.55 + 0.03)
    print(json.dumps({"workload": f"text+style implicit h-Edit + P2P + CLIP-Gram reward, SD-1.5 512^2, {T} steps, K={K}, batch {B}, 1 GPU", "images_per_s": B / (ms / 1e3),
                      "ms_per_batch": ms, "reward_branch_ms": ms_reward, "reward_share": ms_reward / ms, "unet_sample_forwards_per_image": fwd / B,
                      "achieved_tflops": (tf_unet + tf_reward) / (ms / 1e3), "reward_tflops": tf_reward / (ms_reward / 1e3),
                      "finite": bool(torch.isfinite(ed).all())}))


if __name__ == "__main__":
    main()
