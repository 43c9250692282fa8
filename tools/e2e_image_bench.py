#!/usr/bin/env python
"""Secondary measurement of SURVEY 8d config 2: full image -> image on one GPU, every stage native -- VAE encode (512x512 -> 64x64
latent), edit-friendly DDPM inversion (2T UNet sample-forwards per image, batched), implicit h-Edit + P2P edit (T steps), VAE decode.
SD-1.5 / SD-VAE geometry, random-init weights, synthetic text contexts.  Prints one JSON line (bench.py keeps the headline metric)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedit_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--timesteps", type=int, default=50)
    a = ap.parse_args()
    B, T = a.batch, a.timesteps
    dev = torch.device("cuda", 0)
    cfg = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, heads=8,
               cross_attention_dim=768, norm_groups=32, ctx_len=77)
    vcfg = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_groups=32)
    eng = hedit_b200.UNetEngine(cfg, max_samples=5 * B, max_contexts=1 + 2 * B)
    eng.load_random_weights(0)
    dec = hedit_b200.VaeDecoderEngine(vcfg)
    dec.load_random_weights(1)
    enc = hedit_b200.VaeEncoderEngine(vcfg)
    g = torch.Generator(device="cuda").manual_seed(3)
    # encoder weights: seeded through the oracle-free path (random tensors by expected shape are not enumerable for the encoder,
    # so reuse the decoder-style initialiser on a throw-away torch module layout)
    import ctypes as C
    from hedit_b200 import _lib
    boc = vcfg["block_out_channels"]
    def load(name, shape, scale=None):
        if len(shape) >= 2:
            fan = 1
            for d in shape[1:]:
                fan *= d
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * (3.0 / fan) ** 0.5
        elif name.endswith("weight"):
            t = 0.8 + 0.4 * torch.rand(shape, generator=g, device=dev)
        else:
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * 0.1
        _lib.check(enc.lib.hedit_vae_enc_load_tensor(enc.handle, name.encode(), t.data_ptr(), (C.c_int64 * len(shape))(*shape), len(shape)), name)
    def res(name, cin, cout):
        load(name + ".norm1.weight", (cin,)); load(name + ".norm1.bias", (cin,)); load(name + ".conv1.weight", (cout, cin, 3, 3)); load(name + ".conv1.bias", (cout,))
        load(name + ".norm2.weight", (cout,)); load(name + ".norm2.bias", (cout,)); load(name + ".conv2.weight", (cout, cout, 3, 3)); load(name + ".conv2.bias", (cout,))
        if cin != cout:
            load(name + ".conv_shortcut.weight", (cout, cin, 1, 1)); load(name + ".conv_shortcut.bias", (cout,))
    load("encoder.conv_in.weight", (boc[0], 3, 3, 3)); load("encoder.conv_in.bias", (boc[0],))
    prev = boc[0]
    for i, c in enumerate(boc):
        for l in range(2):
            res(f"encoder.down_blocks.{i}.resnets.{l}", prev if l == 0 else c, c)
        if i < 3:
            load(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (c, c, 3, 3)); load(f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (c,))
        prev = c
    res("encoder.mid_block.resnets.0", prev, prev); res("encoder.mid_block.resnets.1", prev, prev)
    P = "encoder.mid_block.attentions.0"
    load(P + ".group_norm.weight", (prev,)); load(P + ".group_norm.bias", (prev,))
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        load(f"{P}.{nm}.weight", (prev, prev)); load(f"{P}.{nm}.bias", (prev,))
    load("encoder.conv_norm_out.weight", (prev,)); load("encoder.conv_norm_out.bias", (prev,))
    load("encoder.conv_out.weight", (8, prev, 3, 3)); load("encoder.conv_out.bias", (8,))
    load("quant_conv.weight", (8, 8, 1, 1)); load("quant_conv.bias", (8,))
    _lib.check(enc.lib.hedit_vae_enc_finalize(enc.handle), "finalize")

    tok = hedit_b200.WordTokenizer()
    sched = hedit_b200.DDIMTables(T, steps_offset=1)
    ts, coef = hedit_b200.step_tables(sched, T, 1.0, False)
    prompts = ["a green lizard is sitting on a branch", "a brown lizard is sitting on a branch"]
    ctrls = [hedit_b200.make_controller(prompts, False, 0.4, 0.35, blend_word=(("lizard",), ("lizard",)), equilizer_params={"words": ("lizard",), "values": (2.0,)},
                                        num_steps=T, tokenizer=tok) for _ in range(B)]
    plan = hedit_b200.compile_edit_plan(ctrls, T)
    ctx = torch.randn(1 + 2 * B, 77, 768, generator=g, device=dev)
    imgs = torch.tanh(torch.randn(B, 3, 512, 512, generator=g, device=dev))
    ab = sched.alphas_cumprod.to(dev)
    tl = [int(t) for t in sched.timesteps]
    ratio = 1000 // T

    def pipeline():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        w0 = enc.encode(imgs).latent_dist.mode() * 0.18215                                   # main_p2p.py:154-159
        ev[1].record()
        # edit-friendly DDPM inversion (ddpm_inversion.py:5-167), all images and timesteps batched: x_t ~ q(x_t | x_0), then z_t
        noise = torch.randn(B, T + 1, 4, 64, 64, generator=g, device=dev)
        at = torch.stack([ab[t] for t in reversed(tl)])                                       # idx = 1..T  <->  timesteps ascending
        xts = torch.empty(B, T + 1, 4, 64, 64, device=dev)
        xts[:, 0] = w0
        xts[:, 1:] = w0[:, None] * at.sqrt()[None, :, None, None, None] + noise[:, 1:] * (1 - at).sqrt()[None, :, None, None, None]
        x_in = torch.stack([xts[:, T - k] for k in range(T)], dim=1).reshape(B * T, 4, 64, 64)   # step k starts from xts[T-k]
        t_all = tl * B
        cidx_u = [0] * (B * T)
        cidx_c = [1 + 2 * b for b in range(B) for _ in range(T)]
        eps = eng.forward(torch.cat([x_in, x_in]), t_all + t_all, ctx, ctx_index=cidx_u + cidx_c)
        e_u, e_c = eps[:B * T], eps[B * T:]
        npred = (e_u + 1.0 * (e_c - e_u)).reshape(B, T, 4, 64, 64)
        zs = torch.empty(B, T, 4, 64, 64, device=dev)
        for k, t in enumerate(tl):
            idx = T - k - 1
            a_t = ab[t]
            a_p = ab[t - ratio] if t - ratio >= 0 else ab[0]
            var = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
            xt = xts[:, T - k]
            x0_hat = (xt - (1 - a_t).sqrt() * npred[:, k]) / a_t.sqrt()
            mu = a_p.sqrt() * x0_hat + (1 - a_p - var).sqrt() * npred[:, k]
            zs[:, idx] = (xts[:, idx] - mu) / var.sqrt()
        ev[2].record()
        ed, rc = eng.edit(xts[:, T].contiguous(), zs, ctx, ts, coef, [1.0, 5.0, 7.5], plan, 0.1, 1, False, 1)
        ev[3].record()
        out = dec.decode_tensor(ed * (1 / 0.18215))
        ev[4].record()
        torch.cuda.synchronize()
        return out, [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]

    pipeline()
    out, ms = pipeline()
    total = sum(ms)
    print(json.dumps({"workload": f"image -> image, SD-1.5 512^2, DDPM inversion + {T}-step implicit h-Edit + P2P, batch {B}, 1 GPU, all stages native",
                      "images_per_s": B / (total / 1e3), "ms": {"vae_encode": ms[0], "ddpm_inversion": ms[1], "edit": ms[2], "vae_decode": ms[3]},
                      "finite": bool(torch.isfinite(out).all())}))


if __name__ == "__main__":
    main()
