#!/usr/bin/env python
"""Micro-benchmarks of single operators through the C ABI (CUDA events, device-resident inputs).  Used for tuning and
as the target of `ncu` captures.  python tools/op_bench.py [attn|linear|conv|all] [--iters N]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import P, bf, lib  # noqa: E402

DEV = "cuda"


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_attn(iters, S=8):
    for (S, N, H, d) in [(S, 4096, 8, 40), (S, 1024, 8, 80), (S, 256, 8, 160)]:
        C = H * d
        qkv = bf(torch.randn(S, N, 3 * C, device=DEV))
        out = torch.zeros(S, N, C, device=DEV, dtype=qkv.dtype)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        fn = lambda: lib().hedit_op_self_attention(P(q), P(k), P(v), 3 * C, 3 * C, S, N, N, H, d, None, None, None, P(out), None)
        ms = timeit(fn, iters)
        fl = 4.0 * S * H * N * N * d
        # accuracy of the same launch against fp32 torch on the first two samples
        qf, kf, vf = (t[:2].float().reshape(2, N, H, d).permute(0, 2, 1, 3) for t in (q, k, v))
        ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) @ vf).permute(0, 2, 1, 3).reshape(2, N, C)
        err = ((out[:2].float() - ref).norm() / ref.norm()).item()
        print(f"self_attn S={S} N={N} H={H} d={d}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  ({S * H * (N / 128) ** 2 / 148 :.0f} 128x128 blocks/SM, "
              f"{ms * 1e-3 * 1.9e9 / (S * H * (N / 128) ** 2 / 148):.0f} cyc/block @1.9GHz)  rel-err vs fp32 {err:.2e}  [HEDIT_ATTN_V4={os.environ.get('HEDIT_ATTN_V4', '0')}]")


def bench_linear(iters):
    for (M, N, K, res, f32) in [(163840, 960, 320, False, False), (163840, 320, 320, True, True), (40960, 1920, 640, False, False),
                                (40960, 640, 2560, True, True), (163840, 320, 1280, True, True), (10240, 1280, 1280, True, True)]:
        A, W = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV) * K ** -0.5)
        bias = torch.randn(N, device=DEV) if f32 or res else None      # fp16-out projections (q/k/v) carry no bias
        r = torch.randn(M, N, device=DEV) if res else None
        o32 = torch.empty(M, N, device=DEV) if f32 else None
        o16 = None if f32 else torch.empty(M, N, device=DEV, dtype=A.dtype)
        fn = lambda: lib().hedit_op_linear(P(A), P(W), P(bias), P(r), P(o32), P(o16), M, N, K, None)
        ms = timeit(fn, iters)
        print(f"linear M={M} N={N} K={K} res={res} f32out={f32}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


def bench_geglu(iters):
    """Feed-forward GEGLU projections of the three UNet levels at 40 samples."""
    for (M, N2, K) in [(163840, 2560, 320), (40960, 5120, 640), (10240, 10240, 1280)]:
        A, W = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N2, K, device=DEV) * K ** -0.5)
        bias = torch.randn(N2, device=DEV)
        out = torch.empty(M, N2 // 2, device=DEV, dtype=A.dtype)
        fn = lambda: lib().hedit_op_linear_geglu(P(A), P(W), P(bias), P(out), M, N2, K, None)
        ms = timeit(fn, iters)
        print(f"geglu M={M} N2={N2} K={K}: {ms:.3f} ms  {2.0 * M * N2 * K / ms / 1e9:.1f} TFLOP/s")


def bench_conv(iters):
    for (S, H, C, Co, st) in [(16, 64, 320, 320, 1), (16, 32, 640, 640, 1), (16, 16, 1280, 1280, 1), (16, 8, 2560, 1280, 1), (16, 64, 960, 320, 1)]:
        x = bf(torch.randn(S, H, H, C, device=DEV))
        w = bf(torch.randn(Co, 3, 3, C, device=DEV) * (9 * C) ** -0.5)
        bias = torch.randn(Co, device=DEV)
        out = torch.empty(S, H // st, H // st, Co, device=DEV)
        fn = lambda: lib().hedit_op_conv3x3(P(x), P(w), P(bias), P(out), S, H, H, C, Co, st, None)
        ms = timeit(fn, iters)
        print(f"conv3x3 S={S} {H}x{H} C={C}->{Co}: {ms:.3f} ms  {2.0 * S * (H // st) ** 2 * Co * 9 * C / ms / 1e9:.1f} TFLOP/s")


def bench_cross(iters):
    for (S, N, H, d, pairs) in [(40, 4096, 8, 40, False), (40, 4096, 8, 40, True), (40, 1024, 8, 80, True)]:
        C = H * d
        n_ctx = 17
        q = bf(torch.randn(S, N, C, device=DEV))
        kv = bf(torch.randn(n_ctx, 77, 2 * C, device=DEV))
        ctx_idx = torch.randint(0, n_ctx, (S,), dtype=torch.int32, device=DEV)
        if pairs:      # per image: 3 singles + one (source, target) pair
            us0, us1, uimg = [], [], []
            for b in range(S // 5):
                s = 5 * b
                us0 += [s, s + 1, s + 4, s + 2]; us1 += [-1, -1, -1, s + 3]; uimg += [b] * 4
        else:
            us0, us1, uimg = list(range(S)), [-1] * S, [0] * S
        t = lambda v, dt=torch.int32: torch.tensor(v, dtype=dt, device=DEV)
        us0, us1, uimg = t(us0), t(us1), t(uimg)
        nimg = S // 5
        mapper = torch.randint(0, 77, (nimg, 80), dtype=torch.int32, device=DEV)
        cb, ct = torch.rand(nimg, 80, device=DEV), torch.rand(nimg, 80, device=DEV)
        isr = torch.zeros(nimg, dtype=torch.int32, device=DEV)
        out = torch.zeros(S, N, C, device=DEV, dtype=q.dtype)
        fn = lambda: lib().hedit_op_cross_attention_p2p(P(q), P(kv), S, n_ctx, N, H, d, len(us0), P(us0), P(us1), P(uimg), P(ctx_idx), P(mapper),
                                                        P(cb), P(ct), None, P(isr), None, None, -1, 0, P(out), None)
        ms = timeit(fn, iters)
        items = S * H * (N // 128)
        print(f"cross_attn S={S} N={N} H={H} d={d} pairs={pairs}: {ms:.3f} ms  ({items} tile-phases, {ms * 1e-3 * 1.9e9 * 148 / items:.0f} SM-cycles/tile-phase)")


def bench_vae(iters):
    """SD-1.x VAE decoder geometry, 64x64 latent -> 512x512 image: decode and decode+backward (random-init weights)."""
    import hedit_b200
    eng = hedit_b200.VaeDecoderEngine(dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_groups=32))
    eng.load_random_weights(0)
    for B in (1, 4):
        z = torch.randn(B, 4, 64, 64, device=DEV) * 5
        img = eng.decode_tensor(z)
        f_fwd = eng.last_stats["flops"]
        dimg = torch.randn_like(img)
        eng.backward(dimg)
        f_all = eng.last_stats["flops"]
        ms_f = timeit(lambda: eng.decode_tensor(z), iters)
        ms_fb = timeit(lambda: (eng.decode_tensor(z), eng.backward(dimg)), iters)
        print(f"vae B={B}: decode {ms_f:.2f} ms ({f_fwd / ms_f / 1e9:.0f} TFLOP/s, {f_fwd / B / 1e12:.3f} TFLOP/image) | decode+backward {ms_fb:.2f} ms "
              f"({f_all / ms_fb / 1e9:.0f} TFLOP/s) | backward alone ~{ms_fb - ms_f:.2f} ms ({(f_all - f_fwd) / max(ms_fb - ms_f, 1e-6) / 1e9:.0f} TFLOP/s)")


def bench_face(iters):
    """CelebA-HQ DDPM UNet geometry (ch 128 x (1,1,2,2,4,4), 256x256), random-init weights: one denoiser call."""
    import hedit_b200
    eng = hedit_b200.FaceUNetEngine(dict(ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, attn_resolution=16, image_size=256, in_channels=3, out_ch=3))
    eng.load_random_weights(0)
    for S in (1, 8):
        x = torch.randn(S, 3, 256, 256, device=DEV)
        eng(x, 500.0)
        fl = eng.last_stats["flops"]
        ms = timeit(lambda: eng(x, 500.0), iters)
        print(f"face unet S={S}: {ms:.2f} ms  {fl / ms / 1e9:.0f} TFLOP/s  ({fl / S / 1e12:.3f} TFLOP/sample, {eng.last_stats['kernel_launches']} launches)")
        try:                     # the same launches replayed from a CUDA graph: how much of the time is launch gaps
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eng(x, 500.0)
            msg = timeit(g.replay, iters)
            print(f"   graph replay: {msg:.2f} ms  {fl / msg / 1e9:.0f} TFLOP/s")
        except Exception as ex:
            print("   graph capture failed:", ex)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--samples", type=int, default=8, help="attn: samples per launch (40 = the bench's UNet launch: 17.3 CTA waves instead of 3.5)")
    a = ap.parse_args()
    torch.manual_seed(0)
    if a.what in ("attn", "all"):
        bench_attn(a.iters, a.samples)
    if a.what in ("linear", "all"):
        bench_linear(a.iters)
    if a.what in ("conv", "all"):
        bench_conv(a.iters)
    if a.what in ("cross", "all"):
        bench_cross(a.iters)
    if a.what in ("vae", "all"):
        bench_vae(a.iters)
    if a.what in ("geglu", "all"):
        bench_geglu(a.iters)
    if a.what in ("face", "all"):
        bench_face(a.iters)
