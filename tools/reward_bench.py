#!/usr/bin/env python
"""Face-swapping reward networks (IR-SE50 identity loss, LPIPS-VGG16): loss + image gradient of a batch of 256x256 images, native kernels
(csrc/reward.cu, CUDA-graph replay) against torch autograd on the same weights (cuDNN, fp32 and TF32/fp16-autocast), same box."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hedit_b200 import reward, reward_nets

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, 3, 256, 256, generator=g) * 0.4).clamp(-1, 1).to(dev)
ref = (torch.randn(1, 3, 256, 256, generator=g) * 0.4).clamp(-1, 1).to(dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def autograd_of(loss):
    def fn():
        with torch.enable_grad():
            xx = x.detach().clone().requires_grad_(True)
            return torch.autograd.grad(loss(xx).sum(), xx)[0]
    return fn


out = {"batch": B}
idl = reward_nets.SyntheticIDLoss(ref, 0).to(dev)
idl.facenet.to(memory_format=torch.channels_last)
arc = reward.ArcFaceEngine.from_facenet(idl.facenet); arc.set_reference(ref)
gb, lb = torch.empty_like(x), torch.empty(B, device=dev)
with torch.no_grad():
    rf = reward_nets.id_features(idl.facenet, ref)
id_loss = lambda t: 1 - F.cosine_similarity(rf, reward_nets.id_features(idl.facenet, t), dim=-1)
out["arcface_native_ms"] = timeit(lambda: arc.loss_grad(x, grad_out=gb, loss_out=lb))
out["arcface_gflop"] = arc.last_stats["flops"] / 1e9
out["arcface_native_tflops"] = out["arcface_gflop"] / out["arcface_native_ms"]
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
out["arcface_torch_fp32_ms"] = timeit(autograd_of(id_loss), 5)
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
out["arcface_torch_tf32_ms"] = timeit(autograd_of(id_loss), 5)

vgg = reward_nets._seed_init(reward_nets.LPIPSVGG16(), 1).to(dev).to(memory_format=torch.channels_last)
lp = reward.LpipsEngine.from_module(vgg); lp.set_source(ref)
with torch.no_grad():
    taps = [f / (f.pow(2).sum(1, keepdim=True).sqrt() + 1e-10) for f in vgg.taps(ref)]
out["lpips_native_ms"] = timeit(lambda: lp.loss_grad(x, grad_out=gb, loss_out=lb))
out["lpips_gflop"] = lp.last_stats["flops"] / 1e9
out["lpips_native_tflops"] = out["lpips_gflop"] / out["lpips_native_ms"]
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
out["lpips_torch_fp32_ms"] = timeit(autograd_of(lambda t: vgg(t, taps)), 5)
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
out["lpips_torch_tf32_ms"] = timeit(autograd_of(lambda t: vgg(t, taps)), 5)


def fp16_autocast():
    with torch.enable_grad(), torch.autocast("cuda", dtype=torch.float16):
        xx = x.detach().clone().requires_grad_(True)
        return torch.autograd.grad(vgg(xx, taps).sum(), xx)[0]


out["lpips_torch_fp16_autocast_ms"] = timeit(fp16_autocast, 5)
print(json.dumps(out))
