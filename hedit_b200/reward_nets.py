"""Face-swapping reward networks at their real geometry, for callers that have no pretrained weights at hand (benchmarks, tests).

The reference differentiates two frozen networks 600 times per image (face-swapping/inversion/h_edit_R.py:109-110,128-129):
  * ArcFace identity loss  -- `IDLoss.get_cosine_loss` (arcface/arcface_model.py:12-70): crop [35:223, 32:220] of the 256x256 image,
    adaptive-average-pool to 112x112, IR-SE50 backbone (arcface/facial_recognition/model_irse.py:9-84) -> 512-d feature, 1 - cosine
    similarity with the reference face's feature;
  * LPIPS-VGG16 perceptual loss -- `LPIPS_Loss.get_lpips_loss` (arcface_model.py:72-94; the `lpips` package, net='vgg'): ImageNet-style input
    scaling, VGG16 conv features at relu1_2 / 2_2 / 3_3 / 4_3 / 5_3, channel-unit-normalised, squared difference, non-negative 1x1 "lin"
    weights, spatial mean, sum over the five taps.
In production the caller hands `h_Edit_R` its own `idloss` / `lpipsloss` objects (weights: model_ir_se50.pth, the lpips package).  None of
those weights exist offline, so this module builds the same architectures with seeded random weights; `make_reward_grads` wires them to
the gradient hooks `FaceUNetEngine.edit` takes.  torch modules (cuDNN / cuBLAS through autograd): the native loop calls them as
reward-model plug-ins exactly like the caller's own objects."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _SE(nn.Module):
    def __init__(self, c, r=16):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c // r, 1, bias=False)
        self.fc2 = nn.Conv2d(c // r, c, 1, bias=False)

    def forward(self, x):
        w = torch.sigmoid(self.fc2(F.relu(self.fc1(x.mean(dim=(2, 3), keepdim=True)))))
        return x * w


class _IRSEUnit(nn.Module):
    """bottleneck_IR_SE (helpers.py:100-125): BN -> 3x3 -> PReLU -> 3x3 (stride) -> BN -> SE, plus a strided-identity (MaxPool2d(1, s)) or
    1x1-conv+BN shortcut.  Attribute names give the reference's state_dict keys (`shortcut_layer.0.weight`, `res_layer.5.fc1.weight`)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.shortcut_layer = nn.MaxPool2d(1, stride) if cin == cout else nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))
        self.res_layer = nn.Sequential(nn.BatchNorm2d(cin), nn.Conv2d(cin, cout, 3, 1, 1, bias=False), nn.PReLU(cout),
                                       nn.Conv2d(cout, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout), _SE(cout))

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class IRSE50(nn.Module):
    """IR-SE50 face-recognition backbone (model_irse.py:9-56 with num_layers=50, mode='ir_se'): 112x112 input, stages of (3, 4, 14, 3)
    units at 64/128/256/512 channels, 512-d l2-normalised output.  state_dict-compatible with the reference's `Backbone`."""

    def __init__(self):
        super().__init__()
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        self.output_layer = nn.Sequential(nn.BatchNorm2d(512), nn.Dropout(0.6), nn.Flatten(), nn.Linear(512 * 7 * 7, 512), nn.BatchNorm1d(512))
        units, cin = [], 64
        for cout, n in ((64, 3), (128, 4), (256, 14), (512, 3)):
            for k in range(n):
                units.append(_IRSEUnit(cin, cout, 2 if k == 0 else 1))
                cin = cout
        self.body = nn.Sequential(*units)

    def forward(self, x):
        f = self.output_layer(self.body(self.input_layer(x)))
        return f / f.norm(dim=1, keepdim=True)            # l2_norm (helpers.py:15-18)


class LPIPSVGG16(nn.Module):
    CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512)
    TAPS = (64, 128, 256, 512, 512)

    def __init__(self):
        super().__init__()
        layers, cin = [], 3
        for v in self.CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, 1, 1), nn.ReLU()]
                cin = v
        self.features = nn.Sequential(*layers)
        self.lins = nn.ParameterList([nn.Parameter(torch.rand(1, c, 1, 1) * 0.02) for c in self.TAPS])      # non-negative, like the trained heads
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188]).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor([.458, .448, .450]).view(1, 3, 1, 1))

    def taps(self, x):
        x = (x - self.shift) / self.scale
        out = []
        for m in self.features:
            if isinstance(m, nn.MaxPool2d):
                out.append(x)
            x = m(x)
        out.append(x)
        return out            # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3

    def forward(self, x, y_taps):
        d = 0.0
        for fx, fy, w in zip(self.taps(x), y_taps, self.lins):
            nx = fx / (fx.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            d = d + ((nx - fy) ** 2 * w).sum(1, keepdim=True).mean(dim=(2, 3))
        return d.reshape(-1)


def _seed_init(m: nn.Module, seed: int):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, (nn.Conv2d, nn.Linear)):
                fan = mod.weight[0].numel()
                mod.weight.copy_((torch.rand(mod.weight.shape, generator=g) * 2 - 1) * (3.0 / fan) ** 0.5)
                if mod.bias is not None:
                    mod.bias.copy_((torch.rand(mod.bias.shape, generator=g) * 2 - 1) * 0.05)
            elif isinstance(mod, (nn.BatchNorm2d, nn.BatchNorm1d)):
                mod.weight.copy_(0.8 + 0.4 * torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_((torch.rand(mod.bias.shape, generator=g) * 2 - 1) * 0.1)
                mod.running_mean.copy_((torch.rand(mod.bias.shape, generator=g) * 2 - 1) * 0.1)
                mod.running_var.copy_(0.8 + 0.4 * torch.rand(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.PReLU):
                mod.weight.copy_(0.1 + 0.3 * torch.rand(mod.weight.shape, generator=g))
        for p in getattr(m, "lins", []):           # LPIPS 1x1 heads: non-negative, like the trained ones (the constructor draws them unseeded)
            p.copy_(torch.rand(p.shape, generator=g) * 0.02)
    return m.eval().requires_grad_(False)


class SyntheticIDLoss(nn.Module):
    """Same protocol and attribute layout as the reference's IDLoss (arcface_model.py:12-70: `.facenet`, `.ref`, `get_cosine_loss`), with
    seeded random weights and a given reference face instead of model_ir_se50.pth / a jpg."""

    def __init__(self, ref_img: torch.Tensor, seed: int = 0):
        super().__init__()
        self.facenet = _seed_init(IRSE50(), seed)
        self.ref = ref_img.detach().reshape(-1, 3, 256, 256)[:1].clone()

    def extract_feats(self, x):
        return id_features(self.facenet, x)

    def get_cosine_loss(self, image):
        return (1 - F.cosine_similarity(F.normalize(self.extract_feats(self.ref.to(image.device)), dim=-1),
                                        F.normalize(self.extract_feats(image), dim=-1), dim=-1)).mean()


class SyntheticLPIPSLoss(nn.Module):
    """Same protocol and attribute layout as the reference's LPIPS_Loss (arcface_model.py:72-95: `.lpips_loss`, `.src`, `get_lpips_loss`)."""

    def __init__(self, src_img: torch.Tensor, seed: int = 1):
        super().__init__()
        self.lpips_loss = _seed_init(LPIPSVGG16(), seed)
        self.src = src_img.detach().clone()

    def get_lpips_loss(self, x):
        src = self.src.to(x.device)
        with torch.no_grad():
            taps = [f / (f.pow(2).sum(1, keepdim=True).sqrt() + 1e-10) for f in self.lpips_loss.taps(src)]
        return self.lpips_loss(x, taps).mean()


def id_features(net: IRSE50, img: torch.Tensor) -> torch.Tensor:
    """IDLoss.extract_feats (arcface_model.py:41-47) for 256x256 inputs."""
    x = img if img.shape[2] == 256 else F.adaptive_avg_pool2d(img, (256, 256))
    return net(F.adaptive_avg_pool2d(x[:, :, 35:223, 32:220], (112, 112)))


def make_reward_grads(ref_img: torch.Tensor, src_img: torch.Tensor, device, seed: int = 0):
    """(id_grad, lpips_grad, description): x0 (B,3,256,256) -> d/dx0 of the per-image-summed identity / LPIPS losses, as
    `FaceUNetEngine.edit(..., id_grad=, lpips_grad=)` expects.  ref_img (1,3,256,256) = the face whose identity is transferred,
    src_img (B,3,256,256) = the images being edited (LPIPS anchors)."""
    arc = _seed_init(IRSE50(), seed).to(device).to(memory_format=torch.channels_last)
    vgg = _seed_init(LPIPSVGG16(), seed + 1).to(device).to(memory_format=torch.channels_last)
    with torch.no_grad():
        ref_feat = id_features(arc, ref_img.to(device))
        src_taps = [f / (f.pow(2).sum(1, keepdim=True).sqrt() + 1e-10) for f in vgg.taps(src_img.to(device))]

    def grad_of(loss):
        def fn(x0):
            with torch.enable_grad():
                x = x0.detach().clone().requires_grad_(True)
                return torch.autograd.grad(loss(x), x)[0]
        return fn

    id_grad = grad_of(lambda x: (1 - F.cosine_similarity(ref_feat, id_features(arc, x), dim=-1)).sum())
    lp_grad = grad_of(lambda x: vgg(x, src_taps).sum())
    return id_grad, lp_grad, "IR-SE50 (112x112 crop) + LPIPS-VGG16 (256x256) at full geometry, seeded random weights, torch fp32 autograd (cuDNN)"
