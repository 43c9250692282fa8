"""Test-only stand-in for the `lpips` package (richzhang/PerceptualSimilarity 0.1.x; a dependency of the reference's
face-swapping/arcface/arcface_model.py:8 that is neither vendored under /root/reference nor installable offline), so that the reference's
`LPIPS_Loss` class imports and runs UNMODIFIED when goldens are generated (tests/make_golden.py --config face_full).

`LPIPS(net='vgg')` follows the package's published algorithm -- ScalingLayer, torchvision VGG16 feature slices at relu1_2 / 2_2 / 3_3 /
4_3 / 5_3, channel-unit-normalisation with eps 1e-10, squared difference, 1x1 `lin` heads, spatial average, sum over the five taps --
through hedit_b200.reward_nets.LPIPSVGG16 (pinned tap by tap to torchvision's vgg16 in tests/test_oracle_pin.py).  Weights are whatever
the caller loads (seeded random in the tests: the trained weights do not exist offline)."""
import torch
import torch.nn as nn

from hedit_b200.reward_nets import LPIPSVGG16


class LPIPS(nn.Module):
    def __init__(self, net="vgg", **_):
        super().__init__()
        assert net in ("vgg", "vgg16"), "the shim restates the VGG16 variant only"
        self.pnet_type, self.spatial, self.lpips = "vgg", False, True
        self.net = LPIPSVGG16()

    def state_dict(self, *a, **k):
        return self.net.state_dict(*a, **k)

    def forward(self, in0, in1, normalize=False):
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        taps1 = [f / (f.pow(2).sum(1, keepdim=True).sqrt() + 1e-10) for f in self.net.taps(in1)]
        return self.net(in0, taps1).reshape(-1, 1, 1, 1)
